// CTA-pair (cta_group::2) tcgen05 implicit GEMM - the production kernel for every dense contraction of the
// HybridViT forward (problem description and A-operand modes: IgemmParams in kernels.h).
//
// Why a CTA pair: with cta_group::1 a 128x256x16 MMA reads 12 KB of operands from shared memory in 128 cycles
// (96 B/clk) while TMA refills the ring at 94 B/clk - together above the 128 B/clk/SM shared-memory port.  Here one
// UMMA of M = 256 spans two SMs: each CTA stages its own 128 rows of A and only HALF of the B tile, i.e. per SM
// 64 B/clk of operand reads + 62 B/clk of TMA fill, and the L2 -> SM traffic per FLOP drops by a third
// (measured: 1.50 PFLOP/s on a 31744x4096x4096 fp16 GEMM vs 1.35 for the single-CTA kernel).
//
// Why a TMA-store epilogue: row-per-thread global stores (32 distinct 128-byte lines per warp instruction) cost more
// than the whole K = 512 main loop (qkv projection: 86 us, 41 us with the stores removed).  The epilogue therefore
// writes 128-byte-wide column blocks of the tile into 128B-swizzled shared memory and one thread issues a TMA store;
// the fp32 residual of proj / fc2 / patch-embedding is TMA-loaded into the same staging buffer and added in place.
//
// Pair protocol (per smem stage s / accumulator stage a):
//   full[s]       lives in the LEADER (cluster rank 0); the leader's producer posts expect_tx for the bytes of both
//                 CTAs, both producers' TMA loads complete_tx on it (cp.async.bulk.tensor ... .cta_group::2)
//   empty[s]      one per CTA; the leader's MMA thread frees the stage in both CTAs with a multicast tcgen05.commit
//   tmem_full[a]  one per CTA, arrived by a multicast commit after the last k-block
//   tmem_empty[a] lives in the leader; the 8 epilogue warps of BOTH CTAs arrive on it (remote mbarrier arrive)
// Tiles are enumerated in pairs of adjacent 128-row M-tiles (rank r takes tile 2i+r) that share the N-tile; pairs never
// straddle an upsample parity class; an odd leftover tile is padded with an out-of-range tile (TMA zero fill on load,
// clipped on store).
//
// Epilogue organisation: 8 warps = 2 groups x 4 warps (a group covers the four TMEM lane quarters).  Group g owns
// the column blocks j = g, g+2, ... of the tile (a block is 128 bytes of every output row: 64 16-bit or 32 fp32
// columns), with two 16 KB staging buffers, its own named barrier and its own store-issuing thread.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace hvit {

namespace {

constexpr int BLOCK_M = 128;  // rows per CTA (UMMA M = 256 across the pair)
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;
constexpr int STG_BUF_BYTES = 128 * 128;  // 128 rows x 128 B

template <int BLOCK_N, int EPI = 0>
struct Cfg2 {
  static constexpr int B_HALF_ROWS = BLOCK_N / 2;
  static constexpr int B_STAGE_BYTES = B_HALF_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // The K = 512 GEMM main loops are bound by the ring's round trip (MMA done -> empty -> TMA issue -> L2 -> full:
  // ~2 200 cycles against 512 MMA cycles per stage), i.e. by the bytes in flight, not by L2 bandwidth (a variant with
  // the weights resident in shared memory - half the L2 traffic, 4 A stages - was no faster).  Measured: 3 -> 4
  // stages +10 %, 4 -> 5 +5 %.  So shared memory goes to stages: the per-channel constants live in shared memory only
  // for the CURRENT tile (two small buffers, by accumulator parity), and where the epilogue is far from critical -
  // the plain 16-bit one (EPI_LIN16: qkv, to_feature_map, skip projections) and the fp32 one of long-K GEMMs
  // (EPI_F32D: fc2) - each epilogue group gives up its second staging buffer (it waits for the previous TMA store
  // to release the buffer).
  static constexpr bool DEEP = BLOCK_N == 256 && (EPI == 1 || EPI == 8 || EPI == 9);
  static constexpr int NBUF = DEEP ? 1 : 2;                   // staging buffers per epilogue group
  static constexpr bool HAS_SCALE = EPI == 0 || EPI == 3 || EPI == 4;  // epilogues with a per-channel multiplier
  static constexpr int STAGES = BLOCK_N == 256 ? (HAS_SCALE ? 4 : (DEEP ? 6 : 5)) : (BLOCK_N == 128 ? 6 : 7);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int STAGING_BYTES = 2 * NBUF * STG_BUF_BYTES;    // [group][buffer]
  static constexpr int CONST_N = BLOCK_N;                           // shift (| scale) of one tile, x 2 buffers
  static constexpr int CONST_BUF = (HAS_SCALE ? 2 : 1) * CONST_N;   // floats per buffer
  static constexpr int CONST_BYTES = 2 * CONST_BUF * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + CONST_BYTES + 512;
  static_assert(SMEM_BYTES <= 232448, "igemm_tc2: shared memory budget");
};

struct TileCoord {
  int n0, m0, b, h0, w0, par;
};

// cluster tile -> this CTA's 128-row tile.  group = upsample parity (IG_UP2) or 0.
__device__ __forceinline__ TileCoord decode_ctile(const IgemmParams& p, int ct, int n_tiles_n, int block_n,
                                                  int pairs_per_group, int rank) {
  TileCoord c;
  int pr, nt, g, ig;
  fast_divmod(p.fd_ntn, ct, pr, nt);
  c.n0 = nt * block_n;
  fast_divmod(p.fd_ppg, pr, g, ig);
  int i = 2 * ig + rank;
  c.m0 = 0; c.b = 0; c.h0 = 0; c.w0 = 0; c.par = g;
  if (p.mode == IG_PLAIN) {
    c.m0 = i * BLOCK_M;  // >= M for the padding tile: zero fill on load, clipped on store
  } else {
    int tw, th;
    fast_divmod(p.fd_tw, i, i, tw);
    c.w0 = tw * p.Wt;
    fast_divmod(p.fd_th, i, i, th);
    c.h0 = th * p.Ht;
    c.b = i;  // == B for the padding tile
  }
  return c;
}


// Epilogue specialisations.  The hot combinations of the HybridViT plan are compiled without run-time branches in
// the per-element code (so the 32 columns of a TMEM chunk are scheduled together); any other parameter combination
// takes EPI_GENERIC.
enum {
  EPI_GENERIC = 0,   // everything decided at run time
  EPI_LIN16 = 1,     // + shift,            16-bit out (qkv, to_feature_map, skip projections)
  EPI_GELU16 = 2,    // + shift, erf-GELU,  16-bit out (fc1)
  EPI_BN16 = 3,      // * scale + shift, max(., lo), 16-bit out (3x3 convs; lo = 0 for ReLU, -inf for none)
  EPI_BNPOOL16 = 4,  // EPI_BN16 + fused 2x2 max-pool
  EPI_F32 = 5,       // + shift (+ fp32 residual), fp32 out (proj, fc2, patch embedding)
  EPI_SH16 = 6,      // + shift, max(., lo), 16-bit out: 3x3 convs whose BN scale is folded into the weights
  EPI_SHPOOL16 = 7,  // EPI_SH16 + fused 2x2 max-pool
  EPI_F32D = 8,      // EPI_F32 without a TMA-loaded residual and K >= 1024 (fc2): deep-ring configuration
  EPI_LNLIN16 = 9,   // EPI_LIN16 with the preceding LayerNorm folded in (IgemmParams::ln_stats_in): qkv, to_feature_map
  EPI_LNGELU16 = 10  // EPI_GELU16 with the preceding LayerNorm folded in: fc1
};

// GELU(v) = relu(v) - 0.5*|v|*erfc(|v|/sqrt2), erfc(u/sqrt2) = 2^q(u) with a weighted-minimax degree-5 q on [0, 6]
// (max |GELU error| 6.9e-7 in fp32 evaluation, verified against scipy - DESIGN.md): 10 instructions per element.
__device__ __forceinline__ float gelu_erfc5(float v) {
  const float u = fminf(fabsf(v), 6.0f);
  float q = -4.837184678763151e-4f;
  q = fmaf(q, u, 7.163475267589092e-3f);
  q = fmaf(q, u, -5.204327404499054e-2f);
  q = fmaf(q, u, -4.5973172783851624e-1f);
  q = fmaf(q, u, -1.150922417640686f);
  q = fmaf(q, u, -1.5133146916923579e-5f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  return fmaf(-0.5f * u, e, fmaxf(v, 0.0f));
}
// The same on a pair of elements with packed fp32 arithmetic (FFMA2 / FMUL2: one issue slot for two elements; the
// fc1 epilogue competes with the MMA and TMA threads for issue slots): 6.5 instead of 10 instructions per element,
// bit-identical results (same operations, same order).
__device__ __forceinline__ float2 gelu_erfc5_x2(float2 v) {
  const float2 u = make_float2(fminf(fabsf(v.x), 6.0f), fminf(fabsf(v.y), 6.0f));
  float2 q = make_float2(-4.837184678763151e-4f, -4.837184678763151e-4f);
  q = __ffma2_rn(q, u, make_float2(7.163475267589092e-3f, 7.163475267589092e-3f));
  q = __ffma2_rn(q, u, make_float2(-5.204327404499054e-2f, -5.204327404499054e-2f));
  q = __ffma2_rn(q, u, make_float2(-4.5973172783851624e-1f, -4.5973172783851624e-1f));
  q = __ffma2_rn(q, u, make_float2(-1.150922417640686f, -1.150922417640686f));
  q = __ffma2_rn(q, u, make_float2(-1.5133146916923579e-5f, -1.5133146916923579e-5f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
  return __ffma2_rn(__fmul2_rn(u, make_float2(-0.5f, -0.5f)), e, make_float2(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f)));
}

__device__ __forceinline__ void tile_out_coords(const IgemmParams& p, const TileCoord& c, bool pool, int& o1, int& o2,
                                                int& o3) {
  if (p.mode == IG_PLAIN) {
    o1 = c.m0; o2 = 0; o3 = 0;
  } else if (pool) {
    o1 = c.w0 >> 1; o2 = c.h0 >> 1; o3 = c.b;
  } else {
    o1 = c.w0; o2 = c.h0; o3 = c.b;
  }
}

// max of two packed 16-bit pairs (fp16x2 or bf16x2)
template <bool F16>
__device__ __forceinline__ uint32_t max_16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  if (F16) asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// 32 accumulator columns -> 64 bytes of this thread's 128-byte staging row (16-bit output).
// sc / sh: shared-memory per-channel constants of these 32 columns; half: which 64-byte half of the row.
template <int EPI, bool F16>
__device__ __forceinline__ void emit16(const uint32_t (&cur)[32], const float* sc, const float* sh,
                                       const IgemmParams& p, float relu_lo, uint8_t* row, int half, int sw,
                                       bool writer, int hxor = 16, float ln_rs = 0.f) {
  float v[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(sh + 4 * q);
    if (EPI == EPI_LNLIN16 || EPI == EPI_LNGELU16) {
      // folded LayerNorm: rs_m * acc + c_n (the weights are gamma-scaled AND centred over k, so acc is already the
      // contraction of x - mean(x))
      const float2 r2 = make_float2(ln_rs, ln_rs);
      const float2 s0 = __ffma2_rn(make_float2(__uint_as_float(cur[4 * q + 0]), __uint_as_float(cur[4 * q + 1])), r2,
                                   make_float2(b.x, b.y));
      const float2 s1 = __ffma2_rn(make_float2(__uint_as_float(cur[4 * q + 2]), __uint_as_float(cur[4 * q + 3])), r2,
                                   make_float2(b.z, b.w));
      v[4 * q + 0] = s0.x; v[4 * q + 1] = s0.y; v[4 * q + 2] = s1.x; v[4 * q + 3] = s1.y;
    } else if (EPI == EPI_LIN16 || EPI == EPI_GELU16 || EPI == EPI_SH16 || EPI == EPI_SHPOOL16) {
      const float2 s0 = __fadd2_rn(make_float2(__uint_as_float(cur[4 * q + 0]), __uint_as_float(cur[4 * q + 1])),
                                   make_float2(b.x, b.y));
      const float2 s1 = __fadd2_rn(make_float2(__uint_as_float(cur[4 * q + 2]), __uint_as_float(cur[4 * q + 3])),
                                   make_float2(b.z, b.w));
      v[4 * q + 0] = s0.x; v[4 * q + 1] = s0.y; v[4 * q + 2] = s1.x; v[4 * q + 3] = s1.y;
    } else {
      const float4 a = *reinterpret_cast<const float4*>(sc + 4 * q);
      v[4 * q + 0] = fmaf(__uint_as_float(cur[4 * q + 0]), a.x, b.x);
      v[4 * q + 1] = fmaf(__uint_as_float(cur[4 * q + 1]), a.y, b.y);
      v[4 * q + 2] = fmaf(__uint_as_float(cur[4 * q + 2]), a.z, b.z);
      v[4 * q + 3] = fmaf(__uint_as_float(cur[4 * q + 3]), a.w, b.w);
    }
  }
  if (EPI == EPI_BNPOOL16 || EPI == EPI_SHPOOL16) {
    // 2x2 max-pool on the PACKED 16-bit values: rounding is monotonic, so max(round(a), round(b)) == round(max(a, b))
    // and relu commutes with both - bit-identical to pooling in fp32, with half the shuffles (the warp shuffle unit,
    // ~1 instruction per clock and SM, bounds this epilogue: 1 024 -> 512 shuffles per 128 x 128 tile)
    uint32_t u[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) u[e] = F16 ? pack_f16x2(v[2 * e], v[2 * e + 1]) : pack_bf16x2(v[2 * e], v[2 * e + 1]);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      u[e] = max_16x2<F16>(u[e], __shfl_xor_sync(0xFFFFFFFFu, u[e], 1));
      u[e] = max_16x2<F16>(u[e], __shfl_xor_sync(0xFFFFFFFFu, u[e], hxor));
    }
    if (relu_lo == 0.0f) {
#pragma unroll
      for (int e = 0; e < 16; ++e) u[e] = max_16x2<F16>(u[e], 0u);
    }
    if (writer) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(row + (((half * 4 + q) ^ sw) << 4)) =
            make_uint4(u[4 * q + 0], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
    }
    return;
  }
  if (EPI == EPI_GELU16 || EPI == EPI_LNGELU16) {
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      const float2 g = gelu_erfc5_x2(make_float2(v[e], v[e + 1]));
      v[e] = g.x;
      v[e + 1] = g.y;
    }
  } else if (EPI == EPI_BN16 || EPI == EPI_BNPOOL16 || EPI == EPI_SH16 || EPI == EPI_SHPOOL16) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], relu_lo);
  } else if (EPI == EPI_GENERIC) {
    if (p.act == ACT_RELU) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.0f);
    } else if (p.act == ACT_GELU) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = gelu_erfc5(v[e]);
    }
  }
  if (EPI == EPI_BNPOOL16 || EPI == EPI_SHPOOL16 || (EPI == EPI_GENERIC && p.pool)) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      v[e] = fmaxf(v[e], __shfl_xor_sync(0xFFFFFFFFu, v[e], 1));
      v[e] = fmaxf(v[e], __shfl_xor_sync(0xFFFFFFFFu, v[e], hxor));
    }
  }
  if (writer) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 pk;
      if (F16) {
        pk.x = pack_f16x2(v[8 * q + 0], v[8 * q + 1]); pk.y = pack_f16x2(v[8 * q + 2], v[8 * q + 3]);
        pk.z = pack_f16x2(v[8 * q + 4], v[8 * q + 5]); pk.w = pack_f16x2(v[8 * q + 6], v[8 * q + 7]);
      } else {
        pk.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]); pk.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
        pk.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]); pk.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
      }
      *reinterpret_cast<uint4*>(row + (((half * 4 + q) ^ sw) << 4)) = pk;
    }
  }
}

// 32 accumulator columns -> this thread's whole 128-byte staging row (fp32 output); the residual tile, when present,
// has been TMA-loaded into the same (swizzled) staging row and is added in place.
// LNP (LayerNorm producer, IgemmParams::ln_stats_out): the final values x_new of these 32 columns also go out as 16-bit
// (x16_dst: this row's 64 bytes, two 32-byte stores = whole sectors) and their mean / centred sum of squares is merged
// into the running statistics (cnt, mean, M2) of this thread's row (Chan's update; two passes within the chunk).
struct LnRun {
  float cnt, mean, m2;
};
__device__ __forceinline__ void st_global_v8(void* dst, const uint32_t (&u)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(u[0]), "r"(u[1]), "r"(u[2]),
               "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}
template <int EPI, bool LNP = false>
__device__ __forceinline__ void emit32(const uint32_t (&cur)[32], const float* sc, const float* sh,
                                       const IgemmParams& p, float relu_lo, uint8_t* row, int sw, bool has_res,
                                       LnRun* ln = nullptr, uint8_t* x16_dst = nullptr) {
  float v[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(sh + 4 * q);
    if (EPI == EPI_F32 || EPI == EPI_F32D) {
      const float2 s0 = __fadd2_rn(make_float2(__uint_as_float(cur[4 * q + 0]), __uint_as_float(cur[4 * q + 1])),
                                   make_float2(b.x, b.y));
      const float2 s1 = __fadd2_rn(make_float2(__uint_as_float(cur[4 * q + 2]), __uint_as_float(cur[4 * q + 3])),
                                   make_float2(b.z, b.w));
      v[4 * q + 0] = s0.x; v[4 * q + 1] = s0.y; v[4 * q + 2] = s1.x; v[4 * q + 3] = s1.y;
    } else {
      const float4 a = *reinterpret_cast<const float4*>(sc + 4 * q);
      v[4 * q + 0] = fmaf(__uint_as_float(cur[4 * q + 0]), a.x, b.x);
      v[4 * q + 1] = fmaf(__uint_as_float(cur[4 * q + 1]), a.y, b.y);
      v[4 * q + 2] = fmaf(__uint_as_float(cur[4 * q + 2]), a.z, b.z);
      v[4 * q + 3] = fmaf(__uint_as_float(cur[4 * q + 3]), a.w, b.w);
    }
  }
  if (EPI == EPI_GENERIC) {
    if (p.act == ACT_RELU) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.0f);
    } else if (p.act == ACT_GELU) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = gelu_erfc5(v[e]);
    }
  }
  if (has_res) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4* d = reinterpret_cast<float4*>(row + ((q ^ sw) << 4));
      const float4 r = *d;
      v[4 * q + 0] += r.x; v[4 * q + 1] += r.y; v[4 * q + 2] += r.z; v[4 * q + 3] += r.w;
      *d = make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<float4*>(row + ((q ^ sw) << 4)) =
          make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  if (LNP) {
    // (packed fp32 pipes: 16 FADD2 for the sum, 16 FADD2 + 16 FFMA2 for the centred squares)
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      a0 = __fadd2_rn(a0, make_float2(v[e], v[e + 1]));
      a1 = __fadd2_rn(a1, make_float2(v[e + 2], v[e + 3]));
    }
    const float mc = ((a0.x + a0.y) + (a1.x + a1.y)) * (1.0f / 32.0f);
    const float2 nm = make_float2(-mc, -mc);
    float2 q0 = make_float2(0.f, 0.f), q1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      const float2 d0 = __fadd2_rn(make_float2(v[e], v[e + 1]), nm);
      const float2 d1 = __fadd2_rn(make_float2(v[e + 2], v[e + 3]), nm);
      q0 = __ffma2_rn(d0, d0, q0);
      q1 = __ffma2_rn(d1, d1, q1);
    }
    const float m2c = (q0.x + q0.y) + (q1.x + q1.y);
    const float tot = ln->cnt + 32.0f;
    const float delta = mc - ln->mean;
    const float f = 32.0f / tot;
    ln->mean = fmaf(delta, f, ln->mean);
    ln->m2 += fmaf(delta * delta, ln->cnt * f, m2c);
    ln->cnt = tot;
    if (x16_dst != nullptr) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t u[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) u[e] = pack_16x2(v[16 * h + 2 * e], v[16 * h + 2 * e + 1], p.f16);
        st_global_v8(x16_dst + 32 * h, u);
      }
    }
  }
}

template <int BLOCK_N, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
igemm_tc2_kernel(const __grid_constant__ IgemmMaps maps, const IgemmParams p, int num_ctiles, int n_tiles_n,
                 int pairs_per_group) {
  using C = Cfg2<BLOCK_N, EPI>;
  const long long k_t0 = clock64();
  // (declared aligned instead of rounding the pointer up by hand: integer arithmetic on the address loses the
  // shared-memory address space and turns every staging store / constant load of the epilogue into a generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B swizzle needs 1024-byte aligned tiles
  uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;
  uint8_t* consts = staging + C::STAGING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(consts + C::CONST_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]   (used in the leader)
  uint64_t* empty_bar = bars + C::STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]        (used in the leader)
  uint64_t* res_bar = tmem_empty + 2;           // [group][buffer]: residual tile landed in the staging buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 4);
  volatile long long* t_issue = reinterpret_cast<volatile long long*>(bars + 32);  // [STAGES] diagnostics: TMA issue time per stage

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int kblocks = p.K / BLOCK_K;
  const int cblocks = p.Cin / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a);
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.out[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * NUM_EPI_WARPS);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits and TMEM allocation of both CTAs visible before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // PDL: the next kernel's prologue may overlap this kernel ...
  griddep_wait();               // ... and everything above overlapped the previous kernel's tail
  const long long k_t1 = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      long long pw = 0;
      const long long pt0 = clock64();
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        const TileCoord c = decode_ctile(p, ct, n_tiles_n, BLOCK_N, pairs_per_group, rank);
        const int py = c.par >> 1, px = c.par & 1;
        const int b_row = c.n0 + c.par * p.N + rank * C::B_HALF_ROWS;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (p.prof) {
            const long long t = clock64();
            mbar_wait(&empty_bar[stage], phase ^ 1);
            pw += clock64() - t;
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
          }
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (p.prof) t_issue[stage] = clock64();
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          if (p.mode == IG_PLAIN) {
            tma_load_2d_2sm(sa, &maps.a, &full_bar[stage], kb * BLOCK_K, c.m0);
            // long-K GEMMs stream their A operand from HBM (fc2: 130 MB of hidden activations): pull the tile that the
            // ring will ask for PF k-blocks from now into L2, so the ring's own loads see L2 latency
            if (p.a_prefetch > 0) {
              const int pk = kb + p.a_prefetch;
              if (pk < kblocks) {
                tma_prefetch_2d(&maps.a, pk * BLOCK_K, c.m0);
              } else if (ct + num_clusters < num_ctiles) {  // first k-blocks of this CTA's next tile
                const TileCoord cn = decode_ctile(p, ct + num_clusters, n_tiles_n, BLOCK_N, pairs_per_group, rank);
                if (cn.m0 != c.m0) tma_prefetch_2d(&maps.a, (pk - kblocks) * BLOCK_K, cn.m0);
              }
            }
          } else {
            const int tap = kb / cblocks;
            const int c0 = (kb - tap * cblocks) * BLOCK_K;
            if (p.mode == IG_CONV3) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              tma_load_4d_2sm(sa, &maps.a, &full_bar[stage], c0, c.w0 + dx, c.h0 + dy, c.b);
            } else if (p.mode == IG_UP2) {
              const int dy = (tap >> 1) + py - 1, dx = (tap & 1) + px - 1;
              tma_load_4d_2sm(sa, &maps.a, &full_bar[stage], c0, c.w0 + dx, c.h0 + dy, c.b);
            } else {  // IG_PATCH
              const int ky = tap / p.patch, kx = tap % p.patch;
              tma_load_5d_2sm(sa, &maps.a, &full_bar[stage], c0, kx, c.w0, ky, c.b * p.Hq + c.h0);
              // the patch gather streams the whole last encoder map from HBM (262 MB at 64 x 4 s) through this ring:
              // pull the box the ring will ask for a_prefetch k-blocks from now into L2 (same reason as for fc2 above)
              if (p.a_prefetch > 0) {
                const int pk = kb + p.a_prefetch;
                if (pk < kblocks) {
                  const int t2 = pk / cblocks;
                  tma_prefetch_5d(&maps.a, (pk - t2 * cblocks) * BLOCK_K, t2 % p.patch, c.w0, t2 / p.patch, c.b * p.Hq + c.h0);
                } else if (ct + num_clusters < num_ctiles) {
                  const TileCoord cn = decode_ctile(p, ct + num_clusters, n_tiles_n, BLOCK_N, pairs_per_group, rank);
                  const int pk2 = pk - kblocks, t2 = pk2 / cblocks;
                  if (cn.b != c.b || cn.h0 != c.h0 || cn.w0 != c.w0)
                    tma_prefetch_5d(&maps.a, (pk2 - t2 * cblocks) * BLOCK_K, t2 % p.patch, cn.w0, t2 / p.patch, cn.b * p.Hq + cn.h0);
                }
              }
            }
          }
          tma_load_2d_2sm(sb, &maps.b, &full_bar[stage], kb * BLOCK_K, b_row);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (p.prof) {
        p.prof[blockIdx.x * 16 + 0] = pw;
        p.prof[blockIdx.x * 16 + 1] = clock64() - pt0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc_16(2 * BLOCK_M, BLOCK_N, 0, 0, p.f16);
      // This thread's own instruction stream competes with the epilogue warps for issue slots (fc1's GELU epilogue
      // keeps the schedulers 70 % busy), so the loop is kept lean: descriptors are (lo, hi) words, lo = a per-stage
      // base + a compile-time k offset; diagnostics flags are hoisted.
      constexpr uint32_t D_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version bit 46, 128B swizzle
      constexpr uint32_t D_LBO = (16u >> 4) << 16;
      const uint32_t a_lo_base = ((smem_u32(smem) & 0x3FFFFu) >> 4) | D_LBO;
      const bool prof = p.prof != nullptr;
      const bool skip_mma = (p.dbg & 4) != 0;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long mw_e = 0, mw_f = 0, lat_sum = 0, lat_n = 0;
      const long long mt0 = clock64();
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        if (prof) {
          const long long t = clock64();
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          mw_e += clock64() - t;
        } else {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (prof) {
            const long long t = clock64();
            mbar_wait(&full_bar[stage], phase);
            const long long t2 = clock64();
            mw_f += t2 - t;
            if (t2 - t > 100) {  // the load was still in flight: issue -> landed (an upper bound of the TMA round trip)
              lat_sum += t2 - t_issue[stage];
              ++lat_n;
            }
          } else {
            mbar_wait(&full_bar[stage], phase);
          }
          tc_fence_after();
          const uint32_t a_lo = a_lo_base + stage * (C::STAGE_BYTES >> 4);
          const uint32_t b_lo = a_lo + (A_STAGE_BYTES >> 4);
          if (!skip_mma) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_16_2sm_lh(d_tmem, a_lo + ((k * UMMA_K * 2) >> 4), D_HI, b_lo + ((k * UMMA_K * 2) >> 4), D_HI, idesc,
                             k != 0 ? 1u : (kb != 0 ? 1u : 0u));
          }
          umma_commit_2sm(&empty_bar[stage], 3);  // stage free in both CTAs once these MMAs have read it
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tmem_full[acc], 3);  // accumulator halves complete in both CTAs
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (p.prof) {
        p.prof[blockIdx.x * 16 + 2] = mw_e;
        p.prof[blockIdx.x * 16 + 3] = mw_f;
        p.prof[blockIdx.x * 16 + 4] = clock64() - mt0;
        p.prof[blockIdx.x * 16 + 13] = lat_sum;
        p.prof[blockIdx.x * 16 + 14] = lat_n;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> regs -> smem -> TMA store
    const int ew = warp - 2;
    const int sub = warp & 3;            // TMEM lane quarter this warp may read
    const int grp = ew >> 2;             // column-block parity owned by this warp's group
    const int m = sub * 32 + lane;       // accumulator row == TMEM lane
    const bool issuer = (ew & 3) == 0 && lane == 0;
    constexpr bool kF32 = EPI == EPI_F32 || EPI == EPI_F32D;
    const bool out_f32 = EPI == EPI_GENERIC ? p.out_f32 != 0 : kF32;
    const bool pool = EPI == EPI_GENERIC ? p.pool != 0 : EPI == EPI_BNPOOL16;
    const int wcols = out_f32 ? 32 : 64;            // columns per 128-byte block
    const int nblk = BLOCK_N / wcols;
    const int J = (nblk - grp + 1) / 2;             // blocks of this group per tile
    uint8_t* stg0 = staging + grp * C::NBUF * STG_BUF_BYTES;
    // In-place residual (x += f(x), the transformer's proj / fc2): nothing is loaded - the block is added into
    // global memory by a TMA reduce-add.  Otherwise the fp32 residual tile is TMA-loaded into the staging buffer.
    const bool red_add = (EPI == EPI_GENERIC || kF32) && p.residual != nullptr && p.res_inplace != 0;
    const bool has_res = (EPI == EPI_GENERIC || kF32) && p.residual != nullptr && !red_add;
    // staging row of this thread (pooled convolutions only keep the pooled pixels)
    const bool writer = pool ? ((lane & 17) == 0) : true;
    const int srow = pool ? (sub * 8 + ((lane & 15) >> 1)) : m;
    const int sw = srow & 7;
    const float relu_lo = p.act == ACT_RELU ? 0.0f : -INFINITY;

    // Per-channel constants of the CURRENT tile in shared memory: [acc parity][scale BLOCK_N | shift BLOCK_N].  Every
    // epilogue thread loads one channel of the tile before it waits for the accumulator; one named barrier per tile.
    float* cbuf = reinterpret_cast<float*>(consts);

    const float ln_inv_slots = p.ln_slots > 0 ? 1.0f / static_cast<float>(p.ln_slots) : 0.f;
    const float ln_per = p.ln_slots > 0 ? static_cast<float>(p.K / p.ln_slots) : 0.f;
    const float ln_inv_k = 1.0f / static_cast<float>(p.K);
    float2 ln_pre[4];          // folded LayerNorm: the next tile's statistics partials of this thread's row
    bool ln_have_pre = false;
    uint32_t it = 0;  // running column-block counter of this group: buffer = it & 1, residual phase = (it >> 1) & 1
    int acc = 0;
    uint32_t acc_phase = 0;
    long long ew_full = 0, e_cst = 0, e_blk = 0, e_tiles = 0;
    const long long et0 = clock64();
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
      const long long tA = clock64();
      const TileCoord c = decode_ctile(p, ct, n_tiles_n, BLOCK_N, pairs_per_group, rank);
      // output coordinates of the tile origin (dims 1..3 of the 4-D output map; dim 0 is the channel)
      int o1, o2, o3;
      tile_out_coords(p, c, pool, o1, o2, o3);
      const CUtensorMap* omap = &maps.out[c.par];
      const int r3 = p.res_mod > 0 ? 0 : o3;  // positional table: same rows for every clip
      float* cshift = cbuf + acc * C::CONST_BUF;
      float* cscale = C::HAS_SCALE ? cshift + C::CONST_N : cshift;  // (never read by the shift-only epilogues)
      for (int e = ew * 32 + lane; e < BLOCK_N; e += NUM_EPI_WARPS * 32) {
        if (C::HAS_SCALE) cscale[e] = p.scale != nullptr ? __ldg(p.scale + c.n0 + e) : 1.0f;
        cshift[e] = p.shift != nullptr ? __ldg(p.shift + c.n0 + e) : 0.0f;
      }
      constexpr int cbase = 0;
      // folded LayerNorm (consumer): this thread's row statistics from the producer's per-slot partials (equal counts:
      // the mean is the mean of the slot means, M2 adds the between-slot term)
      constexpr bool kLnIn = EPI == EPI_LNLIN16 || EPI == EPI_LNGELU16;
      float ln_rs = 0.f;
      if (kLnIn) {
        // (all partials in flight at once; and the NEXT tile's partials are loaded now into four registers, a tile ahead:
        // a load consumed right away costs a full L2 round trip under load - measured 1 700 cycles per tile in this
        // phase, which put the fc1 epilogue on the critical path; prefetch.global.L1 did not help)
        float2 t[8];
        if (ln_have_pre) {
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = i < 4 ? ln_pre[i] : make_float2(0.f, 0.f);
        } else {
          const float2* st = reinterpret_cast<const float2*>(p.ln_stats_in) +
                             static_cast<long long>(min(c.m0 + m, p.M - 1)) * p.ln_slots;
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = i < p.ln_slots ? __ldg(&st[i]) : make_float2(0.f, 0.f);
        }
        // short dependency chains and no IEEE division / square root: this runs once per tile on the epilogue's
        // critical path (the serial form - div, 8 dependent FMAs, div, sqrt, div - measured ~1 000 cycles per tile)
        const float mean = (((t[0].x + t[1].x) + (t[2].x + t[3].x)) + ((t[4].x + t[5].x) + (t[6].x + t[7].x))) * ln_inv_slots;
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = t[i].x - mean;
          e[i] = i < p.ln_slots ? fmaf(ln_per * d, d, t[i].y) : 0.f;
        }
        const float m2 = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
        ln_rs = c.m0 + m < p.M ? rsqrtf(fmaf(m2, ln_inv_k, p.ln_eps)) : 0.f;
        ln_have_pre = p.ln_slots <= 4 && ct + num_clusters < num_ctiles;
        if (ln_have_pre) {
          const TileCoord cn = decode_ctile(p, ct + num_clusters, n_tiles_n, BLOCK_N, pairs_per_group, rank);
          const float2* st = reinterpret_cast<const float2*>(p.ln_stats_in) +
                             static_cast<long long>(min(cn.m0 + m, p.M - 1)) * p.ln_slots;
#pragma unroll
          for (int i = 0; i < 4; ++i) ln_pre[i] = i < p.ln_slots ? __ldg(&st[i]) : make_float2(0.f, 0.f);
        }
      }
      const bool lnp = EPI == EPI_F32 && p.ln_stats_out != nullptr;  // LayerNorm producer (see emit32)
      const bool ln_row_ok = c.m0 + m < p.M;
      LnRun ln_run = {0.f, 0.f, 0.f};
      if (has_res && issuer && J > 0) {
        // residual tile of the first block -> staging buffer; the NEXT tile's residual blocks -> L2, so that the
        // per-block TMA loads of the next tile are L2 hits instead of exposed HBM latency
        uint8_t* buf = stg0 + (C::NBUF == 2 ? (it & 1) : 0) * STG_BUF_BYTES;
        mbar_expect_tx(&res_bar[grp * 2 + (it & 1)], STG_BUF_BYTES);
        tma_load_4d(buf, &maps.res, &res_bar[grp * 2 + (it & 1)], c.n0 + grp * wcols, o1, o2, r3);
        const int nct = ct + num_clusters;
        if (nct < num_ctiles) {
          const TileCoord cn = decode_ctile(p, nct, n_tiles_n, BLOCK_N, pairs_per_group, rank);
          int n1, n2, n3;
          tile_out_coords(p, cn, pool, n1, n2, n3);
          for (int j = 0; j < J; ++j)
            tma_prefetch_4d(&maps.res, cn.n0 + (2 * j + grp) * wcols, n1, n2, p.res_mod > 0 ? 0 : n3);
        }
      }

      const long long tB = clock64();
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      named_bar_sync(3, NUM_EPI_WARPS * 32);  // this tile's constants are in place (and the buffer of two tiles ago is free)
      const long long tC = clock64();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BLOCK_N;

      // hand a finished column block to the TMA: all 128 threads of the group have written (and fenced) their rows
      auto finish_block = [&](int j, uint8_t* buf, int col0) {
        fence_proxy_async_smem();            // staging writes -> visible to the TMA (async proxy)
        if (issuer) tma_store_wait_read0();  // the previous store of this group has released the other buffer
        named_bar_sync(1 + grp, 128);
        if (issuer) {
          if (red_add) tma_reduce_add_4d(omap, buf, col0, o1, o2, o3);
          else if (!(p.dbg & 1)) tma_store_4d(omap, buf, col0, o1, o2, o3);
          tma_store_commit();
        }
        ++it;
      };

      uint32_t ra[32], rb[32];
      if (J > 0) tmem_ld32(t_base + grp * wcols, ra);
      if (!out_f32) {
        // 16-bit output: a block is two 32-column TMEM chunks; the next chunk is always in flight
        for (int j = 0; j < J; ++j) {
          const int blk = 2 * j + grp;
          uint8_t* buf = stg0 + (C::NBUF == 2 ? (it & 1) : 0) * STG_BUF_BYTES;
          uint8_t* row = buf + srow * 128;
          const float* sc = cscale + cbase + blk * 64;
          const float* sh = cshift + cbase + blk * 64;
          if (C::NBUF == 1) {  // single staging buffer: the previous TMA store must have finished reading it
            if (issuer) tma_store_wait_read0();
            named_bar_sync(1 + grp, 128);
          }
          tmem_ld_wait(ra);
          tmem_ld32(t_base + blk * 64 + 32, rb);
          if (p.f16) emit16<EPI, true>(ra, sc, sh, p, relu_lo, row, 0, sw, writer, 16, ln_rs);
          else emit16<EPI, false>(ra, sc, sh, p, relu_lo, row, 0, sw, writer, 16, ln_rs);
          tmem_ld_wait(rb);
          if (j + 1 < J) tmem_ld32(t_base + (blk + 2) * 64, ra);
          if (p.f16) emit16<EPI, true>(rb, sc + 32, sh + 32, p, relu_lo, row, 1, sw, writer, 16, ln_rs);
          else emit16<EPI, false>(rb, sc + 32, sh + 32, p, relu_lo, row, 1, sw, writer, 16, ln_rs);
          finish_block(j, buf, c.n0 + blk * 64);
        }
      } else {
        // fp32 output: a block is one chunk; ra / rb alternate as current / prefetch buffer
        auto block32 = [&](int j, uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
          const int blk = 2 * j + grp;
          uint8_t* buf = stg0 + (C::NBUF == 2 ? (it & 1) : 0) * STG_BUF_BYTES;
          if (C::NBUF == 1) {  // single staging buffer: the previous TMA store must have finished reading it
            if (issuer) tma_store_wait_read0();
            named_bar_sync(1 + grp, 128);
          }
          if (has_res && issuer && j + 1 < J) {
            // residual tile of the NEXT block -> the other buffer, issued before this block is processed so that the load
            // overlaps it (the buffer's last reader is the TMA store of block j - 1, committed a moment ago)
            tma_store_wait_read0();
            const uint32_t nb = (it + 1) & 1;
            mbar_expect_tx(&res_bar[grp * 2 + nb], STG_BUF_BYTES);
            tma_load_4d(stg0 + nb * STG_BUF_BYTES, &maps.res, &res_bar[grp * 2 + nb], c.n0 + (blk + 2) * 32, o1, o2, r3);
          }
          if (has_res) mbar_wait(&res_bar[grp * 2 + (it & 1)], (it >> 1) & 1);
          tmem_ld_wait(cur);
          if (j + 1 < J) tmem_ld32(t_base + (blk + 2) * 32, nxt);
          if (EPI == EPI_F32 && lnp) {
            uint8_t* x16 = ln_row_ok ? reinterpret_cast<uint8_t*>(p.ln_x16_out) +
                                           (static_cast<long long>(c.m0 + m) * p.ld16 + c.n0 + blk * 32) * 2
                                     : nullptr;
            emit32<EPI, true>(cur, cscale + cbase + blk * 32, cshift + cbase + blk * 32, p, relu_lo, buf + srow * 128, sw,
                              has_res, &ln_run, x16);
          } else {
            emit32<EPI>(cur, cscale + cbase + blk * 32, cshift + cbase + blk * 32, p, relu_lo, buf + srow * 128, sw,
                        has_res);
          }
          finish_block(j, buf, c.n0 + blk * 32);
        };
        for (int j = 0; j < J; j += 2) {
          block32(j, ra, rb);
          if (j + 1 < J) block32(j + 1, rb, ra);
        }
        if (lnp && ln_row_ok)  // this group's 128 columns of the row: slot = 2 * n_tile + group
          reinterpret_cast<float2*>(p.ln_stats_out)[static_cast<long long>(c.m0 + m) * (p.N >> 7) + 2 * (c.n0 / BLOCK_N) + grp] =
              make_float2(ln_run.mean, ln_run.m2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);  // the leader CTA owns the accumulator-free barrier
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
      e_cst += tB - tA; ew_full += tC - tB; e_blk += clock64() - tC; ++e_tiles;
    }
    if (p.prof && ew == 0 && lane == 0) {
      p.prof[blockIdx.x * 16 + 5] = ew_full;
      p.prof[blockIdx.x * 16 + 6] = e_cst;
      p.prof[blockIdx.x * 16 + 7] = e_blk;
      p.prof[blockIdx.x * 16 + 8] = clock64() - et0;
      p.prof[blockIdx.x * 16 + 9] = e_tiles;
    }
    if (issuer) tma_store_wait_all();  // global writes complete before the kernel (and its smem) goes away
  }

  const long long k_t2 = clock64();
  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer can still touch this CTA's smem / barriers
  if (warp == 2) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
  if (p.prof && threadIdx.x == 64) {
    p.prof[blockIdx.x * 16 + 10] = k_t1 - k_t0;
    p.prof[blockIdx.x * 16 + 11] = k_t2 - k_t1;
    p.prof[blockIdx.x * 16 + 12] = clock64() - k_t2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3 convolution with the input tile + halo held in shared memory (IgemmParams::halo).
//
// The tap-shifted TMA boxes of igemm_tc2_kernel read every input pixel nine times out of L2; these reads are private
// to an SM (no L2 request merging), which pins encoder.1 at the L2 -> SM throughput cap (measured 36 B/clk/SM,
// 35 % tensor-pipe activity).  Here a CTA loads its 8 x 16 pixel tile plus a one-pixel halo ONCE per 64 input
// channels - one 5-D TMA box (8 ch, 10, 18, 8 chunks) into the un-swizzled K-major "chunk plane" layout
// [chunk][hh][ww][8 ch = 16 B] - and the nine taps are nine UMMA descriptors into that buffer: with 8-pixel-wide tiles
// every 8-row core matrix is one contiguous 128-byte run (row h + dy, columns w + dx ..), consecutive core matrices
// are a constant WW * 16 bytes apart (SBO) and the two 8-channel chunks of a K = 16 step a constant plane apart (LBO),
// for any (dy, dx).  Weights still stream through a ring of 128B-swizzled half tiles (shared by all CTAs: these L2
// reads merge).  Epilogue: 16-bit BN (+ReLU) (+2x2 pool) paths of the kernel above with the 8 x 16 tile's lane map.
// ---------------------------------------------------------------------------------------------------------------
// 16 epilogue warps: the conv epilogues (short K, pool) are latency-bound with two warps per scheduler and were
// slower than the MMA main loop (encoder.1: ~4 000 vs 2 300 cycles per tile); four warps per scheduler, one 32-column
// TMEM chunk per warp and block, hide those latencies.
constexpr int NUM_EPI_WARPS_H = 16;
constexpr int NUM_THREADS_H = 64 + NUM_EPI_WARPS_H * 32;

template <int BLOCK_N, bool UP2>
struct CfgH {
  static constexpr int WW = 10, HH = 18;                      // tile 8 x 16 + halo
  static constexpr int PLANE = HH * WW * 16;                  // bytes per 8-channel chunk plane
  static constexpr int HALO_BYTES = 8 * PLANE;                // 64 channels: 23040 B
  static constexpr int A_STAGES = BLOCK_N == 256 ? 2 : 3;
  static constexpr int B_HALF_ROWS = BLOCK_N / 2;
  static constexpr int B_STAGE_BYTES = B_HALF_ROWS * BLOCK_K * 2;
  static constexpr int B_STAGES = BLOCK_N == 256 ? 6 : (UP2 ? 8 : 9);  // 9 = all taps of a 64-channel conv (resident mode)
  // UP2 (nearest x2 + 3x3 as four parity 2x2 convs): the four parity outputs of a low-resolution tile share ONE halo
  // tile and accumulate side by side in TMEM (4 * BLOCK_N columns per stage; a single stage when BLOCK_N = 128)
  static constexpr int NPAR = UP2 ? 4 : 1;
  static constexpr int NTAP = UP2 ? 4 : 9;
  static constexpr int ACC_COLS = NPAR * BLOCK_N;
  static constexpr int ACC_STAGES = 2 * ACC_COLS <= 512 ? 2 : 1;
  static constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "accumulators must fit the 512 TMEM columns");
  static constexpr int STAGING_BYTES = 4 * STG_BUF_BYTES;
  static constexpr int CONST_N = 2048;
  static constexpr int CONST_BYTES = 2 * CONST_N * 4;
  static constexpr int OFF_B = 0;
  static constexpr int OFF_STG = OFF_B + B_STAGES * B_STAGE_BYTES;
  static constexpr int OFF_A = OFF_STG + STAGING_BYTES;
  static constexpr int OFF_CONST = OFF_A + A_STAGES * HALO_BYTES;
  static constexpr int OFF_BAR = OFF_CONST + CONST_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
};

// shared-memory matrix descriptor, no swizzle (layout_type 0), K-major: 8-row x 16-byte core matrices;
// lbo = bytes between core matrices adjacent in K, sbo = bytes between core matrices adjacent in M
__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t saddr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

template <int BLOCK_N, int EPI, bool UP2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS_H, 1)
igemm_halo_kernel(const __grid_constant__ IgemmMaps maps, const IgemmParams p, int num_ctiles, int n_tiles_n,
                  int pairs_per_group) {
  using C = CfgH<BLOCK_N, UP2>;
  // (declared aligned instead of rounding the pointer up by hand: integer arithmetic on the address loses the
  // shared-memory address space and turns every staging store / constant load of the epilogue into a generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B swizzle needs 1024-byte aligned tiles
  uint8_t* staging = smem + C::OFF_STG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* a_full = bars;                          // [A_STAGES] (leader)
  uint64_t* a_empty = a_full + C::A_STAGES;         // [A_STAGES]
  uint64_t* b_full = a_empty + C::A_STAGES;         // [B_STAGES] (leader)
  uint64_t* b_empty = b_full + C::B_STAGES;         // [B_STAGES]
  uint64_t* tmem_full = b_empty + C::B_STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2] (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int cblocks = p.Cin / BLOCK_K;
  // all weight tiles of the layer fit the ring and every tile of this CTA pair uses the same ones (64 input channels,
  // one n-tile: encoder.1): load them once and keep them - no per-tap barrier traffic, a quarter of the L2 reads
  const bool resident = C::NPAR * C::NTAP == C::B_STAGES && cblocks == 1 && n_tiles_n == 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a);
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.out[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::A_STAGES; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < C::B_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * NUM_EPI_WARPS_H);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      const bool prof = p.prof != nullptr;
      int as = 0, bs = 0;                  // ring slots (running counters: no modulo in the loop)
      uint32_t a_par = 1, b_par = 1;       // parity to wait for on the "empty" barriers
      bool first = true;
      long long pw_a = 0, pw_b = 0;
      const long long pt0 = clock64();
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        const TileCoord c = decode_ctile(p, ct, n_tiles_n, BLOCK_N, pairs_per_group, rank);
        const int b_row = c.n0 + rank * C::B_HALF_ROWS;
        for (int cb = 0; cb < cblocks; ++cb) {
          if (prof) {
            const long long t = clock64();
            mbar_wait(&a_empty[as], a_par);
            pw_a += clock64() - t;
          } else {
            mbar_wait(&a_empty[as], a_par);
          }
          if (rank == 0) mbar_expect_tx(&a_full[as], 2 * C::HALO_BYTES);
          tma_load_5d_2sm(smem + C::OFF_A + as * C::HALO_BYTES, &maps.a, &a_full[as], 0, c.w0 - 1, c.h0 - 1, cb * 8, c.b);
          if (++as == C::A_STAGES) {
            as = 0;
            a_par ^= 1;
          }
          if (resident && !first) continue;  // the weight tiles of the first tile stay in their ring slots
#pragma unroll
          for (int pt = 0; pt < C::NPAR * C::NTAP; ++pt) {
            const int par = pt / C::NTAP, tap = pt % C::NTAP;
            if (prof) {
              const long long t = clock64();
              mbar_wait(&b_empty[bs], b_par);
              pw_b += clock64() - t;
            } else {
              mbar_wait(&b_empty[bs], b_par);
            }
            if (rank == 0) mbar_expect_tx(&b_full[bs], 2 * C::B_STAGE_BYTES);
            tma_load_2d_2sm(smem + C::OFF_B + bs * C::B_STAGE_BYTES, &maps.b, &b_full[bs], (tap * cblocks + cb) * BLOCK_K,
                            par * p.N + b_row);
            if (++bs == C::B_STAGES) {
              bs = 0;
              b_par ^= 1;
            }
          }
        }
        first = false;
      }
      if (prof) {
        p.prof[blockIdx.x * 16 + 0] = pw_a;
        p.prof[blockIdx.x * 16 + 1] = pw_b;
        p.prof[blockIdx.x * 16 + 2] = clock64() - pt0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    // At BLOCK_N <= 128 the four MMAs of a tap take 128-256 cycles, so this thread's own instruction stream is on the
    // critical path (measured: 420 cycles per tap with a rolled loop that rebuilt both descriptors and divided by
    // NTAP): the tap loop is fully unrolled, descriptors are a 32-bit add of a compile-time offset to a per-slot
    // base word, ring slots are running counters.
    if (rank == 0 && elect_one()) {
      const bool prof = p.prof != nullptr;
      const uint32_t idesc = make_idesc_16(2 * BLOCK_M, BLOCK_N, 0, 0, p.f16);
      // descriptor words: lo = start address >> 4 | LBO >> 4 << 16, hi = SBO >> 4 | 1 << 14 (| 128B swizzle << 29)
      constexpr uint32_t A_HI = static_cast<uint32_t>((C::WW * 16) >> 4) | (1u << 14);
      constexpr uint32_t A_LBO = static_cast<uint32_t>(C::PLANE >> 4) << 16;
      constexpr uint32_t B_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t B_LBO = (16u >> 4) << 16;
      const uint32_t a_lo_base = ((smem_u32(smem + C::OFF_A) & 0x3FFFFu) >> 4) | A_LBO;
      const uint32_t b_lo_base = ((smem_u32(smem + C::OFF_B) & 0x3FFFFu) >> 4) | B_LBO;
      int as = 0, bs = 0;
      uint32_t a_par = 0, b_par = 0;
      bool first = true;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long mw_e = 0, mw_a = 0, mw_b = 0, m_tiles = 0;
      const long long mt0 = clock64();
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        long long t = prof ? clock64() : 0;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        if (prof) mw_e += clock64() - t;
        ++m_tiles;
        tc_fence_after();
        const uint32_t d_acc = tmem_base + acc * C::ACC_COLS;
        for (int cb = 0; cb < cblocks; ++cb) {
          t = prof ? clock64() : 0;
          mbar_wait(&a_full[as], a_par);
          if (prof) mw_a += clock64() - t;
          tc_fence_after();
          const uint32_t a_lo = a_lo_base + as * (C::HALO_BYTES >> 4);
#pragma unroll
          for (int pt = 0; pt < C::NPAR * C::NTAP; ++pt) {
            const int par = pt / C::NTAP, tap = pt % C::NTAP;
            if (!resident || first) {
              if (prof) {
                t = clock64();
                mbar_wait(&b_full[bs], b_par);
                mw_b += clock64() - t;
              } else {
                mbar_wait(&b_full[bs], b_par);
              }
              tc_fence_after();
            }
            const uint32_t b_lo = b_lo_base + bs * (C::B_STAGE_BYTES >> 4);
            // tap position in halo coordinates (tile pixel (h, w) sits at (h + 1, w + 1)):
            // conv: (h + ky, w + kx); parity (py, px) of the x2 upsample, tap (a, b): (h + a + py, w + b + px)
            const int dy = UP2 ? (tap >> 1) + (par >> 1) : tap / 3, dx = UP2 ? (tap & 1) + (par & 1) : tap % 3;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_16_2sm_lh(d_acc + par * BLOCK_N, a_lo + (((dy * C::WW + dx) * 16 + k * 2 * C::PLANE) >> 4), A_HI,
                             b_lo + ((k * UMMA_K * 2) >> 4), B_HI, idesc, (tap | k) != 0 ? 1u : (cb != 0 ? 1u : 0u));
            if (!resident) umma_commit_2sm(&b_empty[bs], 3);
            if (++bs == C::B_STAGES) {
              bs = 0;
              b_par ^= 1;
            }
          }
          umma_commit_2sm(&a_empty[as], 3);
          if (++as == C::A_STAGES) {
            as = 0;
            a_par ^= 1;
          }
        }
        first = false;
        umma_commit_2sm(&tmem_full[acc], 3);
        if (++acc == C::ACC_STAGES) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (prof) {
        p.prof[blockIdx.x * 16 + 3] = mw_e;
        p.prof[blockIdx.x * 16 + 4] = mw_a;
        p.prof[blockIdx.x * 16 + 5] = mw_b;
        p.prof[blockIdx.x * 16 + 6] = clock64() - mt0;
        p.prof[blockIdx.x * 16 + 9] = m_tiles;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (16-bit output, BN / ReLU / pool)
    const int ew = warp - 2;
    const int sub = warp & 3;            // TMEM lane quarter this warp may read
    const int grp = ew >> 3;             // 8 warps per group; group g owns the 64-column blocks g, g + 2, ...
    const int half = (ew >> 2) & 1;      // which 32-column half of a block this warp converts
    const int m = sub * 32 + lane;       // accumulator row: pixel (h = m >> 3, w = m & 7) of the 8 x 16 tile
    const bool issuer = (ew & 7) == 0 && lane == 0;
    constexpr bool pool = EPI == EPI_BNPOOL16 || EPI == EPI_SHPOOL16;
    constexpr int nblk = BLOCK_N / 64;
    const int J = (nblk - grp + 1) / 2;
    uint8_t* stg0 = staging + grp * 2 * STG_BUF_BYTES;
    // pooled tile: 4 x 8 pixels, row = (h / 2) * 4 + w / 2; the lanes with even h and even w write
    const bool writer = pool ? ((lane & 9) == 0) : true;
    const int srow = pool ? ((sub * 2 + (lane >> 4)) * 4 + ((lane & 7) >> 1)) : m;
    const int sw = srow & 7;
    const float relu_lo = p.act == ACT_RELU ? 0.0f : -INFINITY;
    float* cscale = reinterpret_cast<float*>(smem + C::OFF_CONST);
    float* cshift = cscale + C::CONST_N;
    for (int e = ew * 32 + lane; e < p.N; e += NUM_EPI_WARPS_H * 32) {
      cscale[e] = p.scale != nullptr ? __ldg(p.scale + e) : 1.0f;
      cshift[e] = p.shift != nullptr ? __ldg(p.shift + e) : 0.0f;
    }
    named_bar_sync(3, NUM_EPI_WARPS_H * 32);

    uint32_t it = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long ew_full = 0, e_ld = 0, e_emit = 0, e_st = 0;
    const long long et0 = clock64();
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
      const TileCoord c = decode_ctile(p, ct, n_tiles_n, BLOCK_N, pairs_per_group, rank);
      int o1, o2, o3;
      tile_out_coords(p, c, pool, o1, o2, o3);
      const long long t = p.prof ? clock64() : 0;
      mbar_wait(&tmem_full[acc], acc_phase);
      if (p.prof) ew_full += clock64() - t;
      tc_fence_after();
      for (int par = 0; par < C::NPAR; ++par) {
        const uint32_t t_base =
            tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * C::ACC_COLS + par * BLOCK_N + half * 32;
        for (int j = 0; j < J; ++j) {
          const int blk = 2 * j + grp;
          uint8_t* buf = stg0 + (it & 1) * STG_BUF_BYTES;
          uint8_t* row = buf + srow * 128;
          const float* sc = cscale + c.n0 + blk * 64 + half * 32;
          const float* sh = cshift + c.n0 + blk * 64 + half * 32;
          uint32_t ra[32];
          const long long q0 = p.prof ? clock64() : 0;
          tmem_ld32(t_base + blk * 64, ra);
          tmem_ld_wait(ra);
          const long long q1 = p.prof ? clock64() : 0;
          if (p.f16) emit16<EPI, true>(ra, sc, sh, p, relu_lo, row, half, sw, writer, 8);
          else emit16<EPI, false>(ra, sc, sh, p, relu_lo, row, half, sw, writer, 8);
          const long long q2 = p.prof ? clock64() : 0;
          fence_proxy_async_smem();
          if (issuer) tma_store_wait_read0();
          named_bar_sync(1 + grp, 256);
          if (issuer) {
            tma_store_4d(&maps.out[par], buf, c.n0 + blk * 64, o1, o2, o3);
            tma_store_commit();
          }
          if (p.prof) {
            const long long q3 = clock64();
            e_ld += q1 - q0; e_emit += q2 - q1; e_st += q3 - q2;
          }
          ++it;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
      if (++acc == C::ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (issuer) tma_store_wait_all();
    if (p.prof && ew == 0 && lane == 0) {
      p.prof[blockIdx.x * 16 + 7] = ew_full;
      p.prof[blockIdx.x * 16 + 8] = clock64() - et0;
      p.prof[blockIdx.x * 16 + 10] = e_ld;
      p.prof[blockIdx.x * 16 + 11] = e_emit;
      p.prof[blockIdx.x * 16 + 12] = e_st;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
}

template <int BLOCK_N, int EPI, bool UP2>
int launch_halo_impl(const IgemmParams& p, const IgemmMaps& maps, int num_ctiles, int n_tiles_n, int pairs_per_group,
                     int num_sms, cudaStream_t stream) {
  using C = CfgH<BLOCK_N, UP2>;
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_halo_kernel<BLOCK_N, EPI, UP2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) {
      once.retry();
      set_error("igemm_halo: cudaFuncSetAttribute(%d B smem) failed: %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return -4;
    }
  }
  const int max_clusters = num_sms / 2;
  const int clusters = num_ctiles < max_clusters ? num_ctiles : max_clusters;
  cudaError_t le = launch_pdl(igemm_halo_kernel<BLOCK_N, EPI, UP2>, dim3(2 * clusters), dim3(NUM_THREADS_H), C::SMEM_BYTES,
                              stream, maps, p, num_ctiles, n_tiles_n, pairs_per_group);
  if (le == cudaSuccess) le = cudaGetLastError();
  if (le != cudaSuccess) {
    set_error("igemm_halo<%d,%d>: %s (dyn smem %d)", BLOCK_N, EPI, cudaGetErrorString(le), C::SMEM_BYTES);
    return -4;
  }
  return 0;
}

template <int BLOCK_N>
int launch_halo_n(const IgemmParams& p, const IgemmMaps& maps, int a, int n_tiles_n, int b, int num_sms,
                  cudaStream_t stream) {
  const bool sh = p.scale == nullptr;  // BN scale folded into the weights: the epilogue only adds the shift
  if (p.pool)
    return sh ? launch_halo_impl<BLOCK_N, EPI_SHPOOL16, false>(p, maps, a, n_tiles_n, b, num_sms, stream)
              : launch_halo_impl<BLOCK_N, EPI_BNPOOL16, false>(p, maps, a, n_tiles_n, b, num_sms, stream);
  if (p.mode == IG_UP2) {
    if constexpr (BLOCK_N <= 128)
      return sh ? launch_halo_impl<BLOCK_N, EPI_SH16, true>(p, maps, a, n_tiles_n, b, num_sms, stream)
                : launch_halo_impl<BLOCK_N, EPI_BN16, true>(p, maps, a, n_tiles_n, b, num_sms, stream);
    set_error("igemm_halo: x2-upsample convs need block_n <= 128");
    return -1;
  }
  return sh ? launch_halo_impl<BLOCK_N, EPI_SH16, false>(p, maps, a, n_tiles_n, b, num_sms, stream)
            : launch_halo_impl<BLOCK_N, EPI_BN16, false>(p, maps, a, n_tiles_n, b, num_sms, stream);
}

template <int BLOCK_N, int EPI>
int launch_impl2(const IgemmParams& p, const IgemmMaps& maps, int num_ctiles, int n_tiles_n, int pairs_per_group,
                 int num_sms, cudaStream_t stream) {
  using C = Cfg2<BLOCK_N, EPI>;
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_tc2_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) {
      once.retry();
      set_error("igemm_tc2: cudaFuncSetAttribute(%d B smem) failed: %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return -4;
    }
  }
  const int max_clusters = num_sms / 2;
  const int clusters = num_ctiles < max_clusters ? num_ctiles : max_clusters;
  cudaError_t le = launch_pdl(igemm_tc2_kernel<BLOCK_N, EPI>, dim3(2 * clusters), dim3(NUM_THREADS), C::SMEM_BYTES,
                              stream, maps, p, num_ctiles, n_tiles_n, pairs_per_group);
  if (le == cudaSuccess) le = cudaGetLastError();
  if (le != cudaSuccess) {
    cudaFuncAttributes fa;
    memset(&fa, 0, sizeof(fa));
    cudaFuncGetAttributes(&fa, igemm_tc2_kernel<BLOCK_N, EPI>);
    set_error("igemm_tc2<%d,%d>: %s (grid %d, threads %d, dyn smem %d, regs %d, maxThreadsPerBlock %d)", BLOCK_N, EPI,
              cudaGetErrorString(le), 2 * clusters, NUM_THREADS, C::SMEM_BYTES, fa.numRegs, fa.maxThreadsPerBlock);
    return -4;
  }
  return 0;
}

// which compiled epilogue serves this parameter combination (see the EPI_* enum)
int pick_epi(const IgemmParams& p) {
  if (p.dbg & 16) return EPI_GENERIC;
  if (p.out_f32) {
    if (!(p.act == ACT_NONE && p.scale == nullptr && !p.pool)) return EPI_GENERIC;
    return (p.K >= 1024 && p.mode == IG_PLAIN && (p.residual == nullptr || p.res_inplace)) ? EPI_F32D : EPI_F32;
  }
  if (p.residual != nullptr) return EPI_GENERIC;
  if (p.ln_stats_in != nullptr) return p.act == ACT_GELU ? EPI_LNGELU16 : EPI_LNLIN16;
  if (p.pool) return (p.act != ACT_GELU) ? EPI_BNPOOL16 : EPI_GENERIC;
  if (p.scale != nullptr) return (p.act != ACT_GELU) ? EPI_BN16 : EPI_GENERIC;
  if (p.act == ACT_GELU) return EPI_GELU16;
  if (p.act == ACT_NONE) return EPI_LIN16;
  return EPI_BN16;  // ReLU without scale: scale constants default to 1
}

template <int BLOCK_N>
int launch_n(const IgemmParams& p, const IgemmMaps& maps, int a, int n_tiles_n, int b, int num_sms,
             cudaStream_t stream) {
  switch (pick_epi(p)) {
    case EPI_LIN16: return launch_impl2<BLOCK_N, EPI_LIN16>(p, maps, a, n_tiles_n, b, num_sms, stream);
    case EPI_GELU16: return launch_impl2<BLOCK_N, EPI_GELU16>(p, maps, a, n_tiles_n, b, num_sms, stream);
    case EPI_BN16: return launch_impl2<BLOCK_N, EPI_BN16>(p, maps, a, n_tiles_n, b, num_sms, stream);
    case EPI_BNPOOL16: return launch_impl2<BLOCK_N, EPI_BNPOOL16>(p, maps, a, n_tiles_n, b, num_sms, stream);
    case EPI_F32: return launch_impl2<BLOCK_N, EPI_F32>(p, maps, a, n_tiles_n, b, num_sms, stream);
    case EPI_F32D: return launch_impl2<BLOCK_N, EPI_F32D>(p, maps, a, n_tiles_n, b, num_sms, stream);
    case EPI_LNLIN16: return launch_impl2<BLOCK_N, EPI_LNLIN16>(p, maps, a, n_tiles_n, b, num_sms, stream);
    case EPI_LNGELU16: return launch_impl2<BLOCK_N, EPI_LNGELU16>(p, maps, a, n_tiles_n, b, num_sms, stream);
    default: return launch_impl2<BLOCK_N, EPI_GENERIC>(p, maps, a, n_tiles_n, b, num_sms, stream);
  }
}

}  // namespace

int launch_igemm_tc2(const IgemmParams& p, const IgemmMaps& maps, int block_n, int num_sms, cudaStream_t stream) {
  if (p.K % BLOCK_K != 0 || p.N % block_n != 0 || (p.mode != IG_PLAIN && p.Cin % BLOCK_K != 0)) {
    set_error("igemm_tc2: unsupported shape N=%d K=%d Cin=%d block_n=%d", p.N, p.K, p.Cin, block_n);
    return -1;
  }
  if (p.halo && !((p.mode == IG_CONV3 || (p.mode == IG_UP2 && block_n <= 128 && !p.pool)) && p.Wt == 8 && p.Ht == 16 &&
                  !p.out_f32 && p.residual == nullptr && p.act != ACT_GELU && 4 * p.N <= 2048)) {
    set_error("igemm_tc2: the halo convolution path needs a 16-bit 3x3 conv with an 8x16 tile");
    return -1;
  }
  if (p.pool && !p.halo && !(p.mode == IG_CONV3 && p.Wt == 16 && p.Ht == 8)) {
    set_error("igemm_tc2: fused pool needs a 16x8 spatial tile");
    return -1;
  }
  if (p.residual != nullptr && !p.out_f32) {
    set_error("igemm_tc2: a residual input needs an fp32 output");
    return -1;
  }
  IgemmParams pp = p;
  // L2 prefetch of an HBM-streamed A operand a few k-blocks ahead of the ring (IG_PLAIN and IG_PATCH producers): OFF by
  // default.  Round 1 measured a gain for fc2 (69.3 -> 65.3 us at distance 4); with the residual stream pinned in L2 and
  // the current ring depths it now costs time - round 2, same box, back to back: fc2 64.6 us without vs 68.6-70.7 us with
  // (distance 2 / 4 / 8 / 12), patch embedding 143.6 vs 170-181 us (profiles/r2_a_prefetch.json): the prefetched boxes
  // compete with the ring's own loads for the same L2 request slots.  HVIT_A_PREFETCH=k re-enables it for experiments.
  pp.a_prefetch = 0;
  if (const char* e = getenv("HVIT_A_PREFETCH")) pp.a_prefetch = atoi(e);
  pp.res_inplace = (p.residual != nullptr && p.residual == p.out && p.ldr == p.ldc && p.res_mod == 0 &&
                    !(p.dbg & 32) && p.ln_stats_out == nullptr) ? 1 : 0;
  if (p.ln_stats_out != nullptr &&
      !(p.mode == IG_PLAIN && p.out_f32 && p.residual != nullptr && p.act == ACT_NONE && p.scale == nullptr && !p.pool &&
        block_n == 256 && p.ln_x16_out != nullptr && p.ld16 % 16 == 0)) {
    set_error("igemm_tc2: the LayerNorm-producer epilogue needs a plain fp32 GEMM with a residual and 256-column tiles");
    return -1;
  }
  if (p.ln_stats_in != nullptr &&
      !(p.mode == IG_PLAIN && !p.out_f32 && p.residual == nullptr && p.scale == nullptr && !p.pool &&
        p.shift != nullptr && p.ln_slots > 0 && p.ln_slots <= 8 && p.K % p.ln_slots == 0 && (p.act == ACT_NONE || p.act == ACT_GELU))) {
    set_error("igemm_tc2: the folded-LayerNorm epilogue needs a plain 16-bit GEMM with the c vector as shift and row statistics");
    return -1;
  }
  const int n_tiles_n = p.N / block_n;
  long long group_tiles;
  int groups = 1;
  if (p.mode == IG_PLAIN) {
    group_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  } else {
    group_tiles = static_cast<long long>(p.B) * p.tiles_h * p.tiles_w;
    if (p.mode == IG_UP2 && !p.halo) groups = 4;  // (the halo kernel handles the four parities inside a tile)
  }
  const long long pairs_per_group = (group_tiles + 1) / 2;
  const long long nct = pairs_per_group * groups * n_tiles_n;
  if (nct <= 0 || nct > 0x7FFFFFFF) {
    set_error("igemm_tc2: bad tile count %lld", nct);
    return -1;
  }
  const int a = static_cast<int>(nct), b = static_cast<int>(pairs_per_group);
  pp.fd_ntn = make_fastdiv(n_tiles_n);
  pp.fd_ppg = make_fastdiv(b);
  pp.fd_tw = make_fastdiv(p.mode == IG_PLAIN ? 1 : p.tiles_w);
  pp.fd_th = make_fastdiv(p.mode == IG_PLAIN ? 1 : p.tiles_h);
  if (getenv("HVIT_PROF") != nullptr && pp.prof == nullptr) {
    // diagnostics: per-role cycle counters of this launch (mean over the leader CTAs), printed to stderr
    long long* d = nullptr;
    cudaMalloc(&d, sizeof(long long) * 16 * num_sms);
    cudaMemset(d, 0, sizeof(long long) * 16 * num_sms);
    IgemmParams q = pp;
    q.prof = d;
    const int r = launch_igemm_tc2(q, maps, block_n, num_sms, stream);
    cudaDeviceSynchronize();
    std::vector<long long> h(16 * num_sms);
    cudaMemcpy(h.data(), d, sizeof(long long) * 16 * num_sms, cudaMemcpyDeviceToHost);
    cudaFree(d);
    double s[16] = {0};
    int n = 0;
    for (int c = 0; c < num_sms; c += 2) {
      if (h[c * 16 + 9] == 0) continue;
      for (int k = 0; k < 16; ++k) s[k] += static_cast<double>(h[c * 16 + k]);
      ++n;
    }
    for (int k = 0; k < 16; ++k) s[k] /= (n > 0 ? n : 1);
    if (pp.halo) {
      fprintf(stderr,
              "[halo prof mode=%d M=%d N=%d K=%d bn=%d epi=%d] tiles/cta %.1f | producer wait_a_empty %.0f wait_b_empty %.0f "
              "total %.0f | mma wait_tmem_empty %.0f wait_a_full %.0f wait_b_full %.0f total %.0f | epi wait_full %.0f ld %.0f "
              "emit %.0f fence+bar+store %.0f total %.0f (cycles, leader CTAs)\n",
              p.mode, p.M, p.N, p.K, block_n, pick_epi(pp), s[9], s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[10],
              s[11], s[12], s[8]);
      return r;
    }
    fprintf(stderr,
            "[igemm prof mode=%d M=%d N=%d K=%d bn=%d epi=%d] tiles/cta %.1f | producer wait_empty %.0f total %.0f | mma "
            "wait_tmem_empty %.0f wait_full %.0f total %.0f | epi cst %.0f wait_full %.0f blocks %.0f total %.0f | kernel: prologue "
            "%.0f roles %.0f exit %.0f | TMA issue -> landed when the MMA thread was waiting: %.0f cycles (%.0f samples/CTA)\n",
            p.mode, p.M, p.N, p.K, block_n, pick_epi(pp), s[9], s[0], s[1], s[2], s[3], s[4], s[6], s[5], s[7], s[8], s[10],
            s[11], s[12], s[14] > 0 ? s[13] / s[14] : 0.0, s[14]);
    return r;
  }
  if (pp.halo) {
    switch (block_n) {
      case 256: return launch_halo_n<256>(pp, maps, a, n_tiles_n, b, num_sms, stream);
      case 128: return launch_halo_n<128>(pp, maps, a, n_tiles_n, b, num_sms, stream);
      case 64: return launch_halo_n<64>(pp, maps, a, n_tiles_n, b, num_sms, stream);
      default: set_error("igemm_tc2: block_n must be 64/128/256"); return -1;
    }
  }
  switch (block_n) {
    case 256: return launch_n<256>(pp, maps, a, n_tiles_n, b, num_sms, stream);
    case 128: return launch_n<128>(pp, maps, a, n_tiles_n, b, num_sms, stream);
    case 64: return launch_n<64>(pp, maps, a, n_tiles_n, b, num_sms, stream);
    default: set_error("igemm_tc2: block_n must be 64/128/256"); return -1;
  }
}

}  // namespace hvit
