// Encoder block 0 ("stem", components.py:15-99 as built at hybrid_vit.py:196-209) on the tensor cores:
//   Conv3x3(1 -> 64, pad 1, no bias) -> BatchNorm(eval) -> ReLU -> MaxPool2 (floor), fp32 spectrogram in, NHWC 16-bit out.
//
// The CUDA-core version of this block is FFMA-bound (4.75 G FMA per 64 x 4 s batch: 343 us measured against a 50 us
// HBM floor).  Here the block is restated as four small GEMMs, one per pre-pool position pos = (py, px) in {0,1}^2:
//
//   D_pos[m, c] = sum_k Awin[m, k] * B_pos[c, k]
//
//   Awin[m, :]  one row per POOLED pixel m of the tile: the 4x4 input window x[2ph-1 .. 2ph+2][2pw-1 .. 2pw+2] (zero
//               padded, already divided by mag_max), k = part*16 + (wy*4 + wx).  fp16 mode: the window rounded to fp16
//               (K = 16: one rounding of the normalised magnitude, the same error class as every other activation of that
//               mode); bf16 mode: part 0 = bf16 "hi", part 1 = the bf16 residual "lo" (K = 32, x = hi + lo to ~2^-16)
//   B_pos[c, k] = bn_scale[c] * w[c, ky, kx] at wy = py + ky, wx = px + kx (both parts), else 0
//
// TMEM lane m then holds, in the four accumulators, the four pre-pool conv outputs of all 64 channels of pooled pixel m:
// pooling is a per-thread max over the four accumulators, BN shift and ReLU are max(., -t) + t (the BN scale is folded
// into the 16-bit weights, like every other conv weight of the 16-bit modes), and a thread's 32 channels are 64
// contiguous bytes of the NHWC output.
//
// Persistent CTA per SM, 16 warps: warps 0-7 epilogue (TMEM lane quarter x channel half: TMEM -> max/shift/ReLU ->
// 16-bit -> 128B-swizzled staging tile of a pooled row -> TMA store), warps 8-11 build the window rows in the
// 128B-swizzled K-major layout from an input patch that warps 13-15 stage in shared memory with async copies, warp 12
// issues the MMAs.  Tile = 128 pooled pixels (two pooled rows x 64 pooled columns); 4-stage window ring,
// double-buffered accumulators (2 x 4 x 64 TMEM columns), double-buffered staging per pooled row.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace hvit {
namespace {

constexpr int ST_THREADS = 512;            // warps 0-7 epilogue, 8-11 window builders, 12 MMA issuer, 13-15 patch producers
constexpr int N_PRODUCERS = 3;
constexpr int W_BYTES = 4 * 64 * 128;        // four position matrices B_pos, 64 channel rows x 128 B
constexpr int A_TILE_BYTES = 128 * 128;      // 128 window rows x 128 B (32 or 64 bytes of each row are used)
constexpr int A_STAGES = 4;
constexpr int O_TILE_BYTES = 64 * 128;       // one pooled row of the tile: 64 pooled pixels x 64 channels x 2 B
constexpr int OFF_W = 0;
constexpr int OFF_A = W_BYTES;
constexpr int OFF_O = OFF_A + A_STAGES * A_TILE_BYTES;   // [set][buffer]
constexpr int PATCH_W = 132;                 // input columns per patch row (130 used), 528 B
constexpr int PATCH_BYTES = 2 * 4 * PATCH_W * 4 + 128;  // [set][4 rows][PATCH_W] fp32 + a 128-byte slot for 1/mag_max
constexpr int P_STAGES = 4;
constexpr int OFF_P = OFF_O + 4 * O_TILE_BYTES;
constexpr int OFF_SHIFT = OFF_P + P_STAGES * PATCH_BYTES;
constexpr int OFF_BAR = OFF_SHIFT + 256;
constexpr int ST_SMEM = OFF_BAR + 256;

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&a)[8], uint32_t (&b)[8], uint32_t (&c)[8], uint32_t (&d)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) asm volatile("" : "+r"(a[i]), "+r"(b[i]), "+r"(c[i]), "+r"(d[i])::"memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&a)[16], uint32_t (&b)[16], uint32_t (&c)[16], uint32_t (&d)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  // tie the destination registers to the wait so no use is scheduled above it
#pragma unroll
  for (int i = 0; i < 16; ++i)
    asm volatile("" : "+r"(a[i]), "+r"(b[i]), "+r"(c[i]), "+r"(d[i])::"memory");
}

template <bool F16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}

// x0, x1 -> packed 16-bit hi parts and packed 16-bit residuals (x - hi)
template <bool F16>
__device__ __forceinline__ void split_hi_lo(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack2<F16>(x0, x1);
  float h0, h1;
  if (F16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    h0 = f.x; h1 = f.y;
  } else {
    h0 = __uint_as_float(hi << 16);
    h1 = __uint_as_float(hi & 0xFFFF0000u);
  }
  lo = pack2<F16>(x0 - h0, x1 - h1);
}

// B_pos matrices [4][64][64] (16-bit, plain row-major, k < 16 * parts used) from the fp32 stem weights [3][3][64] and
// the BN scale
__global__ void stem_pack_kernel(const float* __restrict__ w9c, const float* __restrict__ scale, uint16_t* __restrict__ apack,
                                 int f16) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over 4 * 64 * 64
  if (idx >= 4 * 64 * 64) return;
  const int k = idx & 63, c = (idx >> 6) & 63, pos = idx >> 12;
  const int parts = f16 ? 1 : 2;  // fp16: 16-bit input window only; bf16: hi + lo residual
  const int widx = k & 15;
  const int wy = widx >> 2, wx = widx & 3;
  const int py = pos >> 1, px = pos & 1;
  const int ky = wy - py, kx = wx - px;
  float v = 0.f;
  if (k < 16 * parts && ky >= 0 && ky < 3 && kx >= 0 && kx < 3) v = w9c[(ky * 3 + kx) * 64 + c] * scale[c];
  const uint32_t pk = pack_16x2(v, 0.f, f16);
  apack[idx] = static_cast<uint16_t>(pk & 0xFFFFu);
}

template <bool F16>
__global__ void __launch_bounds__(ST_THREADS, 1)
stem_tc_kernel(const float* __restrict__ x, const unsigned* __restrict__ mag_max_bits,
               const uint16_t* __restrict__ apack, const float* __restrict__ shift,
               const __grid_constant__ CUtensorMap tmap_out, int H, int W, int Ho, int Wo, int tiles_w, int row_pairs,
               int num_tiles, long long* prof) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* a_full = bars;                  // [A_STAGES] window rows written (128 builder threads)
  uint64_t* a_empty = bars + A_STAGES;      // [A_STAGES] MMAs that read the stage have completed
  uint64_t* acc_full = a_empty + A_STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2] 256 epilogue threads
  uint64_t* p_full = acc_empty + 2;         // [P_STAGES] input patch landed (32 async-copy arrivals + 1)
  uint64_t* p_empty = p_full + P_STAGES;    // [P_STAGES] 128 builder threads have their windows in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + P_STAGES);
  float* shift_s = reinterpret_cast<float*>(smem + OFF_SHIFT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KSTEPS = F16 ? 1 : 2;  // K = 16 (fp16 window) or 32 (bf16 hi | lo)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&p_full[s], N_PRODUCERS * 32 + 1);
      mbar_init(&p_empty[s], 128);
    }
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(&a_full[s], 128);
      mbar_init(&a_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 256);
    }
    mbar_fence_init();
  }
  if (warp == 12) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  griddep_launch_dependents();
  griddep_wait();  // PDL: everything above overlapped the previous kernel's tail
  // weights -> shared memory in the 128B-swizzled K-major layout (row pitch 128 B, 16-byte chunk c of row r at c ^ (r & 7))
  for (int i = threadIdx.x; i < 4 * 64 * 8; i += ST_THREADS) {
    const int chunk = i & 7, row = (i >> 3) & 63, pos = i >> 9;
    const uint4 v = *reinterpret_cast<const uint4*>(apack + (pos * 64 + row) * 64 + chunk * 8);
    *reinterpret_cast<uint4*>(smem + OFF_W + pos * 8192 + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
  }
  if (threadIdx.x < 64) shift_s[threadIdx.x] = __ldg(shift + threadIdx.x);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int& b, int& rp, int& wt) {
    wt = tile % tiles_w;
    const int t2 = tile / tiles_w;
    rp = t2 % row_pairs;
    b = t2 / row_pairs;
  };

  if (warp == 12) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc_16(128, 64, 0, 0, F16 ? 1 : 0);
      const uint32_t w_addr = smem_u32(smem + OFF_W);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int st = it % A_STAGES, as = it & 1;
        const long long m0 = clock64();
        mbar_wait(&acc_empty[as], ((it >> 1) & 1) ^ 1);
        const long long m1 = clock64();
        mbar_wait(&a_full[st], (it / A_STAGES) & 1);
        tc_fence_after();
        const long long m2 = clock64();
        if (prof != nullptr) { prof[blockIdx.x * 16 + 0] += m1 - m0; prof[blockIdx.x * 16 + 1] += m2 - m1; }
        const uint32_t a_addr = smem_u32(smem + OFF_A + st * A_TILE_BYTES);
#pragma unroll
        for (int pos = 0; pos < 4; ++pos) {
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k)
            umma_bf16(tmem_base + as * 256 + pos * 64, make_smem_desc_sw128(a_addr + k * 32, 1024, 16),
                      make_smem_desc_sw128(w_addr + pos * 8192 + k * 32, 1024, 16), idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(&a_empty[st]);
        umma_commit(&acc_full[as]);
        if (prof != nullptr) prof[blockIdx.x * 16 + 2] += clock64() - m2;
      }
    }
  } else if (warp >= 13) {
    // ------------------------------------------------------------ input patch producers (three warps, async copies)
    // Per tile and pooled row (set): input rows 2*prow-1 .. 2*prow+2, columns 2*pcol0-1 .. +131, zero filled outside
    // the spectrogram (conv padding).  4-byte cp.async because the fp32 spectrogram rows (T = 501 floats) are not
    // 16-byte aligned, which rules out TMA.
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int b, rp, wt;
      decode(tile, b, rp, wt);
      const int ps = it % P_STAGES;
      mbar_wait(&p_empty[ps], ((it / P_STAGES) & 1) ^ 1);
      uint8_t* patch = smem + OFF_P + ps * PATCH_BYTES;
      const float* xb = x + static_cast<long long>(b) * H * W;
      const int c0 = 2 * wt * 64 - 1;
#pragma unroll 1
      for (int rr = warp - 13; rr < 8; rr += N_PRODUCERS) {  // rr = set * 4 + wy, rows dealt round-robin to the warps
        const int r = 2 * (2 * rp + (rr >> 2)) - 1 + (rr & 3);
        const bool rok = r >= 0 && r < H;
        const float* xr = xb + static_cast<long long>(rok ? r : 0) * W;
        const uint32_t drow = smem_u32(patch + rr * PATCH_W * 4);
#pragma unroll
        for (int k = 0; k < (PATCH_W + 31) / 32; ++k) {
          const int j = k * 32 + lane;
          if (j < PATCH_W) {
            const int cidx = c0 + j;
            const bool ok = rok && cidx >= 0 && cidx < W;
            const float* src = xr + (ok ? cidx : 0);
            const uint32_t nbytes = ok ? 4u : 0u;  // 0 source bytes -> the 4 destination bytes are zero filled
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(drow + j * 4), "l"(src), "r"(nbytes) : "memory");
          }
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&p_full[ps])) : "memory");
      if (warp == 13 && lane == 0) {
        float inv = 1.0f;  // enhancer.py:96-101: divide by the clip's magnitude maximum if it is > 1e-8
        if (mag_max_bits != nullptr) {
          const float mv = __uint_as_float(__ldg(mag_max_bits + b));
          inv = 1.0f / (mv > 1e-8f ? mv : 1.0f);
        }
        *reinterpret_cast<float*>(patch + 2 * 4 * PATCH_W * 4) = inv;
        mbar_arrive(&p_full[ps]);
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------ window builders: thread -> row m = set * 64 + n
    const int m = threadIdx.x - 256;
    const int n = m & 63, set = m >> 6;
    const int sw = m & 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int st = it % A_STAGES, ps = it % P_STAGES;
      const long long b0 = clock64();
      mbar_wait(&p_full[ps], (it / P_STAGES) & 1);
      const uint8_t* patch = smem + OFF_P + ps * PATCH_BYTES;
      const float inv = *reinterpret_cast<const float*>(patch + 2 * 4 * PATCH_W * 4);
      float cur[16];
#pragma unroll
      for (int wy = 0; wy < 4; ++wy) {
        const float2* pr = reinterpret_cast<const float2*>(patch + ((set * 4 + wy) * PATCH_W + 2 * n) * 4);
        const float2 a = pr[0], c = pr[1];
        cur[wy * 4 + 0] = a.x * inv; cur[wy * 4 + 1] = a.y * inv;
        cur[wy * 4 + 2] = c.x * inv; cur[wy * 4 + 3] = c.y * inv;
      }
      mbar_arrive(&p_empty[ps]);
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (F16) hi[i] = pack_f16x2(cur[2 * i], cur[2 * i + 1]);
        else split_hi_lo<false>(cur[2 * i], cur[2 * i + 1], hi[i], lo[i]);
      }
      const long long b1 = clock64();
      mbar_wait(&a_empty[st], ((it / A_STAGES) & 1) ^ 1);
      const long long b2 = clock64();
      uint8_t* rowp = smem + OFF_A + st * A_TILE_BYTES + m * 128;  // logical 16-byte chunks 0, 1 = hi (2, 3 = lo)
      *reinterpret_cast<uint4*>(rowp + ((0 ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(rowp + ((1 ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      if (!F16) {
        *reinterpret_cast<uint4*>(rowp + ((2 ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(rowp + ((3 ^ sw) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[st]);
      if (prof != nullptr && m == 0) { prof[blockIdx.x * 16 + 3] += b1 - b0; prof[blockIdx.x * 16 + 4] += b2 - b1; prof[blockIdx.x * 16 + 5] += clock64() - b2; }
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM lane = pooled pixel m, columns = channels
    // warp -> (lane quarter q: rows q*32 .. +32, i.e. pooled row set = q >> 1, pooled columns (q & 1)*32 .. +32;
    //          channel half hf: columns hf*32 .. +32)
    const int q = warp & 3, hf = warp >> 2;
    const int set = q >> 1;
    const int n = (q & 1) * 32 + lane;  // pooled pixel within the tile row = staging row
    const int sw = n & 7;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool issuer = (q & 1) == 0 && hf == 0 && lane == 0;  // one store issuer per pooled row (set)
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      uint8_t* stage = smem + OFF_O + (set * 2 + (it & 1)) * O_TILE_BYTES;
      const long long e0 = clock64();
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store of two tiles ago is done
      named_bar_sync(1 + set, 128);
      mbar_wait(&acc_full[as], (it >> 1) & 1);
      tc_fence_after();
      const long long e1 = clock64();
      const uint32_t tb = tmem_base + lane_addr + as * 256 + hf * 32;
      uint8_t* row = stage + n * 128;
      // 4 steps of 8 channels; the four accumulators of the NEXT step are in flight while this one is converted
      uint32_t d[2][4][8];
      auto load_step = [&](int step, uint32_t (&dd)[4][8]) {
        tmem_ld8(tb + 0 * 64 + step * 8, dd[0]);
        tmem_ld8(tb + 1 * 64 + step * 8, dd[1]);
        tmem_ld8(tb + 2 * 64 + step * 8, dd[2]);
        tmem_ld8(tb + 3 * 64 + step * 8, dd[3]);
      };
      load_step(0, d[0]);
#pragma unroll
      for (int step = 0; step < 4; ++step) {
        uint32_t(&cur)[4][8] = d[step & 1];
        tmem_ld_wait8(cur[0], cur[1], cur[2], cur[3]);
        if (step + 1 < 4) {
          load_step(step + 1, d[(step + 1) & 1]);
        } else {  // this thread's part of the accumulator stage is in registers
          tc_fence_before();
          mbar_arrive(&acc_empty[as]);
        }
        const float4 ta = *reinterpret_cast<const float4*>(shift_s + hf * 32 + step * 8);
        const float4 tb4 = *reinterpret_cast<const float4*>(shift_s + hf * 32 + step * 8 + 4);
        const float t[8] = {ta.x, ta.y, ta.z, ta.w, tb4.x, tb4.y, tb4.z, tb4.w};
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          // max over the 2x2 pool window, + BN shift, ReLU:  max(max4, -t) + t
          const float v0 = max3f(max3f(__uint_as_float(cur[0][i]), __uint_as_float(cur[1][i]), __uint_as_float(cur[2][i])),
                                 __uint_as_float(cur[3][i]), -t[i]) + t[i];
          const float v1 = max3f(max3f(__uint_as_float(cur[0][i + 1]), __uint_as_float(cur[1][i + 1]), __uint_as_float(cur[2][i + 1])),
                                 __uint_as_float(cur[3][i + 1]), -t[i + 1]) + t[i + 1];
          pk[i >> 1] = pack2<F16>(v0, v1);
        }
        // channels hf*32 + step*8 .. +8 = logical 16-byte chunk hf*4 + step of this thread's pixel row
        *reinterpret_cast<uint4*>(row + (((hf * 4 + step) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + set, 128);
      if (issuer) {  // rows / columns beyond the pooled image are clipped by the TMA
        int b, rp, wt;
        decode(tile, b, rp, wt);
        tma_store_4d(&tmap_out, stage, 0, wt * 64, 2 * rp + set, b);
        tma_store_commit();
      }
      if (prof != nullptr && q == 0 && lane == 0 && hf == 0) { prof[blockIdx.x * 16 + 6] += e1 - e0; prof[blockIdx.x * 16 + 7] += clock64() - e1; prof[blockIdx.x * 16 + 8] += 1; }
    }
    if (issuer) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int launch_stem_pack(const float* w9c, const float* scale, void* apack, int f16, cudaStream_t s) {
  stem_pack_kernel<<<(4 * 64 * 64 + 255) / 256, 256, 0, s>>>(w9c, scale, reinterpret_cast<uint16_t*>(apack), f16);
  return check_launch("stem_pack");
}

int launch_stem_tc(const float* x, const unsigned* mag_max_bits, const void* apack, const float* shift,
                   const CUtensorMap& tmap_out, int f16, int B, int H, int W, int num_sms, cudaStream_t s) {
  long long* prof = nullptr;
  if (getenv("HVIT_PROF") != nullptr) {
    cudaMalloc(&prof, sizeof(long long) * 16 * num_sms);
    cudaMemset(prof, 0, sizeof(long long) * 16 * num_sms);
  }
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
    if (e != cudaSuccess) {
      once.retry();
      set_error("stem_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return -4;
    }
  }
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_w = (Wo + 63) / 64, row_pairs = (Ho + 1) / 2;
  const long long nt = static_cast<long long>(B) * row_pairs * tiles_w;
  if (nt <= 0 || nt > 0x7FFFFFFF) {
    set_error("stem_tc: bad tile count");
    return -1;
  }
  const int grid = nt < num_sms ? static_cast<int>(nt) : num_sms;
  const uint16_t* ap = reinterpret_cast<const uint16_t*>(apack);
  if (f16)
    launch_pdl(stem_tc_kernel<true>, dim3(grid), dim3(ST_THREADS), ST_SMEM, s, x, mag_max_bits, ap, shift, tmap_out, H, W,
               Ho, Wo, tiles_w, row_pairs, static_cast<int>(nt), prof);
  else
    launch_pdl(stem_tc_kernel<false>, dim3(grid), dim3(ST_THREADS), ST_SMEM, s, x, mag_max_bits, ap, shift, tmap_out, H, W,
               Ho, Wo, tiles_w, row_pairs, static_cast<int>(nt), prof);
  if (prof != nullptr) {
    cudaDeviceSynchronize();
    long long h[16 * 256];
    cudaMemcpy(h, prof, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost);
    cudaFree(prof);
    double a[16] = {0};
    for (int c = 0; c < grid; ++c)
      for (int k = 0; k < 16; ++k) a[k] += static_cast<double>(h[c * 16 + k]) / grid;
    fprintf(stderr, "[stem prof] tiles/cta %.1f | mma wait_acc_empty %.0f wait_b_full %.0f issue %.0f | builder load+split %.0f wait_b_empty %.0f write %.0f | epi wait_acc_full %.0f drain %.0f (cycles per CTA)\n",
            a[8], a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
  }
  return check_launch("stem_tc");
}

}  // namespace hvit
