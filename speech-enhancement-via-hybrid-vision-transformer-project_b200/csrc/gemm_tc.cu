// tcgen05 implicit-GEMM for sm_100a: 16-bit (bf16 or fp16) operands staged by TMA (128B swizzle), fp32 accumulation in TMEM,
// persistent tiles, warp-specialised (1 TMA warp, 1 MMA warp, 8 epilogue warps), double-buffered accumulator so
// the epilogue of tile i overlaps the main loop of tile i+1.
//
// One kernel covers every dense contraction of the HybridViT forward (reference models/hybrid_vit.py:396-469):
//   IG_PLAIN  qkv / proj / fc1 / fc2 / to_feature_map / skip 1x1 projections   (attention.py:83,109; components.py:223-229)
//   IG_CONV3  encoder and decoder 3x3 convolutions as implicit GEMM            (components.py:54-63,149-158)
//   IG_UP2    nearest-x2 upsample + 3x3 conv as four 2x2 parity convolutions   (components.py:145-158)
//   IG_PATCH  4x4 / stride 4 patch embedding                                   (components.py:275-280)
// The A operand of the conv modes is fetched with shifted multi-dimensional TMA boxes on the NHWC activation;
// out-of-bounds box elements are zero-filled by the TMA unit, which implements the conv zero padding.
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace hvit {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_EPI_WARPS = 8;  // two per TMEM lane quarter: they split the 32-column chunks (even / odd)
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;

template <int BLOCK_N>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * 1024;  // per epilogue warp: scale/shift of its chunks
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256;
};

struct TileCoord {
  int n0, m0, b, h0, w0, par;
};

__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile, int n_tiles_n, int block_n) {
  TileCoord c;
  c.n0 = (tile % n_tiles_n) * block_n;
  int r = tile / n_tiles_n;
  c.m0 = 0; c.b = 0; c.h0 = 0; c.w0 = 0; c.par = 0;
  if (p.mode == IG_PLAIN) {
    c.m0 = r * BLOCK_M;
  } else {
    c.w0 = (r % p.tiles_w) * p.Wt;
    r /= p.tiles_w;
    c.h0 = (r % p.tiles_h) * p.Ht;
    r /= p.tiles_h;
    if (p.mode == IG_UP2) {
      c.par = r & 3;
      r >>= 2;
    }
    c.b = r;
  }
  return c;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
igemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const IgemmParams p, int num_tiles, int n_tiles_n) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + C::STAGING_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kblocks = p.K / BLOCK_K;
  const int cblocks = p.Cin / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], NUM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord c = decode_tile(p, tile, n_tiles_n, BLOCK_N);
        const int py = c.par >> 1, px = c.par & 1;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          if (p.mode == IG_PLAIN) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, c.m0);
          } else {
            const int tap = kb / cblocks;
            const int c0 = (kb - tap * cblocks) * BLOCK_K;
            if (p.mode == IG_CONV3) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              tma_load_4d(sa, &tmap_a, &full_bar[stage], c0, c.w0 + dx, c.h0 + dy, c.b);
            } else if (p.mode == IG_UP2) {
              const int dy = (tap >> 1) + py - 1, dx = (tap & 1) + px - 1;
              tma_load_4d(sa, &tmap_a, &full_bar[stage], c0, c.w0 + dx, c.h0 + dy, c.b);
            } else {  // IG_PATCH
              const int ky = tap / p.patch, kx = tap % p.patch;
              tma_load_5d(sa, &tmap_a, &full_bar[stage], c0, kx, c.w0, ky, c.b * p.Hq + c.h0);
            }
          }
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, c.n0 + c.par * p.N);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (elect_one()) {
      const uint32_t idesc = make_idesc_16(BLOCK_M, BLOCK_N, 0, 0, p.f16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(a_addr + k * UMMA_K * 2, 1024, 16);
            const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * UMMA_K * 2, 1024, 16);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> regs -> global
    // 8 warps: warp w may only read TMEM lanes [32*(w%4), +32); the two warps of a lane quarter take the even / odd
    // 32-column chunks.  Each epilogue warp is alone (or one of two) on its SM sub-partition, so every dependent
    // latency is exposed: per-channel constants are therefore staged in shared memory BEFORE waiting for the
    // accumulator, TMEM loads are double-buffered, and residual loads are issued before the TMEM wait.
    const int ew = warp - 2;
    const int sub = warp & 3;           // TMEM lane quarter this warp may read
    const int grp = ew >> 2;            // chunk parity handled by this warp
    const int m = sub * 32 + lane;      // accumulator row == TMEM lane
    float* cst = reinterpret_cast<float*>(staging + ew * 1024);  // [KCH][scale 32 | shift 32]
    constexpr int NCH = BLOCK_N / 32;
    constexpr int KCH = NCH / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile(p, tile, n_tiles_n, BLOCK_N);
      bool valid;
      long long out_row;
      if (p.mode == IG_PLAIN) {
        const int gm = c.m0 + m;
        valid = gm < p.M;
        out_row = gm;
      } else {
        const int hl = m / p.Wt, wl = m - hl * p.Wt;
        const int h = c.h0 + hl, w = c.w0 + wl;
        if (p.mode == IG_CONV3) {
          if (p.pool) {  // Wt == 16, Ht == 8: this warp holds rows 2*sub, 2*sub+1 of the tile
            const int ph = (c.h0 >> 1) + sub, pw = (c.w0 + (lane & 15)) >> 1;
            valid = ((lane & 17) == 0) && ph < p.Ho && pw < p.Wo;
            out_row = (static_cast<long long>(c.b) * p.HoPitch + ph) * p.Wo + pw;
          } else {
            valid = h < p.H && w < p.W;
            out_row = (static_cast<long long>(c.b) * p.HoPitch + h) * p.Wo + w;
          }
        } else if (p.mode == IG_UP2) {
          valid = h < p.H && w < p.W;
          out_row = (static_cast<long long>(c.b) * p.HoPitch + 2 * h + (c.par >> 1)) * p.Wo + 2 * w + (c.par & 1);
        } else {
          valid = h < p.Hp && w < p.Wp;
          out_row = static_cast<long long>(c.b) * p.Hp * p.Wp + h * p.Wp + w;
        }
      }
      const float* res_row = nullptr;
      if (p.residual != nullptr && valid) {
        const long long rr = p.res_mod > 0 ? (out_row % p.res_mod) : out_row;
        res_row = p.residual + rr * p.ldr;
      }
      // per-channel constants of this warp's chunks -> smem (latency hidden behind this tile's main loop)
#pragma unroll
      for (int k = 0; k < KCH; ++k) {
        const int col = c.n0 + (2 * k + grp) * 32 + lane;
        cst[k * 64 + lane] = p.scale != nullptr ? __ldg(p.scale + col) : 1.0f;
        cst[k * 64 + 32 + lane] = p.shift != nullptr ? __ldg(p.shift + col) : 0.0f;
      }
      __syncwarp();

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BLOCK_N;

      auto process = [&](int k, uint32_t (&cur)[32], uint32_t (&nxt)[32], bool has_next) {
        const int chunk = 2 * k + grp;
        const int col = c.n0 + chunk * 32;
        float4 res[8];
        if (res_row != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) res[j] = *reinterpret_cast<const float4*>(res_row + col + 4 * j);
        }
        tmem_ld_wait(cur);
        if (has_next) tmem_ld32(t_base + (chunk + 2) * 32, nxt);
        float v[32];
        const float4* cs = reinterpret_cast<const float4*>(cst + k * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 sc = cs[j], sh = cs[8 + j];
          v[4 * j + 0] = fmaf(__uint_as_float(cur[4 * j + 0]), sc.x, sh.x);
          v[4 * j + 1] = fmaf(__uint_as_float(cur[4 * j + 1]), sc.y, sh.y);
          v[4 * j + 2] = fmaf(__uint_as_float(cur[4 * j + 2]), sc.z, sh.z);
          v[4 * j + 3] = fmaf(__uint_as_float(cur[4 * j + 3]), sc.w, sh.w);
        }
        if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else if (p.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        }
        if (p.pool) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xFFFFFFFFu, v[j], 1));
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xFFFFFFFFu, v[j], 16));
          }
        }
        if (valid) {
          if (res_row != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] += res[j].x; v[4 * j + 1] += res[j].y; v[4 * j + 2] += res[j].z; v[4 * j + 3] += res[j].w;
            }
          }
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldc + col;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            bf16* o = reinterpret_cast<bf16*>(p.out) + out_row * p.ldc + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 q;
              q.x = pack_16x2(v[j], v[j + 1], p.f16);
              q.y = pack_16x2(v[j + 2], v[j + 3], p.f16);
              q.z = pack_16x2(v[j + 4], v[j + 5], p.f16);
              q.w = pack_16x2(v[j + 6], v[j + 7], p.f16);
              *reinterpret_cast<uint4*>(o + j) = q;
            }
          }
        }
      };

      uint32_t ra[32], rb[32];
      tmem_ld32(t_base + grp * 32, ra);
#pragma unroll
      for (int k = 0; k < KCH; k += 2) {
        process(k, ra, rb, k + 1 < KCH);
        if (k + 1 < KCH) process(k + 1, rb, ra, k + 2 < KCH);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int BLOCK_N>
int launch_impl(const IgemmParams& p, const CUtensorMap& ta, const CUtensorMap& tb, int num_tiles, int n_tiles_n,
                int num_sms, cudaStream_t stream) {
  using C = Cfg<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("igemm_tc: cudaFuncSetAttribute(%d B smem) failed: %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return -4;
    }
    configured = true;
  }
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  igemm_tc_kernel<BLOCK_N><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p, num_tiles, n_tiles_n);
  const cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    cudaFuncAttributes fa;
    memset(&fa, 0, sizeof(fa));
    cudaFuncGetAttributes(&fa, igemm_tc_kernel<BLOCK_N>);
    set_error("igemm_tc<%d>: %s (threads %d, dyn smem %d, regs %d, static smem %zu, local %zu, maxThreadsPerBlock %d, "
              "maxDynSmem %d)", BLOCK_N, cudaGetErrorString(le), NUM_THREADS, C::SMEM_BYTES, fa.numRegs,
              fa.sharedSizeBytes, fa.localSizeBytes, fa.maxThreadsPerBlock, fa.maxDynamicSharedSizeBytes);
    return -4;
  }
  return 0;
}

}  // namespace

int launch_igemm_tc(const IgemmParams& p, const CUtensorMap& ta, const CUtensorMap& tb, int block_n, int num_sms,
                    cudaStream_t stream) {
  if (p.K % BLOCK_K != 0 || p.N % block_n != 0 || (p.mode != IG_PLAIN && p.Cin % BLOCK_K != 0)) {
    set_error("igemm_tc: unsupported shape N=%d K=%d Cin=%d block_n=%d", p.N, p.K, p.Cin, block_n);
    return -1;
  }
  if (p.residual != nullptr && !p.out_f32) {
    set_error("igemm_tc: a residual input needs an fp32 output");
    return -1;
  }
  if (p.pool && !(p.mode == IG_CONV3 && p.Wt == 16 && p.Ht == 8)) {
    set_error("igemm_tc: fused pool needs a 16x8 spatial tile");
    return -1;
  }
  const int n_tiles_n = p.N / block_n;
  long long m_tiles;
  if (p.mode == IG_PLAIN) {
    m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  } else {
    m_tiles = static_cast<long long>(p.B) * p.tiles_h * p.tiles_w * (p.mode == IG_UP2 ? 4 : 1);
  }
  const long long nt = m_tiles * n_tiles_n;
  if (nt <= 0 || nt > 0x7FFFFFFF) {
    set_error("igemm_tc: bad tile count %lld", nt);
    return -1;
  }
  switch (block_n) {
    case 256: return launch_impl<256>(p, ta, tb, static_cast<int>(nt), n_tiles_n, num_sms, stream);
    case 128: return launch_impl<128>(p, ta, tb, static_cast<int>(nt), n_tiles_n, num_sms, stream);
    case 64: return launch_impl<64>(p, ta, tb, static_cast<int>(nt), n_tiles_n, num_sms, stream);
    default: set_error("igemm_tc: block_n must be 64/128/256"); return -1;
  }
}

}  // namespace hvit
