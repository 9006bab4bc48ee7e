// Fused flash-style multi-head self-attention on tcgen05 (head_dim 64), replacing the reference's unfused
// q@k^T -> softmax -> @v that materialises [B, heads, N, N] (models/attention.py:86-105).
//
// One CTA per (clip, head, 128-query tile); 6 warps:
//   warp 0     TMA producer: Q once, K/V blocks of 128 keys double-buffered (boxes straight out of the
//              [B*N, 3D] qkv activation: head h of q|k|v lives at columns s*D + h*64)
//   warp 1     MMA issuer:   S = Q K^T (128x128, fp32 in TMEM), then O_blk = P V (128x64, fp32 in TMEM)
//   warps 2-5  softmax:      one query row per thread (TMEM lane == row): online max/sum in fp32 registers,
//              P written to shared memory as bf16 in the 128B-swizzled K-major layout the PV MMA reads,
//              running output kept in registers and rescaled per block (no TMEM read-modify-write).
// Keys beyond N (last block) are masked to -inf; TMA zero-fills out-of-range rows of Q/K/V.
#include "common.cuh"
#include "kernels.h"

namespace hvit {
namespace {

constexpr int AT_THREADS = 192;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16
constexpr int OFF_Q = 0;
constexpr int OFF_K = TILE_BYTES;          // 2 stages
constexpr int OFF_V = 3 * TILE_BYTES;      // 2 stages
constexpr int OFF_P = 5 * TILE_BYTES;      // 128 x 128 bf16 = two 64-key K-major blocks
constexpr int OFF_BAR = 7 * TILE_BYTES;
constexpr int AT_SMEM = OFF_BAR + 256;
constexpr int TMEM_COLS = 256;             // S: cols [0,128), O_blk: cols [128,192)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, bf16* __restrict__ out, int N, int D, float scale_log2,
               int f16) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B swizzle needs 1024-byte aligned tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* v_full = bars + 3;      // [2]
  uint64_t* kv_empty = bars + 5;    // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (N + 127) / 128;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(smem + OFF_Q, &tmap_qkv, q_full, h * 64, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_3d(smem + OFF_K + st * TILE_BYTES, &tmap_qkv, &k_full[st], D + h * 64, j * 128, b);
        mbar_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_3d(smem + OFF_V + st * TILE_BYTES, &tmap_qkv, &v_full[st], 2 * D + h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_16(128, 128, 0, 0, f16);   // Q (K-major) x K (K-major)
      const uint32_t idesc_pv = make_idesc_16(128, 64, 0, 1, f16);   // P (K-major) x V (MN-major)
      const uint32_t q_addr = smem_u32(smem + OFF_Q);
      const uint32_t p_addr = smem_u32(smem + OFF_P);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + OFF_K + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_s, make_smem_desc_sw128(q_addr + k * 32, 1024, 16),
                    make_smem_desc_sw128(k_addr + k * 32, 1024, 16), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + OFF_V + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_o, make_smem_desc_sw128(p_addr + (k >> 2) * TILE_BYTES + (k & 3) * 32, 1024, 16),
                    make_smem_desc_sw128(v_addr + k * 2048, 1024, 1024), idesc_pv, k != 0 ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
        if (j + 1 < nkv) issue_s(j + 1);
      }
    }
  } else {
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    uint8_t* p_row = smem + OFF_P + row * 128;
    const int sw = row & 7;
    float m_run = -INFINITY, l_run = 0.f;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const int nvalid = min(128, N - j * 128);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // Instruction diet (ncu: this kernel is issue-bound): the row maximum is taken on the raw scores (one FMNMX
      // per element, scaled once), exp2 is a bare MUFU.EX2 fed by one FFMA, and key masking only exists in the code
      // path of a partial last block.
      const bool full = nvalid == 128;
      float mraw = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + c * 32, r);
        tmem_ld_wait(r);
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mraw = fmaxf(mraw, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < nvalid) mraw = fmaxf(mraw, __uint_as_float(r[i]));
        }
      }
      const float mx = fmaxf(m_run, mraw * scale_log2);
      const float alpha = ex2_approx(m_run - mx);
      float rowsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + c * 32, r);
        tmem_ld_wait(r);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pv[i] = ex2_approx(fmaf(__uint_as_float(r[q * 8 + i]), scale_log2, -mx));
            if (!full && c * 32 + q * 8 + i >= nvalid) pv[i] = 0.f;
          }
          rowsum += ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
          const int cidx = c * 4 + q;  // 16-byte chunk index along the 128 keys
          uint4 pk;
          pk.x = pack_16x2(pv[0], pv[1], f16);
          pk.y = pack_16x2(pv[2], pv[3], f16);
          pk.z = pack_16x2(pv[4], pv[5], f16);
          pk.w = pack_16x2(pv[6], pv[7], f16);
          *reinterpret_cast<uint4*>(p_row + (cidx >> 3) * TILE_BYTES + (((cidx & 7) ^ sw) << 4)) = pk;
        }
      }
      l_run = l_run * alpha + rowsum;
      m_run = mx;
      fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor-core async proxy
      tc_fence_before();
      mbar_arrive(p_full);

      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_o + lane_addr + c * 32, r);
        tmem_ld_wait(r);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(r[i]));
      }
    }
    if (q0 + row < N) {
      const float inv = 1.0f / l_run;
      bf16* dst = out + (static_cast<long long>(b) * N + q0 + row) * D + h * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        uint4 pk;
        pk.x = pack_16x2(o[i] * inv, o[i + 1] * inv, f16);
        pk.y = pack_16x2(o[i + 2] * inv, o[i + 3] * inv, f16);
        pk.z = pack_16x2(o[i + 4] * inv, o[i + 5] * inv, f16);
        pk.w = pack_16x2(o[i + 6] * inv, o[i + 7] * inv, f16);
        *reinterpret_cast<uint4*>(dst + i) = pk;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int launch_attn_tc(const CUtensorMap& tmap_qkv, void* out16, int f16, int B, int N, int heads, int D, float scale,
                   cudaStream_t stream) {
  if (D != heads * 64) {
    set_error("attention: head_dim must be 64 (D=%d heads=%d)", D, heads);
    return -1;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_error("attn_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return -4;
    }
    configured = true;
  }
  dim3 grid((N + 127) / 128, heads, B);
  attn_tc_kernel<<<grid, AT_THREADS, AT_SMEM, stream>>>(tmap_qkv, reinterpret_cast<bf16*>(out16), N, D,
                                                         scale * 1.4426950408889634f, f16);
  return check_launch("attn_tc");
}

}  // namespace hvit
