// Fused flash-style multi-head self-attention on tcgen05 (head_dim 64), replacing the reference's unfused
// q@k^T -> softmax -> @v that materialises [B, heads, N, N] (models/attention.py:86-105).
//
// One CTA per (clip, head, PAIR of 128-query tiles), one CTA per SM, 12 warps:
//   warp 8      TMA producer: Q tiles once, K/V blocks of 128 keys double-buffered (boxes straight out of the
//               [B*N, 3D] qkv activation: head h of q|k|v lives at columns s*D + h*64)
//   warp 9      MMA issuer (one thread): S_w = Q_w K^T (128x128 fp32 in TMEM) and O_w += P_w V (128x64 fp32 in TMEM)
//               for the two query tiles w = A, B
//   warps 0-3   softmax group A, warps 4-7 softmax group B: one query row per thread (TMEM lane == row)
//
// The two softmax groups work on different query tiles against the same K/V stream, so while one group is in its
// exp phase (MUFU-bound) the tensor core computes the other group's S / PV and nobody waits on the MMA latency.
// Per key block a thread pulls its whole 128-score row out of TMEM in one go (the S buffer is released to the MMA
// warp immediately, S(j+1) is computed while softmax(j) runs), takes the row max with 3-input max, evaluates
// p = 2^(s*scale*log2e - m) with packed FFMA2 + MUFU.EX2, and writes P as 16-bit into the 128B-swizzled K-major
// layout the PV MMA reads.  O stays in TMEM for the whole key loop (accumulating MMAs); it is only touched when the
// running maximum has grown by more than 2^8 since the last rescale ("lazy rescale": P and the row sum simply use
// the stale maximum, which is exact in real arithmetic and harmless in fp16/bf16/fp32 ranges), and once at the end
// for the 1/l normalisation + TMA store.
// Keys beyond N (last block) are masked to -inf; TMA zero-fills out-of-range rows of Q/K/V and clips the store.
#include "common.cuh"
#include "kernels.h"

namespace hvit {
namespace {

constexpr int AT_THREADS = 384;  // warpgroups: 0 = softmax A, 1 = softmax B, 2 = {TMA producer, MMA issuer, 2 idle warps}
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 16-bit
constexpr int KV_STAGES = 3;
constexpr int OFF_Q = 0;                                  // [2] query tiles
constexpr int OFF_K = 2 * TILE_BYTES;                     // [KV_STAGES]
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;     // [KV_STAGES]
constexpr int OFF_P = OFF_V + KV_STAGES * TILE_BYTES;     // [2] x (128 x 128 16-bit = two 64-key K-major blocks)
constexpr int OFF_O = OFF_P + 4 * TILE_BYTES;                // [2] output staging tiles (TMA store)
constexpr int OFF_BAR = OFF_O + 2 * TILE_BYTES;
constexpr int AT_SMEM = OFF_BAR + 256;
constexpr int TMEM_COLS = 512;  // S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384)
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + r, |r| <= 0.5, degree-4 minimax 2^r (relative
// error 3.4e-6 - far below the 16-bit rounding of P), exponent added as an integer.  x is clamped at -127, where
// the result is exactly 0 (bits 0x3F800000 - 127 * 2^23), like ex2.approx.ftz below its normal range; results for
// x in (-127, -126) are (harmless) fp32 subnormal bit patterns.  Returns the fp32 BIT PATTERN.
__device__ __forceinline__ uint32_t exp2_poly_bits(float x) {
  const float xf = fmaxf(x, -127.0f);
  const float t = xf + 12582912.0f;  // 1.5 * 2^23: the integer n = rint(xf) lands in the low mantissa bits
  const float r = xf - (t - 12582912.0f);
  float p = 0.009570018388330936f;
  p = fmaf(p, r, 0.05591766536235809f);
  p = fmaf(p, r, 0.240247443318367f);
  p = fmaf(p, r, 0.6931218504905701f);
  p = fmaf(p, r, 1.0f);
  return __float_as_uint(p) + (__float_as_uint(t) << 23);
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// exp + row-sum + pack of one 128-score row held in registers; P chunks go to shared memory as they are produced.
//
// The fp32 -> 16-bit conversion instruction (F2FP) shares the 16-lane/clk XU pipe with MUFU.EX2 (measured: with F2FP
// packing the exp phase ran at 26 cycles per warp-element instead of MUFU's 8), so P is converted with integer
// arithmetic on the FMA/ALU pipes instead: the exponent re-bias of fp16 (127 - 15 = 112) is folded into the exp2
// argument (nbias = m + 112; ex2.approx.ftz flushes what would be an fp16 subnormal to exactly 0), after which
//   16-bit float = upper half of (bits(e) * mul + 0x8000)      mul = 8 (fp16: 23 -> 10 mantissa bits), 1 (bf16)
// (round-half-up), and two results are merged with one PRMT.  The row sum uses the same (2^-112-scaled) values; the
// scale is undone in the final 1/l normalisation.
__device__ __forceinline__ float exp_pack_row(uint32_t (&s)[128], float scale_log2, float nbias, uint32_t mul,
                                              uint8_t* p_row, int sw) {
  const float2 sc2 = make_float2(scale_log2, scale_log2);
  const float2 nm2 = make_float2(-nbias, -nbias);
  float2 sum0 = make_float2(0.f, 0.f), sum1 = make_float2(0.f, 0.f);
  // Software pipeline with a lag of LAG chunks between the MUFU stage and the convert / sum / store stage: each
  // softmax group has ONE warp per scheduler, so a MUFU result consumed right after its issue would stall the warp
  // for the MUFU latency (measured: 16 cycles per score instead of the unit's 8).
  constexpr int LAG = 3;
  auto exp_chunk = [&](int c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __ffma2_rn(make_float2(__uint_as_float(s[8 * c + 2 * i]), __uint_as_float(s[8 * c + 2 * i + 1])),
                                  sc2, nm2);
      s[8 * c + 2 * i] = __float_as_uint(ex2_approx(t.x));
      // (measured: moving half of the exponentials to exp2_poly_bits makes this phase 30 % SLOWER - the phase is
      // bound by instruction issue of the single warp per scheduler, not by the MUFU unit)
      s[8 * c + 2 * i + 1] = __float_as_uint(ex2_approx(t.y));
    }
  };
#pragma unroll
  for (int c = 0; c < LAG; ++c) exp_chunk(c);
#pragma unroll
  for (int c = 0; c < 16; ++c) {  // 16-byte chunk c = keys 8c .. 8c+7
    if (c + LAG < 16) exp_chunk(c + LAG);
    uint32_t pk[4];
    float2 t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      t[i] = make_float2(__uint_as_float(s[8 * c + 2 * i]), __uint_as_float(s[8 * c + 2 * i + 1]));
      pk[i] = __byte_perm(s[8 * c + 2 * i] * mul + 0x8000u, s[8 * c + 2 * i + 1] * mul + 0x8000u, 0x7632);
    }
    sum0 = __fadd2_rn(sum0, __fadd2_rn(t[0], t[1]));
    sum1 = __fadd2_rn(sum1, __fadd2_rn(t[2], t[3]));
    *reinterpret_cast<uint4*>(p_row + (c >> 3) * TILE_BYTES + (((c & 7) ^ sw) << 4)) =
        make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  return (sum0.x + sum0.y) + (sum1.x + sum1.y);
}

// PROBS (return_attentions=True, hybrid_vit.py:422-450): the softmax groups also write the attention probabilities
// [B, heads, N, N] fp32 - un-normalised p = 2^(s - m_used) while the key blocks stream by (m_used is the lazily updated
// maximum of that block), then, once the row's final maximum and sum are known, each thread rescales its own row in
// place (its 2 KB are still in L2).  At most PROBS_MAX_KV key blocks.
constexpr int PROBS_MAX_KV = 10;
template <bool PROF, bool PROBS = false>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out, int N,
               int D, int heads, int n_items, float scale_log2, int f16, long long* prof, const int* __restrict__ geo,
               int pingpong, float* __restrict__ probs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B swizzle needs 1024-byte aligned tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                      // [2]  Q tile of group w's current item landed
  uint64_t* q_empty = bars + 2;                 // [2]  every S MMA of group w's current item has completed (Q_w may be reloaded)
  uint64_t* k_full = bars + 4;                  // [KV_STAGES]
  uint64_t* v_full = k_full + KV_STAGES;        // [KV_STAGES]
  uint64_t* kv_empty = v_full + KV_STAGES;      // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;      // [2]  S_w(block) complete in TMEM
  uint64_t* s_free = s_full + 2;                // [2]  S_w(block) copied to registers by all 128 rows
  uint64_t* p_full = s_free + 2;                // [2]  P_w(block) in shared memory (and O_w rescaled if needed)
  uint64_t* pv_done = p_full + 2;               // [2]  O_w += P_w(block) V(block) complete
  uint64_t* o_free = pv_done + 2;               // [2]  final O_w of an item copied to registers by all 128 rows
  uint64_t* lag_bar = o_free + 2;               // group A has finished the row maximum of its first block (see below)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lag_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qpairs = (N + 255) / 256;
  // keys of clip b: all N tokens, or - variable-length batch - its own valid prefix N_b (token order is the clip's own,
  // so validity is a prefix): key blocks beyond N_b are skipped by every role, the tail of the last one is masked.
  // Every QUERY tile is still computed (rows >= N_b are padding: finite, never read by a valid row).
  auto clip_tokens = [&](int b) -> int { return geo != nullptr ? __ldg(geo + b * GEO_STRIDE + GEO_NTOK) : N; };
  // persistent work loop: item -> (clip b, head h, query-tile pair); the pairs of one (b, h) are adjacent items, so
  // they run at the same time on neighbouring CTAs and share K / V in L2
  auto decode = [&](int item, int& b, int& h, int& q0) {
    q0 = (item % qpairs) * 256;
    const int bh = item / qpairs;
    h = bh % heads;
    b = bh / heads;
  };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_out);
  }
  if (warp == 9 && lane == 0) {
    for (int w = 0; w < 2; ++w) {
      mbar_init(&q_full[w], 1);
      mbar_init(&q_empty[w], 1);
    }
    mbar_init(lag_bar, 128);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_full[w], 1);
      mbar_init(&s_free[w], 128);
      mbar_init(&p_full[w], 128);
      mbar_init(&pv_done[w], 1);
      mbar_init(&o_free[w], 128);
    }
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();
  griddep_wait();  // PDL: everything above overlapped the previous kernel's tail

  // Register re-allocation between warpgroups (the kernel is compiled for 168 registers at 384 threads): the
  // producer / MMA warpgroup gives registers back, the softmax warpgroups take them - a 128-score row plus the
  // exp working set must stay in registers (with 168 the row spilled to local memory: 3x slower exp phase).
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 8) {
      // ------------------------------------------------------------ TMA producer
      // Three independent load streams over this CTA's items - Q tiles of group A, Q tiles of group B, K/V blocks -
      // each issued as soon as ITS buffer is free (non-blocking polls): the two softmax groups run half a key block out
      // of phase on purpose (see the MMA issuers), and a producer that waited for one group's buffer before serving the
      // other would pull them back into lock-step at every item boundary.
      if (elect_one()) {
        const int stride = gridDim.x;
        auto has_b = [&](int item) { return (item % qpairs) * 256 + 128 < N; };
        int itemA = blockIdx.x, itemB = blockIdx.x, itemK = blockIdx.x;
        while (itemB < n_items && !has_b(itemB)) itemB += stride;
        uint32_t nA = 0, nB = 0, kv = 0;
        int jK = 0, nkvK = 0;
        if (itemK < n_items) nkvK = (clip_tokens(itemK / qpairs / heads) + 127) / 128;
        while (itemA < n_items || itemB < n_items || itemK < n_items) {
          bool progress = false;
          if (itemK < n_items) {
            const int st = kv % KV_STAGES;
            if (mbar_try_wait(&kv_empty[st], ((kv / KV_STAGES) & 1) ^ 1)) {
              int b, h, q0;
              decode(itemK, b, h, q0);
              mbar_expect_tx(&k_full[st], TILE_BYTES);
              tma_load_3d(smem + OFF_K + st * TILE_BYTES, &tmap_qkv, &k_full[st], D + h * 64, jK * 128, b);
              mbar_expect_tx(&v_full[st], TILE_BYTES);
              tma_load_3d(smem + OFF_V + st * TILE_BYTES, &tmap_qkv, &v_full[st], 2 * D + h * 64, jK * 128, b);
              ++kv;
              if (++jK == nkvK) {
                jK = 0;
                itemK += stride;
                if (itemK < n_items) nkvK = (clip_tokens(itemK / qpairs / heads) + 127) / 128;
              }
              progress = true;
            }
          }
          if (itemA < n_items && mbar_try_wait(&q_empty[0], (nA & 1) ^ 1)) {
            int b, h, q0;
            decode(itemA, b, h, q0);
            mbar_expect_tx(&q_full[0], TILE_BYTES);
            tma_load_3d(smem + OFF_Q, &tmap_qkv, &q_full[0], h * 64, q0, b);
            ++nA;
            itemA += stride;
            progress = true;
          }
          if (itemB < n_items && mbar_try_wait(&q_empty[1], (nB & 1) ^ 1)) {
            int b, h, q0;
            decode(itemB, b, h, q0);
            mbar_expect_tx(&q_full[1], TILE_BYTES);
            tma_load_3d(smem + OFF_Q + TILE_BYTES, &tmap_qkv, &q_full[1], h * 64, q0 + 128, b);
            ++nB;
            itemB += stride;
            while (itemB < n_items && !has_b(itemB)) itemB += stride;
            progress = true;
          }
          if (!progress) __nanosleep(128);  // (a tight poll would take issue slots from the softmax warps on this scheduler)
        }
      }
    } else if (warp == 9 || warp == 10) {
      // ------------------------------------------------------------ MMA issuers: warp 9 serves softmax group A,
      // warp 10 group B.  One issuing thread per group keeps the two S -> softmax -> PV chains independent (a single
      // in-order issuer waits on one group's barrier while the other group's MMA is already due, which locks the two
      // groups' MUFU-bound exp phases together); the tensor core serialises the MMAs of the two threads by itself.
      // Barriers shared by both chains (q_empty, kv_empty) count one commit per issuer.
      if (elect_one()) {
        const int w = warp - 9;
        const uint32_t idesc_s = make_idesc_16(128, 128, 0, 0, f16);   // Q (K-major) x K (K-major)
        const uint32_t idesc_pv = make_idesc_16(128, 64, 0, 1, f16);   // P (K-major) x V (MN-major)
        uint32_t it = 0, kv = 0;
        uint32_t blk = 0;   // key blocks processed so far by this softmax group (barrier phases)
        uint32_t itw = 0;   // items processed so far by this softmax group
        auto issue_s = [&](uint32_t kvi) {
          const int st = kvi % KV_STAGES;
          mbar_wait(&k_full[st], (kvi / KV_STAGES) & 1);
          tc_fence_after();
          const uint32_t q_addr = smem_u32(smem + OFF_Q + w * TILE_BYTES);
          const uint32_t k_addr = smem_u32(smem + OFF_K + st * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + w * 128, make_smem_desc_sw128(q_addr + k * 32, 1024, 16),
                      make_smem_desc_sw128(k_addr + k * 32, 1024, 16), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[w]);
        };
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
          int b, h, q0;
          decode(item, b, h, q0);
          const int nkv = (clip_tokens(b) + 127) / 128;
          if (w == 1 && !(q0 + 128 < N)) {  // no second query tile: only keep the shared K/V barriers' counts
            // (each arrival waits for the matching "full" phase first, so it can never land in an earlier phase)
            for (int j = 0; j < nkv; ++j, ++kv) {
              mbar_wait(&k_full[kv % KV_STAGES], (kv / KV_STAGES) & 1);
              mbar_arrive(&kv_empty[kv % KV_STAGES]);
            }
            continue;
          }
          mbar_wait(&q_full[w], itw & 1);
          if (blk > 0) {  // the group's last score block of its previous item has left TMEM
            mbar_wait(&s_free[w], (blk - 1) & 1);
            tc_fence_after();
          } else if (w == 1) {
            // De-phasing: group B's first S is issued only when group A has finished the row maximum of ITS first block
            // and enters its exp phase.  From then on B runs about half a key block behind A (nothing re-aligns them:
            // separate Q buffers / barriers, the K/V ring absorbs the skew), so one group's MUFU-bound exp phase overlaps
            // the other's TMEM loads / row maximum / barrier waits instead of both hitting the MUFU unit together and
            // both leaving it idle afterwards (measured before: 2 230-cycle exp phases in lock-step, 4 440 cycles per key
            // block against a 2 048-cycle MUFU floor).
            mbar_wait(lag_bar, 0);
          }
          issue_s(kv);
          if (nkv == 1) umma_commit(&q_empty[w]);
          for (int j = 0; j < nkv; ++j, ++kv) {
            const int st = kv % KV_STAGES;
            if (j + 1 < nkv) {  // S_w(j+1) as soon as the softmax group has S_w(j) in registers
              mbar_wait(&s_free[w], (blk + j) & 1);
              tc_fence_after();
              issue_s(kv + 1);
              if (j + 2 == nkv) umma_commit(&q_empty[w]);  // last S MMA of this item issued
            }
            mbar_wait(&v_full[st], (kv / KV_STAGES) & 1);
            mbar_wait(&p_full[w], (blk + j) & 1);
            if (j == 0 && itw > 0) mbar_wait(&o_free[w], (itw - 1) & 1);  // previous item's O_w read out
            tc_fence_after();
            const uint32_t p_addr = smem_u32(smem + OFF_P + w * 2 * TILE_BYTES);
            const uint32_t v_addr = smem_u32(smem + OFF_V + st * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16(tmem_base + 256 + w * 64,
                        make_smem_desc_sw128(p_addr + (k >> 2) * TILE_BYTES + (k & 3) * 32, 1024, 16),
                        make_smem_desc_sw128(v_addr + k * 2048, 1024, 1024), idesc_pv, (j | k) != 0 ? 1u : 0u);
            umma_commit(&pv_done[w]);
            umma_commit(&kv_empty[st]);  // this group's share: K / V stage free once both groups' MMAs completed
          }
          blk += nkv;
          ++itw;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax groups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int w = warp >> 2;  // 0: tile A, 1: tile B
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    const uint32_t tmem_s = tmem_base + w * 128 + lane_addr;
    const uint32_t tmem_o = tmem_base + 256 + w * 64 + lane_addr;
    uint8_t* p_row = smem + OFF_P + w * 2 * TILE_BYTES + row * 128;
    uint8_t* o_stage = smem + OFF_O + w * TILE_BYTES;
    const int sw = row & 7;
    const float ebias = f16 ? 112.0f : 0.0f;   // see exp_pack_row
    const uint32_t emul = f16 ? 8u : 1u;
    const float unbias = f16 ? 1.925929944387236e-34f /* 2^-112 */ : 1.0f;
    if (pingpong && w == 1) named_bar_arrive(3, 256);  // group A holds the token first
    uint32_t blk = 0, itw = 0;
    long long t_sfull = 0, t_ld = 0, t_max = 0, t_pv = 0, t_exp = 0, t_arr = 0, t_fin = 0;
    auto tick = [&]() -> long long { return PROF ? clock64() : 0ll; };
    const long long T0 = tick();

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, h, q0;
      decode(item, b, h, q0);
      const bool hasB = q0 + 128 < N;
      if (w == 1 && !hasB) continue;  // this pair has no second query tile
      const int Nb = clip_tokens(b);
      const int nkv = (Nb + 127) / 128;
      float m_used = -INFINITY, l_run = 0.f;
      float mhist[PROBS ? PROBS_MAX_KV : 1];
      const int qrow = q0 + w * 128 + row;
      float* prow = PROBS ? probs + ((static_cast<long long>(b) * heads + h) * N + qrow) * N : nullptr;
      for (int j = 0; j < nkv; ++j, ++blk) {
        const int nvalid = min(128, Nb - j * 128);
        const long long c0 = tick();
        mbar_wait(&s_full[w], blk & 1);
        tc_fence_after();
        const long long c1 = tick();
        uint32_t s[128];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
          uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
          tmem_ld32(tmem_s, s0);
          tmem_ld32(tmem_s + 32, s1);
          tmem_ld32(tmem_s + 64, s2);
          tmem_ld32(tmem_s + 96, s3);
          tmem_ld_wait(s0);
          tmem_ld_wait(s1);
          tmem_ld_wait(s2);
          tmem_ld_wait(s3);
        }
        tc_fence_before();
        mbar_arrive(&s_free[w]);  // the MMA warp may overwrite S_w with the next block
        const long long c2 = tick();
        if (nvalid < 128) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= nvalid) s[i] = 0xFF800000u;  // -inf: exp2 -> exactly 0
        }
        // row maximum: 8 independent 3-input max chains (short dependency chains, 1 instruction per 2 scores)
        float mr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mr[i] = fmaxf(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
#pragma unroll
        for (int i = 16; i < 128; i += 16) {
#pragma unroll
          for (int k = 0; k < 8; ++k) mr[k] = max3(mr[k], __uint_as_float(s[i + 2 * k]), __uint_as_float(s[i + 2 * k + 1]));
        }
        const float mnew =
            fmaxf(max3(mr[0], mr[1], mr[2]), max3(max3(mr[3], mr[4], mr[5]), mr[6], mr[7])) * scale_log2;
        const long long c3 = tick();
        if (w == 0 && blk == 0) mbar_arrive(lag_bar);  // group B may start (half a key block behind, see the MMA issuers)
        if (j == 0) {
          m_used = mnew;
        } else {
          const bool need = mnew > m_used + RESCALE_THRESHOLD;
          mbar_wait(&pv_done[w], (blk - 1) & 1);  // O_w up to the previous block complete; P_w may be overwritten
          if (__any_sync(0xFFFFFFFFu, need)) {
            const float alpha = need ? ex2_approx(m_used - mnew) : 1.0f;
            if (need) m_used = mnew;
            l_run *= alpha;
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t r[32];
              tmem_ld32(tmem_o + c * 32, r);
              tmem_ld_wait(r);
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st32(tmem_o + c * 32, r);
            }
            tmem_st_wait();
          }
        }
        const long long c4 = tick();
        if (PROF && prof != nullptr && row == 0 && blk >= 1 && blk <= 6) prof[blockIdx.x * 32 + w * 16 + 9 + blk] = c4;
        // Exp-phase token (named barriers 3 / 4, FA3-style ping-pong): the MUFU unit of a scheduler serves one softmax
        // warp at its full rate (10 cycles per score alone, 17 each when both groups' exp phases overlap), so the groups
        // take strict turns - group A's exp phase runs while group B loads / reduces its next block and vice versa.
        // (measured: 74.9 -> 67.0 us stand-alone at 64 x 496 tokens, 397 -> 341 us at 1 248.  NON-strict variants - a
        // shared-memory lock per scheduler, per group, per group and key block; waiting by atomicCAS + __nanosleep, by
        // plain polling, or blocked on an mbarrier that completes a phase per release - all ran at 131-146 us, twice
        // slower than no lock: only the strict order keeps one group's exp phase inside the other's load / max phase)
        if (PROBS) {
          mhist[j] = m_used;
          if (qrow < N) {
            // (the same biased exponent as exp_pack_row, so exactly the terms that enter l_run - the 2^-112 is undone by
            // normalising with 1 / l_run below)
            const float mb = m_used + ebias;
            const int kmax = min(128, N - j * 128);   // keys of this block that exist in the [N, N] map
            float* dst = prow + j * 128;
            if ((N & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 128; i += 4) {
                if (i < kmax)
                  *reinterpret_cast<float4*>(dst + i) =
                      make_float4(ex2_approx(fmaf(__uint_as_float(s[i]), scale_log2, -mb)),
                                  ex2_approx(fmaf(__uint_as_float(s[i + 1]), scale_log2, -mb)),
                                  ex2_approx(fmaf(__uint_as_float(s[i + 2]), scale_log2, -mb)),
                                  ex2_approx(fmaf(__uint_as_float(s[i + 3]), scale_log2, -mb)));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 128; ++i)
                if (i < kmax) dst[i] = ex2_approx(fmaf(__uint_as_float(s[i]), scale_log2, -mb));
            }
          }
        }
        if (pingpong) named_bar_sync(3 + w, 256);
        l_run += exp_pack_row(s, scale_log2, m_used + ebias, emul, p_row, sw);
        // (group B's very last hand-over has no taker: skipped, so no barrier is left half-arrived at exit)
        if (pingpong && !(w == 1 && j + 1 == nkv && item + static_cast<int>(gridDim.x) >= n_items))
          named_bar_arrive(3 + (w ^ 1), 256);
        const long long c5 = tick();
        fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor-core async proxy
        tc_fence_before();
        mbar_arrive(&p_full[w]);
        const long long c6 = tick();
        t_sfull += c1 - c0; t_ld += c2 - c1; t_max += c3 - c2; t_pv += c4 - c3; t_exp += c5 - c4; t_arr += c6 - c5;
      }

      // ---- end of item: O_w / l -> 16-bit -> swizzled staging -> TMA store (clipped at N)
      const long long f0 = tick();
      mbar_wait(&pv_done[w], (blk - 1) & 1);
      tc_fence_after();
      const float inv = (1.0f / l_run) * unbias;
      uint32_t o0[32], o1[32];
      tmem_ld32(tmem_o, o0);
      tmem_ld32(tmem_o + 32, o1);
      tmem_ld_wait(o0);
      tmem_ld_wait(o1);
      tc_fence_before();
      mbar_arrive(&o_free[w]);  // the next item's first PV may overwrite O_w
      if (sub == 0 && lane == 0) tma_store_wait_read0();  // the previous item's store has released the staging tile
      named_bar_sync(1 + w, 128);
      uint8_t* o_row = o_stage + row * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t(&r)[32] = q < 4 ? o0 : o1;
        const int e = (q & 3) * 8;
        uint4 pk;
        pk.x = pack_16x2(__uint_as_float(r[e + 0]) * inv, __uint_as_float(r[e + 1]) * inv, f16);
        pk.y = pack_16x2(__uint_as_float(r[e + 2]) * inv, __uint_as_float(r[e + 3]) * inv, f16);
        pk.z = pack_16x2(__uint_as_float(r[e + 4]) * inv, __uint_as_float(r[e + 5]) * inv, f16);
        pk.w = pack_16x2(__uint_as_float(r[e + 6]) * inv, __uint_as_float(r[e + 7]) * inv, f16);
        *reinterpret_cast<uint4*>(o_row + ((q ^ sw) << 4)) = pk;
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + w, 128);
      if (sub == 0 && lane == 0) {
        tma_store_3d(&tmap_out, o_stage, h * 64, q0 + w * 128, b);
        tma_store_commit();
      }
      if (PROBS && qrow < N) {  // normalise this thread's row of the map: p_j * 2^(m_j - m_final) / l
        for (int j = 0; j < nkv; ++j) {
          const float f = ex2_approx(mhist[j] - m_used) * (1.0f / l_run);
          const int kmax = min(128, N - j * 128);
          float* dst = prow + j * 128;
          if ((N & 3) == 0) {
            for (int i = 0; i < kmax; i += 4) {
              float4 v = *reinterpret_cast<float4*>(dst + i);
              v.x *= f; v.y *= f; v.z *= f; v.w *= f;
              *reinterpret_cast<float4*>(dst + i) = v;
            }
          } else {
            for (int i = 0; i < kmax; ++i) dst[i] *= f;
          }
        }
      }
      ++itw;
      t_fin += tick() - f0;
    }
    if (sub == 0 && lane == 0) tma_store_wait_all();
    if (PROF && prof != nullptr && row == 0) {
      long long* pp = prof + blockIdx.x * 32 + w * 16;
      pp[0] = t_sfull; pp[1] = t_ld; pp[2] = t_max; pp[3] = t_pv; pp[4] = t_exp; pp[5] = t_arr; pp[6] = tick() - T0;
      pp[7] = itw; pp[8] = t_fin;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int launch_attn_tc(const CUtensorMap& tmap_qkv, const CUtensorMap& tmap_out, int f16, int B, int N, int heads, int D,
                   float scale, cudaStream_t stream, long long* prof, const int* geo, float* probs) {
  if (D != heads * 64) {
    set_error("attention: head_dim must be 64 (D=%d heads=%d)", D, heads);
    return -1;
  }
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) {
      once.retry();
      set_error("attn_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return -4;
    }
  }
  const int sms = num_sms();
  const long long items = static_cast<long long>((N + 255) / 256) * heads * B;
  if (items <= 0 || items > 0x7FFFFFFF) {
    set_error("attention: bad problem size B=%d N=%d heads=%d", B, N, heads);
    return -1;
  }
  const int grid = items < sms ? static_cast<int>(items) : sms;
  // exp-phase ping-pong needs both softmax groups to see the same number of key blocks: every query-tile pair must
  // have its second tile, and (variable-length batches aside, where the count is per clip but equal for both groups)
  // that is a property of N alone
  static const int pp_env = [] { const char* e = getenv("HVIT_ATTN_PINGPONG"); return e != nullptr ? atoi(e) : 1; }();
  const int pingpong = (pp_env != 0 && ((N + 255) / 256 - 1) * 256 + 128 < N) ? 1 : 0;
  if (probs != nullptr && (N + 127) / 128 > PROBS_MAX_KV) {
    set_error("attn_tc: attention maps on the tensor-core path need N <= %d (N=%d)", PROBS_MAX_KV * 128, N);
    return -1;
  }
  const float sl2 = scale * 1.4426950408889634f;
  const int ni = static_cast<int>(items);
  cudaError_t le;
  if (probs != nullptr)
    le = launch_pdl(attn_tc_kernel<false, true>, dim3(grid), dim3(AT_THREADS), AT_SMEM, stream, tmap_qkv, tmap_out, N, D,
                    heads, ni, sl2, f16, prof, geo, pingpong, probs);
  else if (prof != nullptr)
    le = launch_pdl(attn_tc_kernel<true>, dim3(grid), dim3(AT_THREADS), AT_SMEM, stream, tmap_qkv, tmap_out, N, D, heads,
                    ni, sl2, f16, prof, geo, pingpong, probs);
  else
    le = launch_pdl(attn_tc_kernel<false>, dim3(grid), dim3(AT_THREADS), AT_SMEM, stream, tmap_qkv, tmap_out, N, D, heads,
                    ni, sl2, f16, prof, geo, pingpong, probs);
  if (le != cudaSuccess) {
    set_error("attn_tc: %s", cudaGetErrorString(le));
    return -4;
  }
  return check_launch("attn_tc");
}

}  // namespace hvit
