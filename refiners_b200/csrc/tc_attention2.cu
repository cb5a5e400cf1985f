// tcgen05 flash attention, second generation: the self- / cross-attention of the SDXL transformer blocks
// (head dim <= 64, bf16 / fp16, non-causal, one K/V set, no bias, Sq > 128).  tc_attention.cu keeps every other case
// (head dim 128, SAM's relative-position bias, IP-Adapter's second K/V set, tiny Sq).
//
//   O[b, q, h, :] = softmax(Q K^T * scale) V        replaces fluxion/layers/attentions.py:115-202 of the reference
//
// What changed against the first kernel, and why (ncu of round 1: tensor pipe 22 %, XU 34 %, issue 38 % - a latency chain):
//   * ONE persistent CTA per SM owns a PAIR of 128-query tiles of one (batch, head): both tiles share every K/V tile
//     that TMA brings in (half the L2 -> smem traffic per flop: two CTAs per SM each streaming their own K/V would need
//     ~15 TB/s of L2 bandwidth at 1 PFLOP/s).
//   * 128 keys per tile: S = Q K^T is a UMMA 128 x 128 x 16 (full-rate instruction shape), P V a 128 x 64 x 16 over 8 k-steps.
//   * FOUR softmax warpgroups (two per query tile, each owning 64 of the tile's 128 key columns) against one MMA-issuing
//     thread: four warps per scheduler keep the MUFU pipe fed (a single warp per scheduler reaches ~50 % of the pipe's
//     rate - measured with ncu - because its own FADD / F2FP / FFMA stream and the MUFU latencies serialise with its
//     MUFU issue), and while the groups of tile A exponentiate the tensor core computes S_B / P_B V, and vice versa.
//   * the running output never leaves TMEM: the P V MMAs accumulate into one O tile per group, and the softmax keeps a
//     STALE running maximum - O is rescaled in TMEM (tcgen05.ld / st) only when a row's maximum grows by more than 2^8
//     (FlashAttention-4's conditional rescaling); the first kernel read 64 fp32 columns of O_j back per thread per tile.
//   * S is read from TMEM exactly once per tile into registers (64 fp32 per thread); the two halves of a row exchange
//     their partial maxima through shared memory (double buffered, one named barrier per tile) and keep partial row sums
//     that are only combined when the row is written out.
//   * optionally (POLY) every fourth exponential runs as a Cody-Waite cubic on the FMA pipe instead of MUFU.EX2: at head
//     dim 64 the exponentials, not the MMAs, bound the kernel (16 MUFU lanes / clk / SM against 4096 MAC / clk / SM).
//
// Warps: warpgroup 0 = {warp 0: TMA producer, warp 1: MMA issuer of tile A + TMEM owner, warp 2: MMA issuer of tile B, warp 3: idle}; warpgroups 1, 2 = softmax
// of query tile A (key columns 0-63 / 64-127); warpgroups 3, 4 = softmax of query tile B.  thread = (query row = TMEM
// lane, column half).
// TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384).
// smem: Q 2 x (2 x 16 KB) (double buffered across work items), K/V ring 2 x (16 + 16 KB), P_A, P_B 32 KB each, 4 KB exchange.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rb200 {
namespace {

using namespace ptx;

constexpr int QT = 128;             // queries per tile (UMMA M)
constexpr int KT = 128;             // keys per tile (UMMA N of S, K extent of P V)
constexpr int HD = 64;              // head-dim slab (columns >= D are zero filled by TMA)
constexpr int STAGES = 2;  // K/V ring: a tile (32 KB) is in use for ~2 us, TMA needs ~1 us to refill its slot
constexpr int NUM_THREADS = 640;
constexpr int Q_TILE_BYTES = QT * HD * 2;      // 16 KB
constexpr int Q_BYTES = 2 * Q_TILE_BYTES;      // both tiles of a pair
constexpr int K_BYTES = KT * HD * 2, V_BYTES = KT * HD * 2;
constexpr int P_SLAB = QT * 64 * 2;            // 64 keys of P for 128 rows: one 128-byte-row swizzle slab
constexpr int P_BYTES = 2 * P_SLAB;            // 128 keys
constexpr int TMEM_COLS = 512;
constexpr int XCHG_BYTES = 2 /*groups*/ * 2 /*tile parity*/ * 2 /*halves*/ * QT * 4;   // partial row maxima (also reused for the row sums)
constexpr size_t SMEM_BYTES = 2 * Q_BYTES + STAGES * (K_BYTES + V_BYTES) + 2 * P_BYTES + 1024 /*align*/ + 256 /*barriers*/ + XCHG_BYTES;
static_assert(SMEM_BYTES <= 227 * 1024, "exceeds the 227 KB of shared memory a CTA can opt in to");
constexpr float RESCALE_LOG2 = 8.0f;           // tolerate a stale maximum until a probability could exceed 2^8

struct Attn2Params {
  void* o;
  int64_t o_sb, o_ss;
  int H;
  int64_t Sq, Sk;
  int n_pairs;          // ceil(Sq / 256)
  int64_t total_work;   // B * H * n_pairs
  int ntiles;           // ceil(Sk / 128)
  float scale_log2e;
  uint32_t idesc_qk, idesc_pv;
  int d_out;
  int stagger;          // RB200_ATTN_STAGGER: 2 (default) = the two tiles' exponential phases alternate strictly (turnstile on
                        // named barriers 3 / 4), 1 = only tile B's first phase of a work item waits for tile A's, 0 = free running
};

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <typename T, int POLY>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_sdpa2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                           // [2 buffers][2 tiles][128 x 64]
  uint8_t* sK = sQ + 2 * Q_BYTES;               // [STAGES][128 x 64]
  uint8_t* sV = sK + STAGES * K_BYTES;
  uint8_t* sP = sV + STAGES * V_BYTES;          // [2 groups][2 slabs][128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
  uint64_t* kv_full = bars;                     // [STAGES] TMA -> MMA
  uint64_t* kv_empty = bars + STAGES;           // [STAGES] MMA -> TMA
  uint64_t* q_full = bars + 2 * STAGES;         // [2]
  uint64_t* q_empty = q_full + 2;               // [2]
  uint64_t* bar_s = q_full + 4;                 // [2 groups] S tile of the group is in TMEM
  uint64_t* bar_sfree = q_full + 6;             // [2] the group has read its S tile into registers
  uint64_t* bar_p = q_full + 8;                 // [2] P tile of the group is in smem (and O has been rescaled if needed)
  uint64_t* bar_o = q_full + 10;                // [2] every P V issued so far for the group has landed in O
  uint64_t* bar_ofree = q_full + 12;            // [2] the finished work item's O rows have been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 14);
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [group][parity][half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = warp >> 2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_k);
    prefetch_tmap(&map_v);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);   // both MMA issuers release a K/V stage
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 2);
      mbar_init(&bar_s[b], 1);
      mbar_init(&bar_sfree[b], 8);   // 2 column halves x 4 warps
      mbar_init(&bar_p[b], 8);
      mbar_init(&bar_o[b], 1);
      mbar_init(&bar_ofree[b], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (wg == 0) {
    setmaxnreg_dec<56>();
    if (warp == 0) {
      // ================================================================================ TMA
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        uint32_t n = 0;  // work items of this CTA so far: Q buffer = n & 1, its phase = (n >> 1) & 1
        for (int64_t w = blockIdx.x; w < p.total_work; w += gridDim.x, ++n) {
          const int pair = int(w % p.n_pairs);
          const int h = int((w / p.n_pairs) % p.H);
          const int b = int(w / (int64_t(p.n_pairs) * p.H));
          const uint32_t qb = n & 1, qph = (n >> 1) & 1;
          mbar_wait(&q_empty[qb], qph ^ 1, 1);
          mbar_arrive_expect_tx(&q_full[qb], Q_BYTES);
          tma_load_4d(sQ + qb * Q_BYTES, &map_q, &q_full[qb], 0, h, pair * 2 * QT, b);
          tma_load_4d(sQ + qb * Q_BYTES + Q_TILE_BYTES, &map_q, &q_full[qb], 0, h, pair * 2 * QT + QT, b);
          for (int j = 0; j < p.ntiles; ++j) {
            mbar_wait(&kv_empty[stage], phase ^ 1, 2);
            mbar_arrive_expect_tx(&kv_full[stage], K_BYTES + V_BYTES);
            tma_load_4d(sK + stage * K_BYTES, &map_k, &kv_full[stage], 0, h, j * KT, b);
            tma_load_4d(sV + stage * V_BYTES, &map_v, &kv_full[stage], 0, h, j * KT, b);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ================================================================================ MMA
      // One issuing thread PER query tile (warp 1: tile A, warp 2: tile B).  Everything an issuer does is a chain of
      // latencies - mbarrier try_wait (~90 cycles even when the phase has completed), tcgen05 fence, descriptor set-up,
      // MMA issue, commit - about 2000 cycles per key tile when a single thread served both tiles: with one thread the
      // softmax warps spent 27 % of their time waiting for S tiles and P V completions while the tensor pipe was 21 %
      // busy (ncu, profiles/r02_ncu_attn_*).  tcgen05.commit tracks the MMAs of the committing thread only, so the two
      // issuers need no coordination beyond the K/V and Q slots, which are released when BOTH have committed (count 2).
      if (lane == 0) {
        const int g = warp - 1;
        int st_k = 0, st_v = 0;     // ring cursors: stage whose K feeds the next S tile / whose V feeds the next P V
        uint32_t ph_k = 0;
        uint32_t t = 0;             // P V tiles issued so far (phases of bar_p / bar_o)
        uint32_t s_issued = 0;      // S tiles issued so far (runs one ahead of t)
        uint32_t n = 0;
        const uint32_t tmem_s = tmem_base + g * 128, tmem_o = tmem_base + 256 + g * 64;
        const uint32_t pbase = smem_u32(sP + g * P_BYTES);
        // S_g(tile) = Q_g K^T: the group must have read its previous S tile out of TMEM (it releases the tile as soon as the
        // values sit in its registers, early in its work on that tile)
        auto issue_s = [&](uint32_t qb, int stage_k) {
          if (s_issued > 0) mbar_wait(&bar_sfree[g], (s_issued - 1) & 1, 4);
          ++s_issued;
          tcgen05_fence_after();
          const uint64_t dq = desc_kmajor(smem_u32(sQ + qb * Q_BYTES + g * Q_TILE_BYTES));
          const uint64_t dk = desc_kmajor(smem_u32(sK + stage_k * K_BYTES));
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_f16(tmem_s, dq + uint64_t(k * 2), dk + uint64_t(k * 2), p.idesc_qk, k > 0);
          umma_commit(&bar_s[g]);
        };
        auto next_k = [&]() {
          if (++st_k == STAGES) {
            st_k = 0;
            ph_k ^= 1;
          }
        };
        for (int64_t w = blockIdx.x; w < p.total_work; w += gridDim.x, ++n) {
          const uint32_t qb = n & 1, qph = (n >> 1) & 1;
          mbar_wait(&q_full[qb], qph, 5);
          mbar_wait(&kv_full[st_k], ph_k, 3);
          issue_s(qb, st_k);
          next_k();
          for (int j = 0; j < p.ntiles; ++j) {
            // S(j + 1) FIRST: it only needs the S tile to be free, which happens long before P(j) is ready - when the
            // group comes back for tile j + 1 the logits are already waiting in TMEM
            if (j + 1 < p.ntiles) {
              mbar_wait(&kv_full[st_k], ph_k, 6);  // K_{j+1}
              issue_s(qb, st_k);
              next_k();
            }
            mbar_wait(&bar_p[g], t & 1, 7);  // P(j) is in smem; O has been rescaled if it had to be
            if (j == 0 && n > 0) mbar_wait(&bar_ofree[g], (n - 1) & 1, 8);  // O still holds the previous work item until read out
            tcgen05_fence_after();
            const uint32_t vbase = smem_u32(sV + st_v * V_BYTES);
#pragma unroll
            for (int k = 0; k < KT / 16; ++k) {
              // A = P (K-major, two 64-key slabs; +32 B per 16 keys inside a swizzle row); B = V (MN-major: +16 key rows * 128 B)
              const uint64_t dp = desc_kmajor(pbase + (k >> 2) * P_SLAB) + uint64_t((k & 3) * 2);
              const uint64_t dv = desc_mnmajor(vbase, V_BYTES) + uint64_t(k * 128);
              umma_f16(tmem_o, dp, dv, p.idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&kv_empty[st_v]);  // this issuer's S (issued earlier) and P V have consumed the stage; count 2
            umma_commit(&bar_o[g]);
            ++t;
            if (++st_v == STAGES) st_v = 0;
          }
          umma_commit(&q_empty[qb]);  // every MMA of this issuer reading the Q buffer has been issued; count 2
        }
      }
    }
  } else {
    // ============================================================================ softmax
    setmaxnreg_inc<104>();
    const int g = (wg - 1) >> 1;          // query tile of the pair
    const int hf = (wg - 1) & 1;          // which 64 key columns of the tile (and which 32 output columns) this thread owns
    const int lg = warp & 3;              // TMEM lane group of this warp
    const int row = lg * 32 + lane;
    const uint32_t lane_off = uint32_t(lg * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * 128 + hf * 64 + lane_off;
    const uint32_t tmem_o = tmem_base + 256 + g * 64 + hf * 32 + lane_off;
    const uint32_t prow = smem_u32(sP + g * P_BYTES + hf * P_SLAB + row * 128);
    const int sw = row & 7;
    // shared-window addresses taken once (see tc_ptx.cuh: each smem_u32 of a generic pointer is ~6 uniform instructions)
    const uint32_t xg = smem_u32(xchg + g * (2 * 2 * QT));  // this group's exchange area: [parity][half][row] floats
    const uint32_t x_mine = xg + uint32_t(hf * QT + row) * 4, x_other = xg + uint32_t((hf ^ 1) * QT + row) * 4;
    const uint32_t a_s = smem_u32(&bar_s[g]), a_sfree = smem_u32(&bar_sfree[g]), a_p = smem_u32(&bar_p[g]);
    const uint32_t a_o = smem_u32(&bar_o[g]), a_ofree = smem_u32(&bar_ofree[g]);
    const uint32_t swz = uint32_t(sw) << 4;
    const uint32_t xbar = 1 + g;          // named barrier of the group's 256 threads
    T* obase = static_cast<T*>(p.o);
    uint32_t t = 0;                       // tiles of this group so far
    uint32_t n = 0;
    for (int64_t w = blockIdx.x; w < p.total_work; w += gridDim.x, ++n) {
      const int pair = int(w % p.n_pairs);
      const int h = int((w / p.n_pairs) % p.H);
      const int64_t b = w / (int64_t(p.n_pairs) * p.H);
      float m_run = -INFINITY, l_run = 0.f;  // l_run: this half's share of the row sum
      for (int j = 0; j < p.ntiles; ++j, ++t) {
        mbar_wait_a(a_s, t & 1);
        tcgen05_fence_after();
        float s[KT / 2];
        {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(tmem_s, r0);
          tmem_ld_32x32(tmem_s + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s[i] = __uint_as_float(r0[i]);
            s[32 + i] = __uint_as_float(r1[i]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed_a(a_sfree);  // the tensor core may overwrite this S tile
        const int64_t left = p.Sk - int64_t(j) * KT - hf * 64;  // valid keys among this thread's 64 columns
        if (left < 64) {
          const int valid = left > 0 ? int(left) : 0;
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= valid) s[i] = -INFINITY;
        }
        float tm0 = fmaxf(s[0], s[1]), tm1 = fmaxf(s[2], s[3]), tm2 = fmaxf(s[4], s[5]), tm3 = fmaxf(s[6], s[7]);
#pragma unroll
        for (int i = 8; i < 64; i += 8) {  // four independent chains: the row maximum is on every tile's critical path
          tm0 = fmaxf(tm0, fmaxf(s[i], s[i + 1]));
          tm1 = fmaxf(tm1, fmaxf(s[i + 2], s[i + 3]));
          tm2 = fmaxf(tm2, fmaxf(s[i + 4], s[i + 5]));
          tm3 = fmaxf(tm3, fmaxf(s[i + 6], s[i + 7]));
        }
        // the row maximum of the tile = max over both column halves: exchange through smem (slot parity t & 1: a thread is
        // never more than one tile ahead of its partner, which may still be reading the previous tile's slot)
        const uint32_t slot = (t & 1) * (2 * QT * 4);
        const float mine = fmaxf(fmaxf(tm0, tm1), fmaxf(tm2, tm3));
        st_shared_f32(x_mine + slot, mine);
        named_bar_sync(xbar, 256);
        const float tmax = fmaxf(mine, ld_shared_f32(x_other + slot));
        bool waited_o = false;
        if (j == 0) {
          m_run = tmax;  // nothing accumulated yet
        } else {
          const float m_new = fmaxf(m_run, tmax);
          const bool grew = (m_new - m_run) * p.scale_log2e > RESCALE_LOG2;
          // TMEM access is warp-collective: the whole warp rescales its 32 rows; the partner warp (same rows, other column
          // half) sees the same maxima and takes the same decision for its 32 output columns
          if (__any_sync(0xffffffffu, grew)) {
            mbar_wait_a(a_o, (t - 1) & 1);  // every P V issued so far has landed in O
            waited_o = true;
            tcgen05_fence_after();
            const float alpha = ex2_approx((m_run - m_new) * p.scale_log2e);
            uint32_t raw[32];
            tmem_ld_32x32(tmem_o, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
            tmem_st_32x32(tmem_o, raw);
            tmem_st_wait();
            l_run *= alpha;
            m_run = m_new;
          }
        }
        const float mb = m_run * p.scale_log2e;
        // the P buffer of this group was last read by P V of tile t - 1 (issued a whole tile ago): wait for it BEFORE the
        // exponentials, so that every 16-byte chunk of P goes to shared memory as soon as it exists (4 live registers
        // instead of 32)
        if ((j > 0 && !waited_o) || (j == 0 && t > 0)) mbar_wait_a(a_o, (t - 1) & 1);  // (j == 0: the previous work item's last tile)
        // Turnstile: the two query tiles take turns in the exponential phase, so that the XU of a scheduler is contended by
        // two warps instead of four and the other tile's TMEM loads / maxima / barriers run underneath (measured +4.5 % at
        // S = 1024, +8 % at S = 4096; RB200_ATTN_STAGGER=0 switches it off, 1 = only the first tile of a work item)
        if (p.stagger == 1) {
          if (g == 1 && j == 0) named_bar_sync(3, 512);
        } else if (p.stagger == 2) {
          if (g == 1) named_bar_sync(3, 512);
          else if (t > 0) named_bar_sync(4, 512);
        }
        // x = s * scale - m and the running sum as packed fp32 pairs (FFMA2 / FADD2: half the issue slots)
        const uint64_t sc2 = f32x2(p.scale_log2e, p.scale_log2e), nmb2 = f32x2(-mb, -mb);
        uint64_t sum_a = f32x2(0.f, 0.f), sum_b = sum_a;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int i = c * 8 + q * 4;
            float x0, x1, x2, x3;
            f32x2_split(fma_f32x2(f32x2(s[i], s[i + 1]), sc2, nmb2), x0, x1);
            f32x2_split(fma_f32x2(f32x2(s[i + 2], s[i + 3]), sc2, nmb2), x2, x3);
            const float p0 = ex2_approx(x0), p2 = ex2_approx(x2);
            const float p1 = POLY >= 2 ? ex2_poly(x1) : ex2_approx(x1);
            const float p3 = POLY >= 1 ? ex2_poly(x3) : ex2_approx(x3);
            sum_a = add_f32x2(sum_a, f32x2(p0, p1));
            sum_b = add_f32x2(sum_b, f32x2(p2, p3));
            pk[q * 2] = pack2<T>(p0, p1);
            pk[q * 2 + 1] = pack2<T>(p2, p3);
          }
          st_shared_v4(prow + ((uint32_t(c) << 4) ^ swz), pk[0], pk[1], pk[2], pk[3]);
        }
        {
          float a0, a1;
          f32x2_split(add_f32x2(sum_a, sum_b), a0, a1);
          l_run += a0 + a1;
        }
        if (p.stagger == 1) {
          if (g == 0 && j == 0) named_bar_arrive(3, 512);
        } else if (p.stagger == 2) {
          named_bar_arrive(g == 0 ? 3 : 4, 512);
        }
        fence_proxy_async();      // generic-proxy writes -> visible to the tensor core (async proxy)
        tcgen05_fence_before();   // also orders a rescale's tcgen05.st before the MMA that the arrive releases
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_p);
      }
      // ---- row sum = both halves' shares; read the finished rows out of TMEM (32 output columns per thread)
      const uint32_t slot = (t & 1) * (2 * QT * 4);   // parity of the NEXT tile: free (its last use was two tiles ago)
      st_shared_f32(x_mine + slot, l_run);
      named_bar_sync(xbar, 256);
      const float l_row = l_run + ld_shared_f32(x_other + slot);
      named_bar_sync(xbar, 256);               // both halves have read before the next work item's first tile rewrites the slot
      mbar_wait_a(a_o, (t - 1) & 1);
      tcgen05_fence_after();
      const float inv = l_row > 0.f ? 1.f / l_row : 0.f;
      float acc[32];
      {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_o, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = inv * __uint_as_float(raw[i]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed_a(a_ofree);  // the MMA warp may start the next work item's P V
      const int64_t qi = (int64_t(pair) * 2 + g) * QT + row;
      if (qi < p.Sq) {
        T* dst = obase + b * p.o_sb + qi * p.o_ss + int64_t(h) * p.d_out + hf * 32;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (p.d_out & 7) == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (hf * 32 + c * 8 >= p.d_out) break;
            uint4 v;
            v.x = pack2<T>(acc[c * 8], acc[c * 8 + 1]);
            v.y = pack2<T>(acc[c * 8 + 2], acc[c * 8 + 3]);
            v.z = pack2<T>(acc[c * 8 + 4], acc[c * 8 + 5]);
            v.w = pack2<T>(acc[c * 8 + 6], acc[c * 8 + 7]);
            reinterpret_cast<uint4*>(dst)[c] = v;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (hf * 32 + i < p.d_out) dst[i] = from_f<T>(acc[i]);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return EncodeTiledFn(nullptr);
    return reinterpret_cast<EncodeTiledFn>(sym);
  }();
  return fn;
}

// [B, S, H, D] view with strides (sb, ss, D, 1) elements; box = 64 x 1 x 128 rows x 1 (columns >= D are zero filled)
int make_map(CUtensorMap* map, int dtype, const void* base, int64_t B, int64_t S, int H, int64_t sb, int64_t ss, int D) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) RB200_FAIL(-4, "cuTensorMapEncodeTiled unavailable");
  const cuuint64_t dims[4] = {cuuint64_t(D), cuuint64_t(H), cuuint64_t(S), cuuint64_t(B)};
  const cuuint64_t strides[3] = {cuuint64_t(D) * 2, cuuint64_t(ss) * 2, cuuint64_t(sb) * 2};
  const cuuint32_t box[4] = {64, 1, 128, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = dtype == RB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult rc = fn(map, dt, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) RB200_FAIL(-4, "sdpa2 tensor map encode failed (%d): S=%lld H=%d ss=%lld sb=%lld", int(rc), (long long)S, H, (long long)ss, (long long)sb);
  return 0;
}

bool ok_operand(const void* ptr, int64_t sb, int64_t ss) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ss % 8 == 0 && sb % 8 == 0;
}

template <typename T, int POLY>
int launch(cudaStream_t st, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const Attn2Params& prm) {
  static PerDeviceOnce configured;
  if (configured.needed()) {
    if (cudaFuncSetAttribute(tc_sdpa2_kernel<T, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)) != cudaSuccess)
      RB200_FAIL(-2, "tc_sdpa2: cannot reserve %zu bytes of shared memory", SMEM_BYTES);
    configured.done();
  }
  const int64_t cap = sm_count();
  const int grid = int(prm.total_work < cap ? prm.total_work : cap);
  tc_sdpa2_kernel<T, POLY><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mq, mk, mv, prm);
  RB200_CHECK_LAUNCH("tc_sdpa2");
  return 0;
}

int env_int(const char* name, int fallback) {
  const char* e = getenv(name);
  return e ? atoi(e) : fallback;
}

}  // namespace

// RB200_ATTN_V2: 0 = always the first-generation kernel, 1 (default) = this kernel where it applies.
// RB200_ATTN_POLY: 1 / 2 = every fourth / every second exponential on the FMA pipe, 0 (default) = all on MUFU (measured: no gain with two
// softmax warps per scheduler).
bool tc_sdpa2_supported(const SdpaProblem& p) {
  static const int enabled = env_int("RB200_ATTN_V2", 1);
  if (!enabled) return false;
  if (p.dtype != RB200_BF16 && p.dtype != RB200_FP16) return false;
  if (p.D < 8 || p.D > 64 || (p.D & 7) != 0 || p.causal) return false;
  if (p.bias_h != nullptr || (p.k2 != nullptr && p.Sk2 > 0)) return false;
  // one key tile per work item (the 77 text tokens of cross-attention): nothing to pipeline across, and the first-generation
  // kernel's two small CTAs per SM hide each other's fill / drain better (49 vs 57 us on [16, 20, 1024] x 77, measured)
  if (p.Sq <= QT || p.Sk <= KT || p.B < 1) return false;
  return ok_operand(p.q, p.q_sb, p.q_ss) && ok_operand(p.k, p.k_sb, p.k_ss) && ok_operand(p.v, p.v_sb, p.v_ss);
}

int tc_sdpa2(cudaStream_t st, const SdpaProblem& p) {
  CUtensorMap mq, mk, mv;
  if (int rc = make_map(&mq, p.dtype, p.q, p.B, p.Sq, p.H, p.q_sb, p.q_ss, p.D)) return rc;
  if (int rc = make_map(&mk, p.dtype, p.k, p.B, p.Sk, p.H, p.k_sb, p.k_ss, p.D)) return rc;
  if (int rc = make_map(&mv, p.dtype, p.v, p.B, p.Sk, p.H, p.v_sb, p.v_ss, p.D)) return rc;
  Attn2Params prm{};
  prm.o = p.o;
  prm.o_sb = p.o_sb;
  prm.o_ss = p.o_ss;
  prm.H = p.H;
  prm.Sq = p.Sq;
  prm.Sk = p.Sk;
  prm.n_pairs = int(ceil_div(p.Sq, 2 * QT));
  prm.total_work = int64_t(prm.n_pairs) * p.H * p.B;
  prm.ntiles = int(ceil_div(p.Sk, KT));
  prm.scale_log2e = p.scale * 1.4426950408889634f;
  const uint32_t fmt = p.dtype == RB200_BF16 ? 1u : 0u;
  const uint32_t common = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(QT >> 4) << 24);
  prm.idesc_qk = common | (uint32_t(KT >> 3) << 17);                // D = 128 x 128, A and B K-major
  prm.idesc_pv = common | (uint32_t(HD >> 3) << 17) | (1u << 16);   // D = 128 x 64, B (= V) MN-major
  prm.d_out = p.D;
  static const int stagger = env_int("RB200_ATTN_STAGGER", 2);
  prm.stagger = stagger;
  static const int poly = env_int("RB200_ATTN_POLY", 0);
  const bool bf = p.dtype == RB200_BF16;
  if (poly >= 2) return bf ? launch<__nv_bfloat16, 2>(st, mq, mk, mv, prm) : launch<__half, 2>(st, mq, mk, mv, prm);
  if (poly == 1) return bf ? launch<__nv_bfloat16, 1>(st, mq, mk, mv, prm) : launch<__half, 1>(st, mq, mk, mv, prm);
  return bf ? launch<__nv_bfloat16, 0>(st, mq, mk, mv, prm) : launch<__half, 0>(st, mq, mk, mv, prm);
}

}  // namespace rb200
