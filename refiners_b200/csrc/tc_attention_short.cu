// Single-pass tensor-core attention for SHORT key sequences (Sk <= 128, head dim <= 64): the text cross-attention of the
// UNets (77 tokens; 70 launches per SDXL step).
//
//   O[b, q, h, :] = softmax(Q K^T * scale) V        replaces fluxion/layers/attentions.py:115-202 of the reference
//
// With at most 128 keys the whole logit row of a query fits in one TMEM tile: no online softmax, no rescaling, no key loop.
// What is left is a STREAM: per 128-query tile 16 KB of Q come in and 16 KB of O go out while K and V (<= 16 KB each, the
// same for every query tile of a (batch, head)) come from L2 - the op is HBM-bound (84 MB per SDXL launch, 13 us at the
// measured copy bandwidth), and the flash kernels, built around a key loop, spent ~5 us of dependent latencies per tile on
// it.  Here the tiles of a CTA are software-pipelined:
//
//   warp 0      TMA producer: {Q tile, K, V} of item i into the next stage of the ring
//   warp 1      MMA issuer (one thread): S_g = Q K^T (128 x Sk_pad x 64), later O_g = P_g V (128 x 64 x Sk_pad); g = i & 1;
//               S tiles run two items ahead of the P V products, so neither the tensor core nor a softmax group waits
//   warps 2-5   softmax group 0 (items 0, 2, 4 ...): thread = query row = TMEM lane; logits -> registers, max, exp2, sum,
//   warps 6-9   softmax group 1 (items 1, 3, 5 ...)  P -> bf16 -> swizzled smem; the group's PREVIOUS output is read from TMEM,
//               scaled by 1 / sum and stored between the exponentials and the P store of the current item
//
// TMEM: S_0, S_1 (128 columns each), O_0, O_1 (64 each).  K and V tiles hold ceil(Sk / 16) * 16 rows (rows >= Sk zero-filled by
// TMA); logit columns >= Sk - computed from those rows, or never written - are set to -inf, so their probabilities are exactly 0.
// The {Q, K, V} stages are sized by Sk: four fit for the 77 text tokens (36 KB each), three at 128 keys.
//
// DUAL instantiations add the IP-Adapter term (latent_diffusion/image_prompt.py:237-309 of the reference):
//   O += scale2 * softmax(Q K2^T * scale) V2        with at most 32 extra keys (4 image tokens, 16 for the "plus" models)
// K2 / V2 ride in the same stage, S2 gets 32 TMEM columns of its own and its own softmax; its probabilities are pre-scaled by
// scale2 * l / l2 and written into the P buffer right behind P's key steps, so that ONE accumulator and ONE 1 / l in the
// epilogue serve both softmaxes.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rb200 {
namespace {

using namespace ptx;

constexpr int QT = 128;
constexpr int HD = 64;
constexpr int MAX_NST = 4;
constexpr int THREADS = 320;
constexpr int TILE_BYTES = QT * HD * 2;         // 16 KB: a Q tile (a K or V tile holds kv_rows <= 128 rows of 128 B)
constexpr int RING_BYTES = 144 * 1024;          // stages of {Q | K | V}: 4 x 36 KB up to 80 keys, 3 x 48 KB up to 128
constexpr int P_SLAB = QT * 64 * 2;             // 64 keys of P for 128 rows
constexpr int P_BYTES = 2 * P_SLAB;
constexpr int TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = RING_BYTES + 2 * P_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 227 * 1024, "exceeds the 227 KB of shared memory a CTA can opt in to");

struct ShortParams {
  void* o;
  int64_t o_sb, o_ss;
  int H;
  int64_t Sq, Sk;
  int n_qt;             // ceil(Sq / 128)
  int64_t total_work;   // B * H * n_qt
  int ksteps;           // ceil(Sk / 16): K extent of P V in MMA steps
  int kv_bytes;         // a K (or V) tile: 16 * ksteps rows of 128 B (the TMA box of the K / V maps)
  int stage_bytes;      // TILE_BYTES + 2 * kv_bytes
  int nst;              // stages that fit the ring: 3 or 4
  // second key / value set (IP-Adapter, DUAL instantiations): o += scale2 * softmax(q k2^T * scale) v2
  int64_t Sk2;
  int ksteps2;          // ceil(Sk2 / 16) <= 2
  int kv2_bytes;        // 16 * ksteps2 rows of 128 B
  float scale2;
  uint32_t idesc_qk2;
  float scale_log2e;
  uint32_t idesc_qk, idesc_pv;
  int d_out;
};

template <typename T> __device__ __forceinline__ uint32_t pack2s(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2s<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2s<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// NC = 32-column chunks of logits per row (Sk <= 32 * NC).  One thread per query row: a variant with two threads per row (16
// softmax warps, row maximum and sum exchanged through shared memory and a named barrier) measured SLOWER - 38.9 vs 32.8 us at
// Sq = 1024, 69.6 vs 53.2 us at Sq = 4096 (session V): the two exchanges per item cost more than the extra warps hide.
template <typename T, int NC, bool DUAL>
__global__ void __launch_bounds__(THREADS, 1)
tc_sdpa_short_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_k2,
                     const __grid_constant__ CUtensorMap map_v2, const ShortParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStage = smem;                                  // [nst][Q | K | V]
  uint8_t* sP = sStage + RING_BYTES;                       // [2 groups][2 slabs][128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
  uint64_t* full = bars;               // [nst] TMA -> MMA
  uint64_t* empty = bars + MAX_NST;    // [nst] MMA -> TMA
  uint64_t* bar_s = bars + 2 * MAX_NST;  // [2] S_g is in TMEM
  uint64_t* bar_sfree = bar_s + 2;     // [2] group g holds its logits in registers
  uint64_t* bar_p = bar_s + 4;         // [2] P_g is in smem
  uint64_t* bar_o = bar_s + 6;         // [2] O_g is in TMEM
  uint64_t* bar_ofree = bar_s + 8;     // [2] group g has read O_g out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_k);
    prefetch_tmap(&map_v);
    for (int s = 0; s < MAX_NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bar_s[g], 1);
      mbar_init(&bar_sfree[g], 4);
      mbar_init(&bar_p[g], 4);
      mbar_init(&bar_o[g], 1);
      mbar_init(&bar_ofree[g], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // items of this CTA: w = blockIdx.x, blockIdx.x + gridDim.x, ... (total_work < 2^31: checked on the host)
  const uint32_t total = uint32_t(p.total_work), n_qt = uint32_t(p.n_qt), heads = uint32_t(p.H);
  const uint32_t n_items = total > blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t nst = uint32_t(p.nst);

  auto producer = [&]() {
    // ==================================================================================== TMA
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (uint32_t i = 0; i < n_items; ++i) {
        const uint32_t w = blockIdx.x + i * gridDim.x;
        const int qt = int(w % n_qt), h = int((w / n_qt) % heads), b = int(w / (n_qt * heads));
        mbar_wait(&empty[st], ph ^ 1, 1);
        mbar_arrive_expect_tx(&full[st], uint32_t(p.stage_bytes));
        uint8_t* dst = sStage + st * p.stage_bytes;
        tma_load_4d(dst, &map_q, &full[st], 0, h, qt * QT, b);
        tma_load_4d(dst + TILE_BYTES, &map_k, &full[st], 0, h, 0, b);
        tma_load_4d(dst + TILE_BYTES + p.kv_bytes, &map_v, &full[st], 0, h, 0, b);
        if constexpr (DUAL) {
          tma_load_4d(dst + TILE_BYTES + 2 * p.kv_bytes, &map_k2, &full[st], 0, h, 0, b);
          tma_load_4d(dst + TILE_BYTES + 2 * p.kv_bytes + p.kv2_bytes, &map_v2, &full[st], 0, h, 0, b);
        }
        if (++st == nst) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  };
  auto issuer = [&]() {
    // ==================================================================================== MMA
    if (lane == 0) {
      // S tiles run TWO items ahead of the P V products: S(i + 2) only needs group (i & 1) to have moved the logits of
      // item i into its registers, which happens at the very start of its work on item i - when the group comes back for
      // its next item the logits are already waiting (with S one item ahead the softmax warps spent 16 % of their time
      // waiting for them: ncu, session N)
      uint32_t s_st = 0, s_ph = 0;     // ring cursor of the next S tile
      auto issue_s = [&](uint32_t i) {
        const uint32_t g = i & 1, k = i >> 1;  // k-th item of group g
        mbar_wait(&full[s_st], s_ph, 2);
        if (k > 0) mbar_wait(&bar_sfree[g], (k - 1) & 1, 3);
        tcgen05_fence_after();
        const uint32_t base = smem_u32(sStage + s_st * p.stage_bytes);
        const uint64_t dq = desc_kmajor(base), dk = desc_kmajor(base + TILE_BYTES);
#pragma unroll
        for (int s = 0; s < HD / 16; ++s) umma_f16(tmem_base + g * 128, dq + uint64_t(s * 2), dk + uint64_t(s * 2), p.idesc_qk, s > 0);
        if constexpr (DUAL) {  // S2_g = Q K2^T into its own 32 columns
          const uint64_t dk2 = desc_kmajor(base + TILE_BYTES + 2 * p.kv_bytes);
#pragma unroll
          for (int s = 0; s < HD / 16; ++s) umma_f16(tmem_base + 384 + g * 32, dq + uint64_t(s * 2), dk2 + uint64_t(s * 2), p.idesc_qk2, s > 0);
        }
        umma_commit(&bar_s[g]);
        if (++s_st == nst) {
          s_st = 0;
          s_ph ^= 1;
        }
      };
      if (n_items > 0) issue_s(0);
      if (n_items > 1) issue_s(1);
      uint32_t st = 0;
      for (uint32_t i = 0; i < n_items; ++i) {
        if (i + 2 < n_items) issue_s(i + 2);
        const uint32_t g = i & 1, k = i >> 1;
        mbar_wait(&bar_p[g], k & 1, 4);
        if (k > 0) mbar_wait(&bar_ofree[g], (k - 1) & 1, 5);
        tcgen05_fence_after();
        const uint32_t pbase = smem_u32(sP + g * P_BYTES);
        const uint32_t vbase = smem_u32(sStage + st * p.stage_bytes + TILE_BYTES + p.kv_bytes);
        for (int s = 0; s < p.ksteps; ++s) {
          // A = P (K-major, two 64-key slabs; +32 B per 16 keys inside a swizzle row); B = V (MN-major: +16 key rows * 128 B)
          const uint64_t dp = desc_kmajor(pbase + (s >> 2) * P_SLAB) + uint64_t((s & 3) * 2);
          const uint64_t dv = desc_mnmajor(vbase, uint32_t(p.kv_bytes)) + uint64_t(s * 128);
          umma_f16(tmem_base + 256 + g * 64, dp, dv, p.idesc_pv, s > 0 ? 1u : 0u);
        }
        if constexpr (DUAL) {  // += P2' V2: P2' sits in the P buffer right behind the key steps of P
          const uint32_t v2base = vbase + p.kv_bytes + p.kv2_bytes;
          for (int s2 = 0; s2 < p.ksteps2; ++s2) {
            const int s = p.ksteps + s2;
            const uint64_t dp = desc_kmajor(pbase + (s >> 2) * P_SLAB) + uint64_t((s & 3) * 2);
            const uint64_t dv = desc_mnmajor(v2base, uint32_t(p.kv2_bytes)) + uint64_t(s2 * 128);
            umma_f16(tmem_base + 256 + g * 64, dp, dv, p.idesc_pv, 1u);
          }
        }
        umma_commit(&empty[st]);
        umma_commit(&bar_o[g]);
        if (++st == nst) st = 0;
      }
    }
  };
  if (warp == 0) producer();
  else if (warp == 1) issuer();
  else {
    // ==================================================================================== softmax + epilogue
    const int g = (warp - 2) >> 2;        // group: items of parity g
    const int lg = warp & 3;              // TMEM lane group this warp may access
    const int row = lg * 32 + lane;
    const uint32_t lane_off = uint32_t(lg * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * 128 + lane_off;
    const uint32_t tmem_o = tmem_base + 256 + g * 64 + lane_off;
    const uint32_t prow = smem_u32(sP + g * P_BYTES + row * 128);
    const uint32_t swz = uint32_t(row & 7) << 4;
    const uint32_t a_s = smem_u32(&bar_s[g]), a_sfree = smem_u32(&bar_sfree[g]), a_p = smem_u32(&bar_p[g]);
    const uint32_t a_o = smem_u32(&bar_o[g]), a_ofree = smem_u32(&bar_ofree[g]);
    T* obase = static_cast<T*>(p.o);
    const int sk = int(p.Sk);
    // The output of item k - 1 is read out of TMEM and written to global memory INSIDE item k, between its exponentials and
    // its P store: the P V of item k - 1 has had a whole softmax phase to finish by then, and its latency (and that of the
    // O load) is never waited for.  `pending` describes the item whose output is still to be written.
    struct Pending { T* dst; float inv; bool live; };
    Pending pending{nullptr, 0.f, false};
    auto flush = [&](uint32_t k_done) {  // output of this group's item k_done: TMEM -> registers -> global
      mbar_wait_a(a_o, k_done & 1);
      tcgen05_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tmem_o, r0);
      tmem_ld_32x32(tmem_o + 32, r1);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed_a(a_ofree);
      if (!pending.live) return;
      T* dst = pending.dst;
      const float inv = pending.inv;
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (p.d_out & 7) == 0) {
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          if (c * 8 >= p.d_out) break;
          const uint32_t* r = c < 4 ? r0 + c * 8 : r1 + (c - 4) * 8;
          uint4 v;
          v.x = pack2s<T>(inv * __uint_as_float(r[0]), inv * __uint_as_float(r[1]));
          v.y = pack2s<T>(inv * __uint_as_float(r[2]), inv * __uint_as_float(r[3]));
          v.z = pack2s<T>(inv * __uint_as_float(r[4]), inv * __uint_as_float(r[5]));
          v.w = pack2s<T>(inv * __uint_as_float(r[6]), inv * __uint_as_float(r[7]));
          reinterpret_cast<uint4*>(dst)[c] = v;
        }
      } else {
#pragma unroll
        for (int e = 0; e < HD; ++e)
          if (e < p.d_out) dst[e] = from_f<T>(inv * __uint_as_float(e < 32 ? r0[e] : r1[e - 32]));
      }
    };
    uint32_t k = 0;
    for (uint32_t i = g; i < n_items; i += 2, ++k) {
      const uint32_t w = blockIdx.x + i * gridDim.x;
      const int qt = int(w % n_qt), h = int((w / n_qt) % heads);
      const int64_t b = w / (n_qt * heads);
      mbar_wait_a(a_s, k & 1);
      tcgen05_fence_after();
      float s[NC * 32];
      {
        uint32_t raw[NC][32];
#pragma unroll
        for (int c = 0; c < NC; ++c) tmem_ld_32x32(tmem_s + c * 32, raw[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e) s[c * 32 + e] = __uint_as_float(raw[c][e]);
      }
      float s2[DUAL ? 32 : 1];
      if constexpr (DUAL) {
        uint32_t raw2[32];
        tmem_ld_32x32(tmem_base + 384 + g * 32 + lane_off, raw2);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) s2[e] = __uint_as_float(raw2[e]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed_a(a_sfree);
      if (sk < NC * 32) {  // (NC - 1) * 32 < Sk: only the last chunk can hold columns beyond the keys
#pragma unroll
        for (int e = (NC - 1) * 32; e < NC * 32; ++e)
          if (e >= sk) s[e] = -INFINITY;
      }
      float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
      for (int e = 4; e < NC * 32; e += 4) {
        m0 = fmaxf(m0, s[e]);
        m1 = fmaxf(m1, s[e + 1]);
        m2 = fmaxf(m2, s[e + 2]);
        m3 = fmaxf(m3, s[e + 3]);
      }
      const float mb = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * p.scale_log2e;
      const uint64_t sc2 = f32x2(p.scale_log2e, p.scale_log2e), nmb2 = f32x2(-mb, -mb);
      uint64_t sum_a = f32x2(0.f, 0.f), sum_b = sum_a;
      uint32_t pk[NC * 16];  // the probabilities of this row, packed to the operand dtype
#pragma unroll
      for (int e = 0; e < NC * 32; e += 4) {
        float x0, x1, x2, x3;
        f32x2_split(fma_f32x2(f32x2(s[e], s[e + 1]), sc2, nmb2), x0, x1);
        f32x2_split(fma_f32x2(f32x2(s[e + 2], s[e + 3]), sc2, nmb2), x2, x3);
        const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
        sum_a = add_f32x2(sum_a, f32x2(p0, p1));
        sum_b = add_f32x2(sum_b, f32x2(p2, p3));
        pk[e / 2] = pack2s<T>(p0, p1);
        pk[e / 2 + 1] = pack2s<T>(p2, p3);
      }
      float l0, l1;
      f32x2_split(add_f32x2(sum_a, sum_b), l0, l1);
      // second key set: its own softmax; o = (P V + P2' V2) / l with P2' = p2 * scale2 * l / l2, so that ONE accumulator and
      // one 1 / l in the epilogue serve both terms
      uint32_t pk2[DUAL ? 16 : 1];
      if constexpr (DUAL) {
        const int sk2 = int(p.Sk2);
        float mm = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (e >= sk2) s2[e] = -INFINITY;
          mm = fmaxf(mm, s2[e]);
        }
        const float mb2 = mm * p.scale_log2e;
        float l2 = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          s2[e] = ex2_approx(fmaf(s2[e], p.scale_log2e, -mb2));
          l2 += s2[e];
        }
        const float f = p.scale2 * (l0 + l1) / l2;
#pragma unroll
        for (int e = 0; e < 32; e += 2) pk2[e / 2] = pack2s<T>(s2[e] * f, s2[e + 1] * f);
      }
      // P_g is free once P V of the group's previous item has completed - which is also when its output can be read
      if (k > 0) flush(k - 1);
#pragma unroll
      for (int c = 0; c < NC * 4; ++c)  // 16-byte chunks of 8 keys
        st_shared_v4(prow + (c >> 3) * P_SLAB + ((uint32_t(c & 7) << 4) ^ swz), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      if constexpr (DUAL) {
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) {  // P2' behind the 16 * ksteps columns of P (overwrites P's zero padding there)
          if (c2 < 2 * p.ksteps2) {
            const int c = 2 * p.ksteps + c2;
            st_shared_v4(prow + (c >> 3) * P_SLAB + ((uint32_t(c & 7) << 4) ^ swz), pk2[c2 * 4], pk2[c2 * 4 + 1], pk2[c2 * 4 + 2], pk2[c2 * 4 + 3]);
          }
        }
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_p);
      const int64_t qi = int64_t(qt) * QT + row;
      pending.live = qi < p.Sq;
      pending.dst = obase + b * p.o_sb + qi * p.o_ss + int64_t(h) * p.d_out;
      pending.inv = 1.0f / (l0 + l1);
    }
    if (k > 0) flush(k - 1);
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

// [B, S, H, D] view with strides (sb, ss, D, 1) elements; box = 64 x 1 x `rows` x 1 (columns >= D and rows >= S are zero filled)
int make_map(CUtensorMap* map, int dtype, const void* base, int64_t B, int64_t S, int H, int64_t sb, int64_t ss, int D, int rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) RB200_FAIL(-4, "cuTensorMapEncodeTiled unavailable");
  const cuuint64_t dims[4] = {cuuint64_t(D), cuuint64_t(H), cuuint64_t(S), cuuint64_t(B)};
  const cuuint64_t strides[3] = {cuuint64_t(D) * 2, cuuint64_t(ss) * 2, cuuint64_t(sb) * 2};
  const cuuint32_t box[4] = {64, 1, cuuint32_t(rows), 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = dtype == RB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult rc = fn(map, dt, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) RB200_FAIL(-4, "short sdpa tensor map encode failed (%d): S=%lld H=%d ss=%lld sb=%lld", int(rc), (long long)S, H, (long long)ss, (long long)sb);
  return 0;
}

bool ok_operand(const void* ptr, int64_t sb, int64_t ss) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ss % 8 == 0 && sb % 8 == 0;
}

template <typename T, int NC, bool DUAL>
int launch(cudaStream_t st, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mk2,
           const CUtensorMap& mv2, const ShortParams& prm) {
  static PerDeviceOnce configured;
  if (configured.needed()) {
    if (cudaFuncSetAttribute(tc_sdpa_short_kernel<T, NC, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)) != cudaSuccess)
      RB200_FAIL(-2, "tc_sdpa_short: cannot reserve %zu bytes of shared memory", SMEM_BYTES);
    configured.done();
  }
  const int64_t cap = sm_count();
  const int grid = int(prm.total_work < cap ? prm.total_work : cap);
  tc_sdpa_short_kernel<T, NC, DUAL><<<grid, THREADS, SMEM_BYTES, st>>>(mq, mk, mv, mk2, mv2, prm);
  RB200_CHECK_LAUNCH("tc_sdpa_short");
  return 0;
}

int env_flag(const char* name, int fallback) {
  const char* e = getenv(name);
  return e ? atoi(e) : fallback;
}

}  // namespace

// RB200_ATTN_SHORT: 0 = short key sequences stay on the first-generation flash kernel, 1 (default) = this kernel
bool tc_sdpa_short_supported(const SdpaProblem& p) {
  static const int enabled = env_flag("RB200_ATTN_SHORT", 1);
  if (!enabled) return false;
  if (p.dtype != RB200_BF16 && p.dtype != RB200_FP16) return false;
  if (p.D < 8 || p.D > 64 || (p.D & 7) != 0 || p.causal) return false;
  if (p.bias_h != nullptr) return false;
  if (p.Sk < 1 || p.Sk > 128 || p.Sq < 1 || p.B < 1) return false;
  if (p.k2 != nullptr && p.Sk2 > 0) {  // second K / V set: at most 32 keys, and both P tiles must fit the 128 columns of the P buffer
    static const int dual = env_flag("RB200_ATTN_SHORT_DUAL", 1);
    if (!dual || p.Sk2 > 32 || p.Sk > 96) return false;  // (16 * ksteps + 16 * ksteps2 <= 128 then holds)
    if (!ok_operand(p.k2, p.k2_sb, p.k2_ss) || !ok_operand(p.v2, p.v2_sb, p.v2_ss)) return false;
  }
  if (ceil_div(p.Sq, QT) * p.H * p.B >= (int64_t(1) << 31)) return false;  // 32-bit item arithmetic in the kernel
  return ok_operand(p.q, p.q_sb, p.q_ss) && ok_operand(p.k, p.k_sb, p.k_ss) && ok_operand(p.v, p.v_sb, p.v_ss);
}

int tc_sdpa_short(cudaStream_t st, const SdpaProblem& p) {
  const int ksteps = int(ceil_div(p.Sk, 16));
  CUtensorMap mq, mk, mv;
  if (int rc = make_map(&mq, p.dtype, p.q, p.B, p.Sq, p.H, p.q_sb, p.q_ss, p.D, QT)) return rc;
  if (int rc = make_map(&mk, p.dtype, p.k, p.B, p.Sk, p.H, p.k_sb, p.k_ss, p.D, 16 * ksteps)) return rc;
  if (int rc = make_map(&mv, p.dtype, p.v, p.B, p.Sk, p.H, p.v_sb, p.v_ss, p.D, 16 * ksteps)) return rc;
  const int nc = int(ceil_div(p.Sk, 32));
  const bool dual = p.k2 != nullptr && p.Sk2 > 0;
  const int ksteps2 = dual ? int(ceil_div(p.Sk2, 16)) : 0;
  CUtensorMap mk2 = mk, mv2 = mv;
  if (dual) {
    if (int rc = make_map(&mk2, p.dtype, p.k2, p.B, p.Sk2, p.H, p.k2_sb, p.k2_ss, p.D, 16 * ksteps2)) return rc;
    if (int rc = make_map(&mv2, p.dtype, p.v2, p.B, p.Sk2, p.H, p.v2_sb, p.v2_ss, p.D, 16 * ksteps2)) return rc;
  }
  ShortParams prm{};
  prm.o = p.o;
  prm.o_sb = p.o_sb;
  prm.o_ss = p.o_ss;
  prm.H = p.H;
  prm.Sq = p.Sq;
  prm.Sk = p.Sk;
  prm.n_qt = int(ceil_div(p.Sq, QT));
  prm.total_work = int64_t(prm.n_qt) * p.H * p.B;
  prm.ksteps = ksteps;
  prm.kv_bytes = 16 * ksteps * 128;
  prm.Sk2 = dual ? p.Sk2 : 0;
  prm.ksteps2 = ksteps2;
  prm.kv2_bytes = 16 * ksteps2 * 128;
  prm.scale2 = p.scale2;
  prm.stage_bytes = TILE_BYTES + 2 * prm.kv_bytes + 2 * prm.kv2_bytes;
  prm.nst = RING_BYTES / prm.stage_bytes >= MAX_NST ? MAX_NST : 3;
  prm.scale_log2e = p.scale * 1.4426950408889634f;
  const uint32_t fmt = p.dtype == RB200_BF16 ? 1u : 0u;
  const uint32_t common = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(QT >> 4) << 24);
  prm.idesc_qk = common | (uint32_t((16 * ksteps) >> 3) << 17);      // D = 128 x (16 ksteps): exactly the rows of the K tile
  prm.idesc_pv = common | (uint32_t(HD >> 3) << 17) | (1u << 16);    // D = 128 x 64, B (= V) MN-major
  prm.idesc_qk2 = common | (uint32_t((16 * (dual ? ksteps2 : 1)) >> 3) << 17);
  prm.d_out = p.D;
  const bool bf = p.dtype == RB200_BF16;
#define RB200_SHORT(NC)                                                                                                      \
  if (dual) return bf ? launch<__nv_bfloat16, NC, true>(st, mq, mk, mv, mk2, mv2, prm) : launch<__half, NC, true>(st, mq, mk, mv, mk2, mv2, prm); \
  return bf ? launch<__nv_bfloat16, NC, false>(st, mq, mk, mv, mk2, mv2, prm) : launch<__half, NC, false>(st, mq, mk, mv, mk2, mv2, prm)
  switch (nc) {
    case 1: RB200_SHORT(1);
    case 2: RB200_SHORT(2);
    case 3: RB200_SHORT(3);
    default: RB200_SHORT(4);
  }
#undef RB200_SHORT
}

}  // namespace rb200
