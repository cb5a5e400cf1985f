// tcgen05 attention for SAM's 14 x 14 windows (the 28 windowed blocks of the ViT-H encoder: 25 windows x 16 heads per
// image, 196 queries x 196 keys, head dim 80, decomposed relative-position bias):
//
//   O[bw, q, h, :] = softmax(Q K^T * scale + rel_h[q, kh] + rel_w[q, kw]) V       k = kh * 14 + kw
//
// replaces the windowed branch of foundationals/segment_anything/image_encoder.py:87-143 of the reference
// (FusedSelfAttention with relative-position embeddings).  The first-generation kernel (tc_attention.cu) ran this at
// 0.02 of the tensor peak: 4 softmax warps, head dim padded to 128, four 64-key tiles of which the last holds 4 keys, a
// run-time kh / kw walk per logit (profiles/r02_ncu_sam_window_attention.txt: 127 us per image, 5.5 % tensor pipe).
//
// This kernel is tc_attention2.cu's machinery (one persistent CTA per SM, a PAIR of 128-query tiles = one whole window
// per work item, 16 softmax warps, one MMA-issuing thread per query tile, O resident in TMEM with lazy rescaling, the
// turnstile between the two tiles' exponential phases) with three changes:
//   * head dim 65..80 = one 64-column slab (128-byte swizzle) + one 16-column slab (32-byte swizzle: 32-byte rows, so
//     a 128-row operand tile is 16 + 4 KB instead of 32 KB and Q, two K/V stages and both P tiles still fit).  S takes
//     4 + 1 k-steps; P V is an N = 64 and an N = 16 MMA per k-step into adjacent TMEM columns.
//   * the window geometry is STATIC (14 x 14 keys, two key tiles of 128): which (kh, kw) a logit belongs to, and
//     whether its key exists at all, is known at compile time per (key tile, column half, element) - no index walk, no
//     masking code, and the 60 non-existent keys of the second tile cost no exponentials.
//   * the bias rows of a query (14 + 14 floats, times log2 e) are staged by the softmax threads themselves into a
//     [row][30-float] table (LDS.64 pairs feed the packed FFMA2 directly; 30-word pitch: conflict free per half warp);
//     the global loads for the NEXT window are issued before the current one is written out.
//
// GEOM 1 is the same kernel for SAM's four GLOBAL-attention blocks (64 x 64 tokens, any map whose rows are 64 wide):
// a 128-key tile is two full rows of the map, so a thread's 64 columns are kw = 0..63 of ONE kh.  rel_w[q, 0..63]
// (times log2 e) sits in the table as fp16 (136-byte rows; fp32 would need 64 KB for the two query tiles, fp16 keeps
// 11 bits of a value that is added to logits rounded to bf16 anyway), rel_h[q, kh] is one fp32 scalar per key tile,
// read from global one tile ahead and folded into the maximum / the exponent offset instead of into every logit.
//
// FUSE (GEOM 0): the bias rows are not read from a table that another kernel wrote - they are computed here.  The 27 + 27
// relative-position embeddings of a layer (the same for every window and head) sit in smem as one more K-major operand
// [64 rows x 80]; per window and query tile one N = 64 MMA  T = Q R^T  lands in the S buffer ahead of the first logit
// tile, and each softmax thread scatters its row of T into the [row][30] table: entry idx of a table belongs to the key
// coordinate  c_q + 13 - idx  (rel_pos[c_q - c_k + 13], image_encoder.py:120-131 of the reference).  That removes the
// rel_bias kernel (as long as the attention itself at batch 4) and its 45 MB fp32 table round trip per block.
//
// TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,336) O_B [384,464).
// smem: Q 2 x 20 KB, K/V ring 2 x 40 KB, P 2 x 32 KB, bias table 23 KB (196 rows; GEOM 1: 34 KB), embeddings 10 KB (FUSE), 4 KB exchange.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rb200 {
namespace {

using namespace ptx;

constexpr int QT = 128;             // queries per tile (UMMA M)
constexpr int KT = 128;             // keys per tile
constexpr int WIN = 14;             // window side
constexpr int SK = WIN * WIN;       // 196 keys = queries
constexpr int STAGES = 2;
constexpr int NUM_THREADS = 640;
constexpr int SLAB0 = QT * 128;     // 64 columns, 128-byte rows
constexpr int SLAB1 = QT * 32;      // 16 columns, 32-byte rows
constexpr int TILE_BYTES = SLAB0 + SLAB1;          // one 128-row operand tile: 20 KB
constexpr int Q_BYTES = 2 * TILE_BYTES;            // both query tiles of the window
constexpr int STAGE_BYTES = 2 * TILE_BYTES;        // K tile + V tile
constexpr int P_SLAB = QT * 64 * 2;
constexpr int P_BYTES = 2 * P_SLAB;
constexpr int BIAS_PITCH = 30;                     // floats per query row: [0,14) rel_w, [14,28) rel_h, both times log2 e
constexpr int BIAS_PITCH_G = 136;                  // GEOM 1: bytes per query row, 64 fp16 rel_w entries (+ 8 B: conflict-free LDS.64)
template <int GEOM> constexpr int bias_bytes() { return GEOM == 0 ? SK * BIAS_PITCH * 4 : 2 * QT * BIAS_PITCH_G; }   // a window has 196 queries
constexpr int R_SLAB0 = 64 * 128, R_SLAB1 = 64 * 32;   // FUSE: rel_h rows at [0,27), rel_w rows at [32,59) of a 64-row operand
constexpr int R_BYTES = R_SLAB0 + R_SLAB1;
constexpr int TMEM_COLS = 512;
constexpr int XCHG_BYTES = 2 * 2 * 2 * QT * 4;
template <int GEOM, int FUSE> constexpr size_t smem_bytes() {
  return Q_BYTES + STAGES * STAGE_BYTES + 2 * P_BYTES + (FUSE ? R_BYTES : 0) + 1024 + 256 + XCHG_BYTES + bias_bytes<GEOM>();
}
static_assert(smem_bytes<0, 1>() <= 227 * 1024 && smem_bytes<1, 0>() <= 227 * 1024, "exceeds the 227 KB of shared memory a CTA can opt in to");
constexpr float RESCALE_LOG2 = 8.0f;
constexpr float L2E = 1.4426950408889634f;

struct WinParams {
  void* o;
  int64_t o_sb, o_ss;
  int H;
  int64_t total_work;   // windows x heads (GEOM 1: images x heads x query-tile pairs)
  int n_pairs;          // pairs of 128-query tiles per (image, head): 1 for a window
  int ntiles;           // 128-key tiles: 2 for a window
  int bias_H;           // rows of the map (GEOM 1)
  float scale_log2e;
  uint32_t idesc_qk, idesc_pv64, idesc_pv16, idesc_t;
  int d_out;
  const float* bias;    // [window * H + head][query tile][bias_H rel_h + bias_W rel_w][128] (sam_attention.cu)
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

// 32-byte-swizzle operand slabs (16 columns of bf16 per row, rows 32 B apart, 8-row groups 256 B apart).  In units of
// 16 B the canonical layouts are K-major ((8,n),2):((2,SBO),1) and MN-major ((2,n),(8,k)):((1,LBO),(2,SBO)): one
// 16-element atom along K (resp. MN), so only SBO = 256 B matters; layout type 6 = SWIZZLE_32B.
__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr) {
  return uint64_t((addr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t(256 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(6) << 61);
}

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint64_t ld_shared_b64(uint32_t addr) {
  uint64_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));   // volatile: reloaded at every use instead of 14 more live registers
  return v;
}
__device__ __forceinline__ float ld_shared_f32_nv(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_b64(uint32_t addr, float lo, float hi) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(lo), "f"(hi) : "memory");
}

template <typename T, int GEOM, int FUSE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_sdpa_win_kernel(const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                   const __grid_constant__ CUtensorMap map_k0, const __grid_constant__ CUtensorMap map_k1,
                   const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_v1,
                   const __grid_constant__ CUtensorMap map_rh0, const __grid_constant__ CUtensorMap map_rh1,
                   const __grid_constant__ CUtensorMap map_rw0, const __grid_constant__ CUtensorMap map_rw1, const WinParams p) {
  static_assert(FUSE == 0 || GEOM == 0, "the fused relative-position product exists for the 14 x 14 windows only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                           // [2 tiles][slab 0 | slab 1]
  uint8_t* sKV = sQ + Q_BYTES;                  // [STAGES][K slab 0 | K slab 1 | V slab 0 | V slab 1]
  uint8_t* sP = sKV + STAGES * STAGE_BYTES;     // [2 groups][2 slabs][128 x 64]
  uint8_t* sR = sP + 2 * P_BYTES;               // FUSE: [64 rows][slab 0 | slab 1] relative-position embeddings
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + (FUSE ? R_BYTES : 0));
  uint64_t* kv_full = bars;                     // [STAGES]
  uint64_t* kv_empty = bars + STAGES;           // [STAGES]
  uint64_t* q_full = bars + 2 * STAGES;         // [1]
  uint64_t* q_empty = q_full + 1;               // [1]
  uint64_t* bar_s = q_full + 2;                 // [2 groups]
  uint64_t* bar_sfree = q_full + 4;
  uint64_t* bar_p = q_full + 6;
  uint64_t* bar_o = q_full + 8;
  uint64_t* bar_ofree = q_full + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 12);
  uint64_t* r_full = q_full + 13;               // FUSE: the embeddings have landed
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  float* sBias = xchg + XCHG_BYTES / 4;         // [2 groups][128 rows][BIAS_PITCH floats | BIAS_PITCH_G bytes]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = warp >> 2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_q0);
    prefetch_tmap(&map_q1);
    prefetch_tmap(&map_k0);
    prefetch_tmap(&map_k1);
    prefetch_tmap(&map_v0);
    prefetch_tmap(&map_v1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);
    mbar_init(r_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_s[b], 1);
      mbar_init(&bar_sfree[b], 8);
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
        uint32_t n = 0;
        if constexpr (FUSE) {
          const uint32_t r = smem_u32(sR);
          mbar_arrive_expect_tx(r_full, R_BYTES);   // rows >= 27 of each 32-row box are out of bounds: zero filled, counted
          tma_load_2d(r, &map_rh0, r_full, 0, 0);
          tma_load_2d(r + 32 * 128, &map_rw0, r_full, 0, 0);
          tma_load_2d(r + R_SLAB0, &map_rh1, r_full, 64, 0);
          tma_load_2d(r + R_SLAB0 + 32 * 32, &map_rw1, r_full, 64, 0);
        }
        for (int64_t w = blockIdx.x; w < p.total_work; w += gridDim.x, ++n) {
          const int pair = int(w % p.n_pairs);
          const int h = int((w / p.n_pairs) % p.H);
          const int b = int(w / (int64_t(p.n_pairs) * p.H));
#pragma unroll 1
          for (int j = 0; j < p.ntiles; ++j) {
            mbar_wait(&kv_empty[stage], phase ^ 1, 2);
            mbar_arrive_expect_tx(&kv_full[stage], STAGE_BYTES);
            uint8_t* st = sKV + stage * STAGE_BYTES;
            tma_load_4d(st, &map_k0, &kv_full[stage], 0, h, j * KT, b);
            tma_load_4d(st + SLAB0, &map_k1, &kv_full[stage], 64, h, j * KT, b);
            tma_load_4d(st + TILE_BYTES, &map_v0, &kv_full[stage], 0, h, j * KT, b);
            tma_load_4d(st + TILE_BYTES + SLAB0, &map_v1, &kv_full[stage], 64, h, j * KT, b);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            if (j == 0) {
              // Q is single buffered: its slot frees when the previous window's last P V has been issued by both MMA
              // threads - later than the first K/V stage, which is therefore requested first
              mbar_wait(q_empty, (n & 1) ^ 1, 1);
              mbar_arrive_expect_tx(q_full, Q_BYTES);
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                tma_load_4d(sQ + g * TILE_BYTES, &map_q0, q_full, 0, h, (pair * 2 + g) * QT, b);
                tma_load_4d(sQ + g * TILE_BYTES + SLAB0, &map_q1, q_full, 64, h, (pair * 2 + g) * QT, b);
              }
            }
          }
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ================================================================================ MMA (one issuer per query tile)
      if (lane == 0) {
        const int g = warp - 1;
        int st_k = 0, st_v = 0;
        uint32_t ph_k = 0;
        uint32_t t = 0, s_issued = 0, n = 0;
        const uint32_t tmem_s = tmem_base + g * 128, tmem_o = tmem_base + 256 + g * 128;
        const uint32_t pbase = smem_u32(sP + g * P_BYTES);
        const uint32_t qbase = smem_u32(sQ + g * TILE_BYTES);
        const uint32_t kvbase = smem_u32(sKV);
        auto issue_s = [&](int stage_k) {
          if (s_issued > 0) mbar_wait(&bar_sfree[g], (s_issued - 1) & 1, 4);
          ++s_issued;
          tcgen05_fence_after();
          const uint32_t kb = kvbase + stage_k * STAGE_BYTES;
          const uint64_t dq = desc_kmajor(qbase), dk = desc_kmajor(kb);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_s, dq + uint64_t(k * 2), dk + uint64_t(k * 2), p.idesc_qk, k > 0);
          umma_f16(tmem_s, desc_sw32(qbase + SLAB0), desc_sw32(kb + SLAB0), p.idesc_qk, 1u);  // columns 64..79
          umma_commit(&bar_s[g]);
        };
        // T_g = Q_g R^T (FUSE): 64 columns of the S buffer, the same hand-over as an S tile
        auto issue_t = [&]() {
          if (s_issued > 0) mbar_wait(&bar_sfree[g], (s_issued - 1) & 1, 4);
          ++s_issued;
          tcgen05_fence_after();
          const uint32_t rb = smem_u32(sR);
          const uint64_t dq = desc_kmajor(qbase), dr = desc_kmajor(rb);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_s, dq + uint64_t(k * 2), dr + uint64_t(k * 2), p.idesc_t, k > 0);
          umma_f16(tmem_s, desc_sw32(qbase + SLAB0), desc_sw32(rb + R_SLAB0), p.idesc_t, 1u);
          umma_commit(&bar_s[g]);
        };
        auto next_k = [&]() {
          if (++st_k == STAGES) {
            st_k = 0;
            ph_k ^= 1;
          }
        };
        for (int64_t w = blockIdx.x; w < p.total_work; w += gridDim.x, ++n) {
          mbar_wait(q_full, n & 1, 5);
          if constexpr (FUSE) {
            if (n == 0) mbar_wait(r_full, 0, 9);
            issue_t();
          }
          mbar_wait(&kv_full[st_k], ph_k, 3);
          issue_s(st_k);
          next_k();
#pragma unroll 1
          for (int j = 0; j < p.ntiles; ++j) {
            if (j + 1 < p.ntiles) {
              mbar_wait(&kv_full[st_k], ph_k, 6);
              issue_s(st_k);
              next_k();
              if (j + 2 == p.ntiles) umma_commit(q_empty);   // the window's last S has been issued: Q frees when it retires (count 2)
            }
            mbar_wait(&bar_p[g], t & 1, 7);
            if (j == 0 && n > 0) mbar_wait(&bar_ofree[g], (n - 1) & 1, 8);
            tcgen05_fence_after();
            const uint32_t vb = kvbase + st_v * STAGE_BYTES + TILE_BYTES;
            const uint64_t dv0 = desc_mnmajor(vb, SLAB0), dv1 = desc_sw32(vb + SLAB0);
#pragma unroll
            for (int k = 0; k < KT / 16; ++k) {
              if (GEOM == 0 && j > 0 && k * 16 >= SK - KT) break;   // a window's second tile holds 68 keys: 5 k-steps
              const uint64_t dp = desc_kmajor(pbase + (k >> 2) * P_SLAB) + uint64_t((k & 3) * 2);
              const uint32_t acc = (j > 0 || k > 0) ? 1u : 0u;
              umma_f16(tmem_o, dp, dv0 + uint64_t(k * 128), p.idesc_pv64, acc);        // 16 key rows x 128 B
              umma_f16(tmem_o + 64, dp, dv1 + uint64_t(k * 32), p.idesc_pv16, acc);    // 16 key rows x 32 B
            }
            umma_commit(&kv_empty[st_v]);
            umma_commit(&bar_o[g]);
            ++t;
            if (++st_v == STAGES) st_v = 0;
          }
        }
      }
    }
  } else {
    // ============================================================================ softmax
    setmaxnreg_inc<104>();
    const int g = (wg - 1) >> 1;          // query tile of the window
    const int hf = (wg - 1) & 1;          // which 64 key columns of a tile (and which 32 + 16 output columns) this thread owns
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const uint32_t lane_off = uint32_t(lg * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * 128 + hf * 64 + lane_off;
    const uint32_t tmem_o = tmem_base + 256 + g * 128 + hf * 32 + lane_off;
    const uint32_t tmem_o1 = tmem_base + 256 + g * 128 + 64 + lane_off;   // columns 64..79: the hf == 0 threads' job
    const uint32_t prow = smem_u32(sP + g * P_BYTES + hf * P_SLAB + row * 128);
    const int sw = row & 7;
    const uint32_t xg = smem_u32(xchg + g * (2 * 2 * QT));
    const uint32_t x_mine = xg + uint32_t(hf * QT + row) * 4, x_other = xg + uint32_t((hf ^ 1) * QT + row) * 4;
    const uint32_t a_s = smem_u32(&bar_s[g]), a_sfree = smem_u32(&bar_sfree[g]), a_p = smem_u32(&bar_p[g]);
    const uint32_t a_o = smem_u32(&bar_o[g]), a_ofree = smem_u32(&bar_ofree[g]);
    const uint32_t swz = uint32_t(sw) << 4;
    const uint32_t xbar = 1 + g;
    const bool live = GEOM == 1 || g * QT + row < SK;                       // the query exists
    // this query's bias row (a window's table has 196 rows: the 60 threads without a query read the last one)
    const uint32_t brow = smem_u32(sBias) + uint32_t(GEOM == 0 && !live ? SK - 1 : g * QT + row) * (GEOM == 0 ? BIAS_PITCH * 4 : BIAS_PITCH_G);
    T* obase = static_cast<T*>(p.o);
    uint32_t t = 0, n = 0;
    uint32_t si = 0;                      // tiles taken out of the S buffer so far (FUSE: one more per window than `t`)
    const int nqt = 2 * p.n_pairs, bias_K = GEOM == 0 ? 2 * WIN : p.bias_H + 64;

    // GEOM 0: this thread's half of the NEXT window's bias rows (hf 0: the 14 rel_w entries, hf 1: the 14 rel_h entries)
    float nb[GEOM == 0 && !FUSE ? WIN : 1];
    auto fetch_bias = [&](int64_t w) {
      const float* blk = p.bias + ((w * 2 + g) * (2 * WIN) + (hf == 0 ? WIN : 0)) * 128 + row;
#pragma unroll
      for (int i = 0; i < WIN; ++i) nb[i] = live ? __ldg(blk + i * 128) : 0.f;
    };
    if constexpr (GEOM == 0 && !FUSE) fetch_bias(blockIdx.x);

    for (int64_t w = blockIdx.x; w < p.total_work; w += gridDim.x, ++n) {
      const int pair = int(w % p.n_pairs);
      const int h = int((w / p.n_pairs) % p.H);
      const int64_t b = w / (int64_t(p.n_pairs) * p.H);
      // this query tile's block of the bias table: [bias_H rel_h rows | bias_W rel_w rows][128 queries]
      const float* blk = p.bias + (((w / p.n_pairs) * nqt + pair * 2 + g) * bias_K) * 128 + row;
      float bh_cur = 0.f;
      // publish the bias rows (the group's previous reads of the table ended before its last named barrier)
      if constexpr (GEOM == 0 && FUSE) {
        // this thread's 32 columns of T = Q R^T: hf 0 the products with rel_h[0..27), hf 1 with rel_w[0..27)
        mbar_wait_a(a_s, si & 1);
        ++si;
        tcgen05_fence_after();
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + g * 128 + hf * 32 + lane_off, r);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed_a(a_sfree);
        const int q = g * QT + row, hq = q / WIN;
        const int c = hf == 0 ? hq : q - hq * WIN;                       // the query's coordinate along this table's axis
        const uint32_t dst = brow + uint32_t((hf == 0 ? WIN : 0) + c + WIN - 1) * 4;   // entry idx -> key coordinate c + 13 - idx
#pragma unroll
        for (int idx = 0; idx < 2 * WIN - 1; ++idx)
          if (live && unsigned(c + WIN - 1 - idx) < unsigned(WIN)) st_shared_f32(dst - uint32_t(idx) * 4, __uint_as_float(r[idx]) * L2E);
      } else if constexpr (GEOM == 0) {
        if (live) {
#pragma unroll
          for (int i = 0; i < WIN; i += 2) st_shared_b64(brow + uint32_t((hf == 0 ? 0 : WIN) + i) * 4, nb[i] * L2E, nb[i + 1] * L2E);
        }
      } else {
        // rel_w[q, hf * 32 .. + 32) as fp16; rel_h of the first key tile
        const float* bw = blk + (p.bias_H + hf * 32) * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const __half2 lo = __floats2half2_rn(__ldg(bw + (c * 4) * 128) * L2E, __ldg(bw + (c * 4 + 1) * 128) * L2E);
          const __half2 hi = __floats2half2_rn(__ldg(bw + (c * 4 + 2) * 128) * L2E, __ldg(bw + (c * 4 + 3) * 128) * L2E);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(brow + uint32_t(hf * 64 + c * 8)), "r"(*reinterpret_cast<const uint32_t*>(&lo)),
                       "r"(*reinterpret_cast<const uint32_t*>(&hi))
                       : "memory");
        }
        bh_cur = __ldg(blk + hf * 128) * L2E;
      }
      named_bar_sync(xbar, 256);
      float m_run = -INFINITY, l_run = 0.f;

      // One key tile.  GEOM 0: J (key tile) and HF (column half) are compile-time, so every logit's (kh, kw) and whether its
      // key exists are constants.  GEOM 1: J = HF = -1, `j` is the run-time tile index and all 64 columns are kw = 0..63 of
      // the map row kh = 2 j + hf.
      auto tile = [&](auto jc, auto hc, int j) {
        constexpr int J = decltype(jc)::value, HF = decltype(hc)::value;
        constexpr int K0 = J * KT + HF * 64;                       // GEOM 0: first key of this thread's 64 columns
        constexpr int NV = GEOM == 1 ? 64 : (SK - K0 >= 64 ? 64 : (SK - K0 > 0 ? SK - K0 : 0));   // keys that exist among them (even)
        const bool first = GEOM == 0 ? J == 0 : j == 0;
        mbar_wait_a(a_s, si & 1);
        ++si;
        tcgen05_fence_after();
        float s[64];
        {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(tmem_s, r0);
          if constexpr (NV > 32) tmem_ld_32x32(tmem_s + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s[i] = __uint_as_float(r0[i]);
            if constexpr (NV > 32) s[32 + i] = __uint_as_float(r1[i]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed_a(a_sfree);
        // logits in the log2 domain: s * scale * log2 e + rel_w[kw] (+ rel_h[kh]); the table holds the bias times log2 e
        const uint64_t sc2 = f32x2(p.scale_log2e, p.scale_log2e);
        float tm0 = -INFINITY, tm1 = -INFINITY, bh = 0.f;
        if constexpr (GEOM == 0) {
#pragma unroll
          for (int i = 0; i < NV; i += 2) {
            const int k = K0 + i, kh = k / WIN, kw = k % WIN;         // constants after unrolling; kw is even
            const float bhv = ld_shared_f32_nv(brow + uint32_t(WIN + kh) * 4);
            const uint64_t x = add_f32x2(fma_f32x2(f32x2(s[i], s[i + 1]), sc2, ld_shared_b64(brow + uint32_t(kw) * 4)), f32x2(bhv, bhv));
            f32x2_split(x, s[i], s[i + 1]);
            if ((i >> 1) & 1) tm1 = fmaxf(tm1, fmaxf(s[i], s[i + 1]));
            else tm0 = fmaxf(tm0, fmaxf(s[i], s[i + 1]));
          }
        } else {
          bh = bh_cur;   // rel_h[q, 2 j + hf]: the same for all 64 columns - it joins the maximum and the exponent offset
          if (j + 1 < p.ntiles) bh_cur = __ldg(blk + (2 * (j + 1) + hf) * 128) * L2E;   // in flight during this tile
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            const uint64_t h4 = ld_shared_b64(brow + uint32_t(i) * 2);
            const uint32_t lo = uint32_t(h4), hi = uint32_t(h4 >> 32);
            const float2 b01 = __half22float2(*reinterpret_cast<const __half2*>(&lo));
            const float2 b23 = __half22float2(*reinterpret_cast<const __half2*>(&hi));
            f32x2_split(fma_f32x2(f32x2(s[i], s[i + 1]), sc2, f32x2(b01.x, b01.y)), s[i], s[i + 1]);
            f32x2_split(fma_f32x2(f32x2(s[i + 2], s[i + 3]), sc2, f32x2(b23.x, b23.y)), s[i + 2], s[i + 3]);
            tm0 = fmaxf(tm0, fmaxf(s[i], s[i + 1]));
            tm1 = fmaxf(tm1, fmaxf(s[i + 2], s[i + 3]));
          }
        }
        const uint32_t slot = (t & 1) * (2 * QT * 4);
        const float mine = fmaxf(tm0, tm1) + bh;
        st_shared_f32(x_mine + slot, mine);
        named_bar_sync(xbar, 256);
        const float tmax = fmaxf(mine, ld_shared_f32(x_other + slot));
        bool waited_o = false;
        if (first) {
          m_run = tmax;
        } else {
          const float m_new = fmaxf(m_run, tmax);
          const bool grew = m_new - m_run > RESCALE_LOG2;
          if (__any_sync(0xffffffffu, grew)) {
            mbar_wait_a(a_o, (t - 1) & 1);
            waited_o = true;
            tcgen05_fence_after();
            const float alpha = ex2_approx(m_run - m_new);
            uint32_t raw[32];
            tmem_ld_32x32(tmem_o, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
            tmem_st_32x32(tmem_o, raw);
            if (hf == 0) {
              uint32_t r2[16];
              tmem_ld_32x16(tmem_o1, r2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) r2[i] = __float_as_uint(__uint_as_float(r2[i]) * alpha);
              tmem_st_32x16(tmem_o1, r2);
            }
            tmem_st_wait();
            l_run *= alpha;
            m_run = m_new;
          }
        }
        if ((!first && !waited_o) || (first && t > 0)) mbar_wait_a(a_o, (t - 1) & 1);   // P buffer free
        if (g == 1) named_bar_sync(3, 512);                                              // turnstile, see tc_attention2.cu
        else if (t > 0) named_bar_sync(4, 512);
        const uint64_t nm2 = f32x2(bh - m_run, bh - m_run);
        uint64_t sum_a = f32x2(0.f, 0.f), sum_b = sum_a;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t pk[4] = {0u, 0u, 0u, 0u};
          if (c * 8 < NV) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int i = c * 8 + q * 4;
              if (i < NV) {
                float x0, x1;
                f32x2_split(add_f32x2(f32x2(s[i], s[i + 1]), nm2), x0, x1);
                const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                sum_a = add_f32x2(sum_a, f32x2(p0, p1));
                pk[q * 2] = pack2<T>(p0, p1);
              }
              if (i + 2 < NV) {
                float x2, x3;
                f32x2_split(add_f32x2(f32x2(s[i + 2], s[i + 3]), nm2), x2, x3);
                const float p2 = ex2_approx(x2), p3 = ex2_approx(x3);
                sum_b = add_f32x2(sum_b, f32x2(p2, p3));
                pk[q * 2 + 1] = pack2<T>(p2, p3);
              }
            }
          }
          st_shared_v4(prow + ((uint32_t(c) << 4) ^ swz), pk[0], pk[1], pk[2], pk[3]);
        }
        {
          float a0, a1;
          f32x2_split(add_f32x2(sum_a, sum_b), a0, a1);
          l_run += a0 + a1;
        }
        named_bar_arrive(g == 0 ? 3 : 4, 512);
        fence_proxy_async();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_p);
        ++t;
      };
      if constexpr (GEOM == 0) {
        if (hf == 0) {
          tile(std::integral_constant<int, 0>(), std::integral_constant<int, 0>(), 0);
          tile(std::integral_constant<int, 1>(), std::integral_constant<int, 0>(), 1);
        } else {
          tile(std::integral_constant<int, 0>(), std::integral_constant<int, 1>(), 0);
          tile(std::integral_constant<int, 1>(), std::integral_constant<int, 1>(), 1);
        }
        // next window's bias rows: in flight while this one is written out
        if constexpr (!FUSE) {
          if (w + gridDim.x < p.total_work) fetch_bias(w + gridDim.x);
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < p.ntiles; ++j) tile(std::integral_constant<int, -1>(), std::integral_constant<int, -1>(), j);
      }

      const uint32_t slot = (t & 1) * (2 * QT * 4);
      st_shared_f32(x_mine + slot, l_run);
      named_bar_sync(xbar, 256);
      const float l_row = l_run + ld_shared_f32(x_other + slot);
      named_bar_sync(xbar, 256);
      mbar_wait_a(a_o, (t - 1) & 1);
      tcgen05_fence_after();
      const float inv = l_row > 0.f ? 1.f / l_row : 0.f;
      float acc[32], acc1[16];
      {
        uint32_t raw[32], r2[16];
        tmem_ld_32x32(tmem_o, raw);
        if (hf == 0) tmem_ld_32x16(tmem_o1, r2);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = inv * __uint_as_float(raw[i]);
        if (hf == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc1[i] = inv * __uint_as_float(r2[i]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed_a(a_ofree);
      if (live) {
        const int64_t qi = (int64_t(pair) * 2 + g) * QT + row;
        T* dst = obase + b * p.o_sb + qi * p.o_ss + int64_t(h) * p.d_out;   // 16-byte aligned: checked on the host
        uint4* d0 = reinterpret_cast<uint4*>(dst + hf * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v;
          v.x = pack2<T>(acc[c * 8], acc[c * 8 + 1]);
          v.y = pack2<T>(acc[c * 8 + 2], acc[c * 8 + 3]);
          v.z = pack2<T>(acc[c * 8 + 4], acc[c * 8 + 5]);
          v.w = pack2<T>(acc[c * 8 + 6], acc[c * 8 + 7]);
          d0[c] = v;
        }
        if (hf == 0) {
          uint4* d1 = reinterpret_cast<uint4*>(dst + 64);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (64 + c * 8 >= p.d_out) break;
            uint4 v;
            v.x = pack2<T>(acc1[c * 8], acc1[c * 8 + 1]);
            v.y = pack2<T>(acc1[c * 8 + 2], acc1[c * 8 + 3]);
            v.z = pack2<T>(acc1[c * 8 + 4], acc1[c * 8 + 5]);
            v.w = pack2<T>(acc1[c * 8 + 6], acc1[c * 8 + 7]);
            d1[c] = v;
          }
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

// [B, S, H, D] view with strides (sb, ss, D, 1) elements; box = cols x 1 x 128 rows x 1 (columns >= D, rows >= S are zero filled)
int make_map(CUtensorMap* map, int dtype, const void* base, int64_t B, int64_t S, int H, int64_t sb, int64_t ss, int D, int cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) RB200_FAIL(-4, "cuTensorMapEncodeTiled unavailable");
  const cuuint64_t dims[4] = {cuuint64_t(D), cuuint64_t(H), cuuint64_t(S), cuuint64_t(B)};
  const cuuint64_t strides[3] = {cuuint64_t(D) * 2, cuuint64_t(ss) * 2, cuuint64_t(sb) * 2};
  const cuuint32_t box[4] = {cuuint32_t(cols), 1, 128, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = dtype == RB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult rc = fn(map, dt, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) RB200_FAIL(-4, "sdpa_win tensor map encode failed (%d): cols=%d S=%lld H=%d ss=%lld sb=%lld", int(rc), cols, (long long)S, H, (long long)ss, (long long)sb);
  return 0;
}

bool ok_operand(const void* ptr, int64_t sb, int64_t ss) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ss % 8 == 0 && sb % 8 == 0;
}

// the relative-position embeddings [2 W - 1, D] as a 2-D map; box = cols x 32 rows
int make_rel_map(CUtensorMap* map, int dtype, const void* base, int rows, int D, int cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) RB200_FAIL(-4, "cuTensorMapEncodeTiled unavailable");
  const cuuint64_t dims[2] = {cuuint64_t(D), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(D) * 2};
  const cuuint32_t box[2] = {cuuint32_t(cols), 32};
  const cuuint32_t es[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == RB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult rc = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) RB200_FAIL(-4, "sdpa_win embedding map encode failed (%d): rows=%d D=%d cols=%d", int(rc), rows, D, cols);
  return 0;
}

template <typename T, int GEOM, int FUSE>
int launch(cudaStream_t st, const CUtensorMap (&m)[10], const WinParams& prm) {
  static PerDeviceOnce configured;
  if (configured.needed()) {
    if (cudaFuncSetAttribute(tc_sdpa_win_kernel<T, GEOM, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes<GEOM, FUSE>())) != cudaSuccess)
      RB200_FAIL(-2, "tc_sdpa_win: cannot reserve %zu bytes of shared memory", smem_bytes<GEOM, FUSE>());
    configured.done();
  }
  const int64_t cap = sm_count();
  const int grid = int(prm.total_work < cap ? prm.total_work : cap);
  tc_sdpa_win_kernel<T, GEOM, FUSE><<<grid, NUM_THREADS, smem_bytes<GEOM, FUSE>(), st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], prm);
  RB200_CHECK_LAUNCH("tc_sdpa_win");
  return 0;
}

// 0: a 14 x 14 window, 1: a map with 64-wide rows (an even number of them), -1: neither
int geometry(const SdpaProblem& p) {
  if (p.bias_H == WIN && p.bias_W == WIN && p.Sq == SK && p.Sk == SK) return 0;
  if (p.bias_W == 64 && p.bias_H >= 4 && p.bias_H % 4 == 0 && p.Sk == int64_t(p.bias_H) * 64 && p.Sq == p.Sk) return 1;
  return -1;
}

}  // namespace

// RB200_ATTN_WIN: 0 = SAM's attention stays on the first-generation kernel, 1 = only the 14 x 14 windows run here,
// 2 (default) = the windows and the global (64-wide map) blocks.
bool tc_sdpa_win_supported(const SdpaProblem& p) {
  static const int enabled = [] {
    const char* e = getenv("RB200_ATTN_WIN");
    return e ? atoi(e) : 2;
  }();
  if (!enabled) return false;
  if (p.dtype != RB200_BF16 && p.dtype != RB200_FP16) return false;
  if (p.D <= 64 || p.D > 80 || (p.D & 7) != 0 || p.causal) return false;
  if (p.bias_h == nullptr || p.k2 != nullptr || p.B < 1) return false;
  const int geom = geometry(p);
  if (geom < 0 || (geom == 1 && enabled < 2)) return false;
  // the fp16 rel_w table of GEOM 1 carries 11 bits: below bf16's own rounding of the probabilities, not below fp16's
  if (geom == 1 && p.dtype != RB200_BF16) return false;
  if ((reinterpret_cast<uintptr_t>(p.o) & 15) != 0 || p.o_ss % 8 != 0 || p.o_sb % 8 != 0) return false;
  return ok_operand(p.q, p.q_sb, p.q_ss) && ok_operand(p.k, p.k_sb, p.k_ss) && ok_operand(p.v, p.v_sb, p.v_ss);
}

// RB200_ATTN_WIN_FUSE: 1 (default) = the kernel computes the windows' relative-position bias itself (the caller then skips
// the rel_bias kernel), 0 = it reads the table.
bool tc_sdpa_win_fuses_bias(const SdpaProblem& p) {
  static const int enabled = [] {
    const char* e = getenv("RB200_ATTN_WIN_FUSE");
    return e ? atoi(e) : 1;
  }();
  auto al = [](const void* q) { return q != nullptr && (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return enabled && tc_sdpa_win_supported(p) && geometry(p) == 0 && al(p.rel_h_emb) && al(p.rel_w_emb);
}

int tc_sdpa_win(cudaStream_t st, const SdpaProblem& p) {
  CUtensorMap m[10];
  const void* base[3] = {p.q, p.k, p.v};
  const int64_t sb[3] = {p.q_sb, p.k_sb, p.v_sb}, ss[3] = {p.q_ss, p.k_ss, p.v_ss};
  for (int i = 0; i < 3; ++i) {
    if (int rc = make_map(&m[2 * i], p.dtype, base[i], p.B, p.Sk, p.H, sb[i], ss[i], p.D, 64)) return rc;
    if (int rc = make_map(&m[2 * i + 1], p.dtype, base[i], p.B, p.Sk, p.H, sb[i], ss[i], p.D, 16)) return rc;
  }
  const int geom = geometry(p);
  const bool fuse = tc_sdpa_win_fuses_bias(p);
  if (fuse) {
    if (int rc = make_rel_map(&m[6], p.dtype, p.rel_h_emb, 2 * WIN - 1, p.D, 64)) return rc;
    if (int rc = make_rel_map(&m[7], p.dtype, p.rel_h_emb, 2 * WIN - 1, p.D, 16)) return rc;
    if (int rc = make_rel_map(&m[8], p.dtype, p.rel_w_emb, 2 * WIN - 1, p.D, 64)) return rc;
    if (int rc = make_rel_map(&m[9], p.dtype, p.rel_w_emb, 2 * WIN - 1, p.D, 16)) return rc;
  } else {
    for (int i = 6; i < 10; ++i) m[i] = m[0];   // unused
  }
  WinParams prm{};
  prm.o = p.o;
  prm.o_sb = p.o_sb;
  prm.o_ss = p.o_ss;
  prm.H = p.H;
  prm.n_pairs = int(ceil_div(p.Sq, 2 * QT));
  prm.ntiles = int(ceil_div(p.Sk, KT));
  prm.bias_H = p.bias_H;
  prm.total_work = int64_t(p.H) * p.B * prm.n_pairs;
  prm.scale_log2e = p.scale * L2E;
  const uint32_t fmt = p.dtype == RB200_BF16 ? 1u : 0u;
  const uint32_t common = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(QT >> 4) << 24);
  prm.idesc_qk = common | (uint32_t(KT >> 3) << 17);               // D = 128 x 128, A and B K-major
  prm.idesc_pv64 = common | (uint32_t(64 >> 3) << 17) | (1u << 16);   // D = 128 x 64, B (= V) MN-major
  prm.idesc_pv16 = common | (uint32_t(16 >> 3) << 17) | (1u << 16);   // D = 128 x 16
  prm.idesc_t = common | (uint32_t(64 >> 3) << 17);                   // D = 128 x 64, A and B K-major
  prm.d_out = p.D;
  prm.bias = p.bias_h;
  const bool bf = p.dtype == RB200_BF16;
  if (geom == 0 && fuse) return bf ? launch<__nv_bfloat16, 0, 1>(st, m, prm) : launch<__half, 0, 1>(st, m, prm);
  if (geom == 0) return bf ? launch<__nv_bfloat16, 0, 0>(st, m, prm) : launch<__half, 0, 0>(st, m, prm);
  return bf ? launch<__nv_bfloat16, 1, 0>(st, m, prm) : launch<__half, 1, 0>(st, m, prm);
}

}  // namespace rb200
