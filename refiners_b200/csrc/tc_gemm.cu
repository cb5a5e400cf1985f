// tcgen05 GEMM / implicit-GEMM conv for sm_100a (bf16 / fp16 operands, fp32 accumulate in TMEM).
//
//   D[M, N] = A[M, Ktot] * B[N, Ktot]^T   with fused epilogue (see include/refiners_b200.h)
//
// Structure (persistent, warp specialised, one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 4D (A) / 3D (B) -> 128B-swizzled smem ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, SS mode)
//   warps 2..9  epilogue: tcgen05.ld accumulator -> registers -> bias / LoRA scale / act /
//               residual -> 16-byte global stores; overlaps the next tile's main loop through
//               two TMEM accumulator stages.
// A is addressed through a 4D tensor map so that the same kernel runs
//   * plain GEMM      dims (K, M, 1, 1), box (64, 128, 1, 1)
//   * conv (NHWC)     dims (C, W, H, B), box (64, TW, TH, TB) with TW*TH*TB = 128 output pixels;
//                     one K step per (filter tap, 64-channel block); padding comes from TMA's
//                     out-of-bounds zero fill and stride-2 from the map's element strides -
//                     im2col is never materialised.
// An optional second K segment (A2, B2) appends extra K blocks; LoRA up-projections use it.
#include <cuda.h>

#include <cstdio>

#include <cstdlib>

#include "common.cuh"

namespace rb200 {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;   // 64 x 2 bytes = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per 32-lane TMEM group)

struct TcParams {
  int64_t M, N;             // logical output rows / accumulator columns
  int tiles_m, tiles_n;
  int k_iters1;             // taps * kblocks
  int kblocks;              // 64-wide K blocks per tap (segment 1)
  int k_iters2;             // K blocks of segment 2
  int kw;                   // filter width S (tap -> (r, s))
  // conv tiling
  int conv;
  int TW, TH, TB;
  int tiles_w, tiles_h;
  int64_t Ho, Wo;
  int stride, pad;
  // epilogue
  const void* bias;
  const float* colscale;
  const void* chan_bias;    // [B, N]
  const void* residual;
  int64_t ldr;
  void* y;
  int64_t ldy;
  int epilogue;
  uint32_t idesc;
};

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __noinline__ void mbar_timeout(int tag, uint32_t parity) {
  printf("rb200 gemm: mbarrier wait timed out: tag=%d parity=%u block=%d thread=%d\n", tag, parity, int(blockIdx.x), int(threadIdx.x));
  __trap();
}
// Bounded wait: a protocol bug reports which barrier starved and traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) mbar_timeout(tag, parity);
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// B-tile half loaded once from L2 and written into the same smem offset of every CTA in `mask`;
// each destination CTA's mbarrier (same offset) receives the complete_tx for these bytes
__device__ __forceinline__ void tma_load_3d_mcast(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// CTA-pair (cta_group::2) loads: data lands in THIS CTA's smem, the transaction bytes are credited to the mbarrier at
// `bar_cluster_addr` - a shared::cluster address, here always the leader CTA's full barrier (see mapa_rank0)
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// shared::cluster address of the variable at the same offset in CTA 0 of the cluster
__device__ __forceinline__ uint32_t mapa_rank0(const void* local) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(local)));
  return r;
}
// Relaxed arrives: releasing a TMEM accumulator orders only tcgen05.ld traffic (wait::ld + fence::before_thread_sync
// precede it); with the default .release the compiler emits MEMBAR + ERRBAR, which parks the warp until its own global
// stores of the tile have drained (ncu source page: 10 % of all warp samples sat there).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
// CTA-pair flavours: one warp of EACH CTA of the pair executes them
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA-pair MMA, issued by the leader only: D[256 x N] = A[256 x 16] B[N x 16]^T.  Rows 0..127 of A / D live in the
// leader's smem / TMEM, rows 128..255 in the peer's (same offsets); B rows [0, N/2) in the leader's smem, [N/2, N) in the peer's.
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// same, arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFF);  // start address
  d |= uint64_t(0) << 16;                    // leading byte offset (unused: one atom along K)
  d |= uint64_t(1024 >> 4) << 32;            // stride byte offset between 8-row atoms
  d |= uint64_t(1) << 46;                    // descriptor version (sm_100)
  d |= uint64_t(2) << 61;                    // SWIZZLE_128B
  return d;
}

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }

// add 8 consecutive T values at p (16-byte aligned) onto v[0..7]
template <typename T>
__device__ __forceinline__ void add8(float* v, const T* p) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = unpack2<T>(w[i]);
    v[2 * i] += f.x;
    v[2 * i + 1] += f.y;
  }
}

template <typename T>
__device__ __forceinline__ void add8v(float* v, const uint4 u) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = unpack2<T>(w[i]);
    v[2 * i] += f.x;
    v[2 * i + 1] += f.y;
  }
}

template <int BN> struct TileCfg;
template <> struct TileCfg<256> { static constexpr int STAGES = 4; };
template <> struct TileCfg<160> { static constexpr int STAGES = 5; };  // 320, 640, 1280 = k * 160 exactly
template <> struct TileCfg<128> { static constexpr int STAGES = 6; };
template <> struct TileCfg<64> { static constexpr int STAGES = 8; };

// CTA-pair mode keeps only half of the B tile per CTA: smaller stages, deeper pipelines
template <int BN> struct PairCfg;
template <> struct PairCfg<256> { static constexpr int STAGES = 6; };
template <> struct PairCfg<160> { static constexpr int STAGES = 8; };
template <> struct PairCfg<128> { static constexpr int STAGES = 8; };
template <> struct PairCfg<64> { static constexpr int STAGES = 8; };

// MODE 1: one CTA per 128 x BN tile.  MODE 2: two CTAs per cluster on vertically adjacent tiles, B tile multicast.
// MODE 3: CTA pair, one 256 x BN tile per pair with cta_group::2 MMAs (B split between the two CTAs' smem).
template <int BN, int MODE>
constexpr int stages_of() { return MODE == 3 ? PairCfg<BN>::STAGES : TileCfg<BN>::STAGES; }
template <int BN, int MODE>
constexpr size_t smem_bytes() {
  return size_t(stages_of<BN, MODE>()) * (BM * BK * 2 + (MODE == 3 ? BN / 2 : BN) * BK * 2) + 1024 /*align slack*/ + 256 /*barriers*/;
}

// --------------------------------------------------------------------------------- the kernel
// CL = CTAs per cluster (1 or 2).  With CL = 2 the two CTAs own vertically adjacent output tiles
// (same N range): each loads its own A tile and HALF of the shared B tile, multicast into both
// CTAs' smem - L2->SM traffic per CTA-iteration drops from A+B to A+B/2 (48 KB -> 32 KB at
// BN = 256), which matters because the kernel is L2-bandwidth- rather than MMA-bound.
//
// MODE 3 (PAIR) replaces the multicast by a cta_group::2 MMA: the pair computes ONE 256 x BN tile, each CTA stages its
// own 128 rows of A and HALF of B (no duplication), only the leader (rank 0) issues MMAs, and every TMA of either CTA
// credits the leader's full barrier.  Per-CTA smem operand reads per flop drop from (128 + BN) to (128 + BN / 2) rows,
// which is what lets 160-wide tiles (8 x 160 = 1280: no wave-quantisation loss on the N = 1280 layers) run at full rate.
template <typename T, int BN, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2,
               const TcParams p) {
  constexpr int CL = MODE == 1 ? 1 : 2;
  constexpr bool PAIR = MODE == 3;
  constexpr int STAGES = stages_of<BN, MODE>();
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
  constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + size_t(STAGES) * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * (A_BYTES + B_BYTES));
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta_rank = CL > 1 ? int(cluster_ctarank()) : 0;
  const int cluster_id = int(blockIdx.x) / CL, num_clusters = int(gridDim.x) / CL;
  // work items are groups of CL vertically adjacent tiles; an odd tail gets a phantom tile whose
  // loads are out-of-bounds zeros and whose stores are masked
  const int num_groups = ((p.tiles_m + CL - 1) / CL) * p.tiles_n;
  const int k_iters = p.k_iters1 + p.k_iters2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    if (p.k_iters2) {
      prefetch_tmap(&map_a2);
      prefetch_tmap(&map_b2);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      // MODE 2: every CTA that reads this slot commits (it is written by all of them); PAIR: the leader's one commit
      mbar_init(&empty_bar[s], PAIR ? 1 : CL);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], PAIR ? 16 : 8);  // PAIR: the epilogue warps of both CTAs release the leader's MMA warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // peers' barriers exist before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================================================================= TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
        const int mg = grp / p.tiles_n, nt = grp - mg * p.tiles_n;
        const int mt = mg * CL + cta_rank;
        int a1, a2, a3;  // A coordinates of dims 1..3 at tap (0, 0)
        if (p.conv) {
          const int per_img = p.tiles_h * p.tiles_w;
          const int tb = mt / per_img, rem = mt - tb * per_img;
          const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
          a1 = tw * p.TW * p.stride - p.pad;
          a2 = th * p.TH * p.stride - p.pad;
          a3 = tb * p.TB;
        } else {
          a1 = mt * BM;
          a2 = 0;
          a3 = 0;
        }
        const int n0 = nt * BN;
        int tap = 0, kb = 0;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          void* da = smem_a + size_t(stage) * A_BYTES;
          void* db = smem_b + size_t(stage) * B_BYTES;
          if constexpr (PAIR) {
            // the leader's barrier collects both CTAs' bytes; the peer only issues its loads
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_BYTES + B_BYTES));
            const uint32_t lead_bar = mapa_rank0(&full_bar[stage]);
            if (it < p.k_iters1) {
              const int r = tap / p.kw, s = tap - r * p.kw;
              tma_load_4d_pair(da, &map_a, lead_bar, kb * BK, a1 + s, a2 + r, a3);
              tma_load_3d_pair(db, &map_b, lead_bar, kb * BK, n0 + cta_rank * (BN / 2), tap);
              if (++kb == p.kblocks) {
                kb = 0;
                ++tap;
              }
            } else {
              const int kb2 = it - p.k_iters1;
              tma_load_4d_pair(da, &map_a2, lead_bar, kb2 * BK, mt * BM, 0, 0);
              tma_load_3d_pair(db, &map_b2, lead_bar, kb2 * BK, n0 + cta_rank * (BN / 2), 0);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
          if (it < p.k_iters1) {
            const int r = tap / p.kw, s = tap - r * p.kw;
            tma_load_4d(da, &map_a, &full_bar[stage], kb * BK, a1 + s, a2 + r, a3);
            if constexpr (CL > 1) {
              // my half of the shared B tile, delivered to both CTAs
              tma_load_3d_mcast(static_cast<uint8_t*>(db) + cta_rank * (B_BYTES / CL), &map_b, &full_bar[stage], kb * BK,
                                n0 + cta_rank * (BN / CL), tap, uint16_t((1u << CL) - 1));
            } else {
              tma_load_3d(db, &map_b, &full_bar[stage], kb * BK, n0, tap);
            }
            if (++kb == p.kblocks) {
              kb = 0;
              ++tap;
            }
          } else {
            const int kb2 = it - p.k_iters1;
            tma_load_4d(da, &map_a2, &full_bar[stage], kb2 * BK, mt * BM, 0, 0);
            if constexpr (CL > 1) {
              tma_load_3d_mcast(static_cast<uint8_t*>(db) + cta_rank * (B_BYTES / CL), &map_b2, &full_bar[stage], kb2 * BK,
                                n0 + cta_rank * (BN / CL), 0, uint16_t((1u << CL) - 1));
            } else {
              tma_load_3d(db, &map_b2, &full_bar[stage], kb2 * BK, n0, 0);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ========================================================================= MMA issuer
    if (lane == 0 && (!PAIR || cta_rank == 0)) {  // PAIR: the leader drives both tensor cores
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc * BN);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase, 3);
          tcgen05_fence_after();
          const uint64_t da = make_sw128_desc(smem_u32(smem_a + size_t(stage) * A_BYTES));
          const uint64_t db = make_sw128_desc(smem_u32(smem_b + size_t(stage) * B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advancing 16 elements (32 bytes) along K inside the swizzle atom: +2 in the
            // 16-byte-granular start-address field
            if constexpr (PAIR) umma_f16_pair(tmem_d, da + uint64_t(k * 2), db + uint64_t(k * 2), p.idesc, (it > 0 || k > 0) ? 1u : 0u);
            else umma_f16(tmem_d, da + uint64_t(k * 2), db + uint64_t(k * 2), p.idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in every CTA that writes into it) when these MMAs retire
          if constexpr (PAIR) umma_commit_pair_mcast(&empty_bar[stage], uint16_t(3));
          else if constexpr (CL > 1) umma_commit_mcast(&empty_bar[stage], uint16_t((1u << CL) - 1));
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue (PAIR: of both CTAs, each reads its own 128 rows from its own TMEM)
        if constexpr (PAIR) umma_commit_pair_mcast(&tfull_bar[acc], uint16_t(3));
        else umma_commit(&tfull_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // =========================================================================== epilogue
    const int lg = warp & 3;             // TMEM lane group this warp may access
    const int half = (warp - 2) >> 2;    // warps w and w+4 share a lane group and interleave the column chunks
    const int row = lg * 32 + lane;      // row of the 128-row tile owned by this thread
    T* y = static_cast<T*>(p.y);
    const T* res = static_cast<const T*>(p.residual);
    const T* bias = static_cast<const T*>(p.bias);
    const T* cbias = static_cast<const T*>(p.chan_bias);
    const bool geglu = p.epilogue == RB200_EPI_GEGLU;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
      const int mg = grp / p.tiles_n, nt = grp - mg * p.tiles_n;
      const int mt = mg * CL + cta_rank;
      int64_t m_lin, b_idx = 0;
      bool valid;
      if (p.conv) {
        const int per_img = p.tiles_h * p.tiles_w;
        const int tb = mt / per_img, rem = mt - tb * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int ib = row / (p.TH * p.TW), irem = row - ib * (p.TH * p.TW);
        const int ih = irem / p.TW, iw = irem - ih * p.TW;
        b_idx = int64_t(tb) * p.TB + ib;
        m_lin = (b_idx * p.Ho + (int64_t(th) * p.TH + ih)) * p.Wo + (int64_t(tw) * p.TW + iw);
        valid = mt < p.tiles_m && m_lin < p.M;
      } else {
        m_lin = int64_t(mt) * BM + row;
        valid = m_lin < p.M;
      }
      // The residual rows of this tile are fetched BEFORE waiting for the accumulator, while the tensor core is
      // still working on it: the epilogue then never sits on a dependent global-load latency per column chunk.
      constexpr int MAXC = (BN / 32 + 1) / 2;
      uint4 rpre[MAXC][4];
      const bool rfast = res != nullptr && !geglu && valid && (int64_t(nt) * BN + BN <= p.N) &&
                         ((reinterpret_cast<uintptr_t>(res + m_lin * p.ldr + int64_t(nt) * BN) & 15) == 0);
      if (rfast) {
        const uint4* rs = reinterpret_cast<const uint4*>(res + m_lin * p.ldr + int64_t(nt) * BN);
#pragma unroll
        for (int cc = 0; cc < MAXC; ++cc) {
          const int c = half + 2 * cc;
          if (c < BN / 32) {
#pragma unroll
            for (int q = 0; q < 4; ++q) rpre[cc][q] = __ldg(rs + c * 4 + q);
          }
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase, 4);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + uint32_t(acc * BN) + (uint32_t(lg * 32) << 16);
#pragma unroll
      for (int cc = 0; cc < MAXC; ++cc) {
        const int c = half + 2 * cc;
        if (c >= BN / 32) break;
        const int64_t n0 = int64_t(nt) * BN + c * 32;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t raw[32];
        tmem_ld_32x32(taddr + uint32_t(c * 32), raw);
        tmem_ld_wait();
        if (!valid) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        const bool full = (n0 + 32 <= p.N);
        if (p.colscale) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || n0 + j < p.N) v[j] *= __ldg(p.colscale + n0 + j);
        }
        if (bias) {
          if (full && ((reinterpret_cast<uintptr_t>(bias + n0) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) add8<T>(v + j, bias + n0 + j);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) v[j] += to_f(bias[n0 + j]);
          }
        }
        if (cbias) {
          const T* cb = cbias + b_idx * p.N + n0;
          if (full && ((reinterpret_cast<uintptr_t>(cb) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) add8<T>(v + j, cb + j);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) v[j] += to_f(cb[j]);
          }
        }
        if (geglu) {
          // packed columns: [16 values | 16 gates]; output column block n0 / 2
          const int64_t no = n0 >> 1;
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = v[j] * gelu_erf_fast(v[16 + j]);
          T* dst = y + m_lin * p.ldy + no;
          if (res) {
            const T* rs = res + m_lin * p.ldr + no;
            if ((reinterpret_cast<uintptr_t>(rs) & 15) == 0) {
              add8<T>(o, rs);
              add8<T>(o + 8, rs + 8);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] += to_f(rs[j]);
            }
          }
          if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            uint4 u0, u1;
            u0.x = pack2<T>(o[0], o[1]);  u0.y = pack2<T>(o[2], o[3]);  u0.z = pack2<T>(o[4], o[5]);   u0.w = pack2<T>(o[6], o[7]);
            u1.x = pack2<T>(o[8], o[9]);  u1.y = pack2<T>(o[10], o[11]); u1.z = pack2<T>(o[12], o[13]); u1.w = pack2<T>(o[14], o[15]);
            reinterpret_cast<uint4*>(dst)[0] = u0;
            reinterpret_cast<uint4*>(dst)[1] = u1;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[j] = from_f<T>(o[j]);
          }
          continue;
        }
        // the kind of activation is decided ONCE per 32-column chunk: a per-element switch put a branch between the 32
        // MUFU chains (ex2, rcp), which then ran one after the other - 444 us instead of 160 us for SAM's 1280 -> 5120
        // Linear + GeLU at 16384 rows (launch list of config 5), 136 out-of-line division calls in the SASS
        if (p.epilogue == RB200_EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        } else if (p.epilogue == RB200_EPI_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_fast(v[j]);
        }
        T* dst = y + m_lin * p.ldy + n0;
        if (rfast) {
#pragma unroll
          for (int q = 0; q < 4; ++q) add8v<T>(v + 8 * q, rpre[cc][q]);
        } else if (res) {
          const T* rs = res + m_lin * p.ldr + n0;
          if (full && ((reinterpret_cast<uintptr_t>(rs) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) add8<T>(v + j, rs + j);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) v[j] += to_f(rs[j]);
          }
        }
        if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack2<T>(v[j], v[j + 1]);
            u.y = pack2<T>(v[j + 2], v[j + 3]);
            u.z = pack2<T>(v[j + 4], v[j + 5]);
            u.w = pack2<T>(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(dst + j) = u;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) dst[j] = from_f<T>(v[j]);
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): release the accumulator
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_rank0(&tempty_bar[acc]));  // the leader's MMA warp waits for both CTAs
        else mbar_arrive_relaxed(&tempty_bar[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // nobody leaves while a peer may still write/signal here
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------- host side
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

// dims/strides listed fastest first; strides in bytes for dims 1..rank-1
int make_map(CUtensorMap* map, int dtype, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, const uint32_t* estrides) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) RB200_FAIL(-4, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  const CUtensorMapDataType dt = dtype == RB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult rc = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), (const cuuint64_t*)dims, (const cuuint64_t*)strides_bytes,
                   (const cuuint32_t*)box, (const cuuint32_t*)estrides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
    RB200_FAIL(-4, "cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", int(rc), rank,
               (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
               (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return 0;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// choose the 128-pixel output tile (TW, TH, TB) of a conv; false if none fits
bool pick_conv_tile(int64_t B, int64_t Ho, int64_t Wo, int* TW, int* TH, int* TB) {
  for (int tw = 128; tw >= 1; tw >>= 1) {
    if (Wo % tw != 0) continue;
    const int rest = 128 / tw;  // = TH * TB
    if (rest <= Ho) {
      if (Ho % rest == 0) { *TW = tw; *TH = rest; *TB = 1; return true; }
      continue;
    }
    if (rest % Ho == 0) {
      const int tb = int(rest / Ho);
      if (B % tb == 0 && tb <= 256) { *TW = tw; *TH = int(Ho); *TB = tb; return true; }
    }
  }
  return false;
}

template <typename T, int BN, int CL>
int launch_tc_cl(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& ma2, const CUtensorMap& mb2,
                 TcParams& prm) {
  constexpr int MODE = CL;  // template argument: 1 single CTA, 2 multicast pair, 3 cta_group::2 pair
  constexpr int CLUSTER = MODE == 1 ? 1 : 2;
  static PerDeviceOnce configured;
  constexpr size_t SMEM = smem_bytes<BN, MODE>();
  if (MODE == 3) prm.idesc = (prm.idesc & ~(0x1Fu << 24)) | (uint32_t(256 >> 4) << 24);  // UMMA M = 256 across the pair
  if (configured.needed()) {
    if (cudaFuncSetAttribute(tc_gemm_kernel<T, BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)) != cudaSuccess)
      RB200_FAIL(-2, "tc_gemm: cannot reserve %zu bytes of shared memory", SMEM);
    configured.done();
  }
  const int64_t groups = int64_t((prm.tiles_m + CLUSTER - 1) / CLUSTER) * prm.tiles_n;
  if (groups > (int64_t(1) << 30)) RB200_FAIL(-1, "tc_gemm: too many tiles");
  const int64_t max_clusters = sm_count() / CLUSTER;
  const int clusters = int(groups < max_clusters ? groups : max_clusters);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(clusters * CLUSTER));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_gemm_kernel<T, BN, CL>, ma, mb, ma2, mb2, prm);
  if (e != cudaSuccess) RB200_FAIL(-2, "tc_gemm launch: %s", cudaGetErrorString(e));
  RB200_CHECK_LAUNCH("tc_gemm");
  return 0;
}

int cluster_mode() {
  static const int mode = [] {
    // 1: one CTA per tile; 2: CTA pairs with B multicast; 3 (default): CTA pairs with cta_group::2 MMAs
    const char* e = getenv("RB200_GEMM_CLUSTER");
    return e ? atoi(e) : 3;
  }();
  return mode;
}

template <typename T, int BN>
int launch_tc(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& ma2, const CUtensorMap& mb2,
              TcParams& prm, int cl) {
  prm.tiles_n = int(ceil_div(prm.N, BN));
  if (cl == 2) return cluster_mode() >= 3 ? launch_tc_cl<T, BN, 3>(st, ma, mb, ma2, mb2, prm) : launch_tc_cl<T, BN, 2>(st, ma, mb, ma2, mb2, prm);
  return launch_tc_cl<T, BN, 1>(st, ma, mb, ma2, mb2, prm);
}

int pick_bn(int64_t M_tiles, int64_t N) {
  static const int forced = [] {
    const char* e = getenv("RB200_GEMM_BN");  // experiments: force one tile width
    return e ? atoi(e) : 0;
  }();
  if (forced == 256 || forced == 160 || forced == 128 || forced == 64) return forced;
  const bool pair = cluster_mode() >= 3 && M_tiles >= 2;
  const int cands[4] = {256, 160, 128, 64};
  int best = 64;
  double best_cost = 1e300;
  const int sms = sm_count();
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const int64_t tn = ceil_div(N, bn);
    // time ~ waves * tile work.  Every MMA re-reads its 128 rows of A from smem whatever its width, so narrower tiles
    // move more operand bytes per flop; measured on B200 (profiles/r01_kernel_probes_bn.txt, r01_kernel_probes_pair.txt)
    // 160- / 128-wide tiles run 1.3x / 1.45x slower per flop than 256-wide ones - in CTA-pair mode as well, where the
    // B half per CTA shrinks but the A rows do not.
    const int64_t waves = pair ? ceil_div(ceil_div(M_tiles, 2) * tn, sms / 2) : ceil_div(M_tiles * tn, sms);
    const double penalty = bn == 256 ? 1.0 : (bn == 160 ? 1.3 : (bn == 128 ? 1.45 : 1.8));
    const double cost = double(waves) * (double(bn) * penalty + 24.0);
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

}  // namespace

bool tc_gemm_supported(const GemmProblem& p) {
  if (p.dtype != RB200_BF16 && p.dtype != RB200_FP16) return false;
  if (p.M <= 0 || p.N <= 0) return false;
  if (!aligned16(p.a) || !aligned16(p.b)) return false;
  if (p.epilogue == RB200_EPI_GEGLU && (p.N % 32 != 0)) return false;
  if (p.conv) {
    if (p.Cin % 8 != 0) return false;
    if (p.stride != 1 && p.stride != 2) return false;
    if (p.K2 != 0) return false;
    int tw, th, tb;
    if (!pick_conv_tile(p.B, p.Ho, p.Wo, &tw, &th, &tb)) return false;
    if (tw * p.stride > 256 || th * p.stride > 256) return false;
  } else {
    if (p.K % 8 != 0 || p.lda % 8 != 0 || p.ldb % 8 != 0) return false;
  }
  if (p.K2 != 0) {
    if (p.K2 % 64 != 0 || p.lda2 % 8 != 0 || p.ldb2 % 8 != 0 || !aligned16(p.a2) || !aligned16(p.b2)) return false;
  }
  return true;
}

int tc_gemm(cudaStream_t st, const GemmProblem& p) {
  CUtensorMap ma, mb, ma2, mb2;
  TcParams prm{};
  prm.M = p.M;
  prm.N = p.N;
  prm.bias = p.bias;
  prm.colscale = p.colscale;
  prm.chan_bias = p.chan_bias;
  prm.residual = p.residual;
  prm.ldr = p.ldr;
  prm.y = p.y;
  prm.ldy = p.ldy;
  prm.epilogue = p.epilogue;
  prm.conv = p.conv;
  const uint32_t fmt = p.dtype == RB200_BF16 ? 1u : 0u;
  const uint32_t ones[4] = {1, 1, 1, 1};

  if (p.conv) {
    int tw, th, tb;
    if (!pick_conv_tile(p.B, p.Ho, p.Wo, &tw, &th, &tb)) RB200_FAIL(-1, "tc conv: no tile for %lldx%lld", (long long)p.Ho, (long long)p.Wo);
    prm.TW = tw; prm.TH = th; prm.TB = tb;
    prm.tiles_w = int(p.Wo / tw);
    prm.tiles_h = int(p.Ho / th);
    prm.tiles_m = int((p.B / tb) * prm.tiles_h * prm.tiles_w);
    prm.Ho = p.Ho; prm.Wo = p.Wo;
    prm.stride = p.stride; prm.pad = p.pad;
    prm.kw = p.S;
    prm.kblocks = int(ceil_div(p.Cin, BK));
    prm.k_iters1 = p.R * p.S * prm.kblocks;
    const uint64_t dims[4] = {uint64_t(p.Cin), uint64_t(p.W), uint64_t(p.H), uint64_t(p.B)};
    const uint64_t strides[3] = {uint64_t(p.Cin) * 2, uint64_t(p.W) * p.Cin * 2, uint64_t(p.H) * p.W * p.Cin * 2};
    // with element strides the box is given in traversed elements: ceil(box / stride) are loaded
    const uint32_t box[4] = {uint32_t(BK), uint32_t((tw - 1) * p.stride + 1), uint32_t((th - 1) * p.stride + 1), uint32_t(tb)};
    const uint32_t es[4] = {1, uint32_t(p.stride), uint32_t(p.stride), 1};
    if (int rc = make_map(&ma, p.dtype, p.a, 4, dims, strides, box, es)) return rc;
    const uint64_t bdims[3] = {uint64_t(p.Cin), uint64_t(p.N), uint64_t(p.R * p.S)};
    const uint64_t bstr[2] = {uint64_t(p.Cin) * 2, uint64_t(p.N) * p.Cin * 2};
    const int bn = pick_bn(prm.tiles_m, p.N);
    const int cl = (cluster_mode() >= 2 && prm.tiles_m >= 2) ? 2 : 1;
    const uint32_t bbox[3] = {uint32_t(BK), uint32_t(bn / cl), 1};  // each CTA of a pair loads half of the B tile
    if (int rc = make_map(&mb, p.dtype, p.b, 3, bdims, bstr, bbox, ones)) return rc;
    ma2 = ma; mb2 = mb;
    prm.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(bn >> 3) << 17) | (uint32_t(BM >> 4) << 24);
    const bool bf = p.dtype == RB200_BF16;
    if (bn == 256) return bf ? launch_tc<__nv_bfloat16, 256>(st, ma, mb, ma2, mb2, prm, cl) : launch_tc<__half, 256>(st, ma, mb, ma2, mb2, prm, cl);
    if (bn == 160) return bf ? launch_tc<__nv_bfloat16, 160>(st, ma, mb, ma2, mb2, prm, cl) : launch_tc<__half, 160>(st, ma, mb, ma2, mb2, prm, cl);
    if (bn == 128) return bf ? launch_tc<__nv_bfloat16, 128>(st, ma, mb, ma2, mb2, prm, cl) : launch_tc<__half, 128>(st, ma, mb, ma2, mb2, prm, cl);
    return bf ? launch_tc<__nv_bfloat16, 64>(st, ma, mb, ma2, mb2, prm, cl) : launch_tc<__half, 64>(st, ma, mb, ma2, mb2, prm, cl);
  }

  prm.tiles_m = int(ceil_div(p.M, BM));
  prm.kw = 1;
  prm.kblocks = int(ceil_div(p.K, BK));
  prm.k_iters1 = prm.kblocks;
  prm.k_iters2 = p.K2 ? int(p.K2 / BK) : 0;
  prm.TW = BM; prm.TH = 1; prm.TB = 1; prm.tiles_w = 1; prm.tiles_h = 1;
  const int bn = pick_bn(prm.tiles_m, p.N);
  const int cl = (cluster_mode() >= 2 && prm.tiles_m >= 2) ? 2 : 1;
  {
    const uint64_t dims[4] = {uint64_t(p.K), uint64_t(p.M), 1, 1};
    const uint64_t strides[3] = {uint64_t(p.lda) * 2, uint64_t(p.lda) * 2 * uint64_t(p.M), uint64_t(p.lda) * 2 * uint64_t(p.M)};
    const uint32_t box[4] = {uint32_t(BK), uint32_t(BM), 1, 1};
    if (int rc = make_map(&ma, p.dtype, p.a, 4, dims, strides, box, ones)) return rc;
    const uint64_t bdims[3] = {uint64_t(p.K), uint64_t(p.N), 1};
    const uint64_t bstr[2] = {uint64_t(p.ldb) * 2, uint64_t(p.ldb) * 2 * uint64_t(p.N)};
    const uint32_t bbox[3] = {uint32_t(BK), uint32_t(bn / cl), 1};
    if (int rc = make_map(&mb, p.dtype, p.b, 3, bdims, bstr, bbox, ones)) return rc;
  }
  if (prm.k_iters2) {
    const uint64_t dims[4] = {uint64_t(p.K2), uint64_t(p.M), 1, 1};
    const uint64_t strides[3] = {uint64_t(p.lda2) * 2, uint64_t(p.lda2) * 2 * uint64_t(p.M), uint64_t(p.lda2) * 2 * uint64_t(p.M)};
    const uint32_t box[4] = {uint32_t(BK), uint32_t(BM), 1, 1};
    if (int rc = make_map(&ma2, p.dtype, p.a2, 4, dims, strides, box, ones)) return rc;
    const uint64_t bdims[3] = {uint64_t(p.K2), uint64_t(p.N), 1};
    const uint64_t bstr[2] = {uint64_t(p.ldb2) * 2, uint64_t(p.ldb2) * 2 * uint64_t(p.N)};
    const uint32_t bbox[3] = {uint32_t(BK), uint32_t(bn / cl), 1};
    if (int rc = make_map(&mb2, p.dtype, p.b2, 3, bdims, bstr, bbox, ones)) return rc;
  } else {
    ma2 = ma; mb2 = mb;
  }
  prm.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(bn >> 3) << 17) | (uint32_t(BM >> 4) << 24);
  const bool bf = p.dtype == RB200_BF16;
  if (bn == 256) return bf ? launch_tc<__nv_bfloat16, 256>(st, ma, mb, ma2, mb2, prm, cl) : launch_tc<__half, 256>(st, ma, mb, ma2, mb2, prm, cl);
  if (bn == 160) return bf ? launch_tc<__nv_bfloat16, 160>(st, ma, mb, ma2, mb2, prm, cl) : launch_tc<__half, 160>(st, ma, mb, ma2, mb2, prm, cl);
  if (bn == 128) return bf ? launch_tc<__nv_bfloat16, 128>(st, ma, mb, ma2, mb2, prm, cl) : launch_tc<__half, 128>(st, ma, mb, ma2, mb2, prm, cl);
  return bf ? launch_tc<__nv_bfloat16, 64>(st, ma, mb, ma2, mb2, prm, cl) : launch_tc<__half, 64>(st, ma, mb, ma2, mb2, prm, cl);
}

}  // namespace rb200
