// SAM decomposed relative-position attention
// (foundationals/segment_anything/image_encoder.py:87-143 in the reference):
//   logits[q, (kh, kw)] = (q . k) d^-1/2 + q . R_v[h - kh + H - 1] + q . R_h[w - kw + W - 1]
// Step 1 (this file): the two rank-H / rank-W bias tables per query, fp32, into the workspace:
//   bias[bw, head, q / 128, kh, q % 128]      = q . rel_h_emb[h - kh + H - 1]
//   bias[bw, head, q / 128, H + kw, q % 128]  = q . rel_w_emb[w - kw + W - 1]
// Step 2: flash attention with the bias added to the scores (never materialising HW x HW logits).
#include "common.cuh"

namespace rb200 {
namespace {

// One block per (bw, head, query row h): the Ww queries of that row, the Hh vertical and 2 Ww - 1 horizontal
// embedding rows they need, all staged in shared memory as fp32 with an odd row pitch (bank-conflict free when the
// lanes of a warp walk consecutive rows).  Every thread owns one query w and four table entries k.
// Output: one contiguous block of (Hh + Ww) x 128 floats per (bw, head, 128-query tile), [k][q % 128]: this kernel's
// stores are coalesced and the attention kernel stages the block of a work item with a single bulk copy.
template <typename T>
__global__ void __launch_bounds__(256) rel_bias_kernel(const T* __restrict__ qkv, const T* __restrict__ rel_h, const T* __restrict__ rel_w,
                                                       float* __restrict__ bias, int Hh, int Ww, int heads, int d) {
  extern __shared__ float sm[];
  const int pitch = d | 1;
  float* qs = sm;                       // [Ww][pitch]
  float* eh = qs + Ww * pitch;          // [Hh][pitch]      row kh  = rel_h[h - kh + Hh - 1]
  float* ew = eh + Hh * pitch;          // [2 Ww - 1][pitch]
  const int h = blockIdx.x % Hh;
  const int head = (blockIdx.x / Hh) % heads;
  const int64_t bw = blockIdx.x / (Hh * heads);
  const int HW = Hh * Ww, C = heads * d;
  const T* qbase = qkv + (bw * HW + int64_t(h) * Ww) * 3 * C + head * d;
  for (int i = threadIdx.x; i < Ww * d; i += blockDim.x) qs[(i / d) * pitch + i % d] = to_f(qbase[int64_t(i / d) * 3 * C + i % d]);
  for (int i = threadIdx.x; i < Hh * d; i += blockDim.x) eh[(i / d) * pitch + i % d] = to_f(rel_h[int64_t(h - i / d + Hh - 1) * d + i % d]);
  for (int i = threadIdx.x; i < (2 * Ww - 1) * d; i += blockDim.x) ew[(i / d) * pitch + i % d] = to_f(rel_w[i]);
  __syncthreads();
  const int K = Hh + Ww;           // combined table index: [0, Hh) vertical, [Hh, Hh + Ww) horizontal
  const int kgroups = (K + 3) / 4;
  const int64_t plane = (bw * heads + head);
  for (int o = threadIdx.x; o < Ww * kgroups; o += blockDim.x) {
    const int w = o % Ww, k0 = (o / Ww) * 4;
    const float* e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + i < K ? k0 + i : K - 1;
      e[i] = k < Hh ? eh + k * pitch : ew + (w - (k - Hh) + Ww - 1) * pitch;
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* qrow = qs + w * pitch;
    for (int c = 0; c < d; ++c) {
      const float qv = qrow[c];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(qv, e[i][c], acc[i]);
    }
    const int q = h * Ww + w, nqt = (HW + 127) / 128;
    float* out = bias + ((plane * nqt + q / 128) * K) * 128 + q % 128;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (k0 + i < K) out[(k0 + i) * 128] = acc[i];
  }
}

template <typename T>
int launch_rel_bias(cudaStream_t st, const void* qkv, const void* rel_h, const void* rel_w, float* bias, int64_t Bw, int Hh, int Ww,
                    int heads, int d) {
  const size_t smem = size_t(3 * Ww - 1 + Hh) * (d | 1) * sizeof(float);
  if (smem > 200 * 1024) RB200_FAIL(-1, "sam_attention: window %dx%d with head dim %d does not fit shared memory", Hh, Ww, d);
  if (smem > 48 * 1024) cudaFuncSetAttribute(rel_bias_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  const int work = Ww * ((Hh + Ww + 3) / 4);
  const int threads = work >= 256 ? 256 : (work + 31) / 32 * 32;
  rel_bias_kernel<T><<<unsigned(Bw * heads * Hh), threads, smem, st>>>(static_cast<const T*>(qkv), static_cast<const T*>(rel_h),
                                                                        static_cast<const T*>(rel_w), bias, Hh, Ww, heads, d);
  RB200_CHECK_LAUNCH("sam_rel_bias");
  return 0;
}

// ---- warp-MMA version (bf16 / fp16 operands, fp32 accumulate) -------------------------------------------------------
// The same products as rel_bias_kernel, as [Ww x d] x [d x (Hh + 2 Ww - 1)] per (bw, head, query row h) on
// mma.sync.m16n8k16: the vertical part multiplies the Hh table rows this query row needs, the horizontal part
// multiplies the WHOLE horizontal table and each accumulator is scattered to its (w, kw = w - j + Ww - 1) slot.
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2], __nv_bfloat16) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2], __half) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <typename T>
__global__ void __launch_bounds__(256) rel_bias_mma_kernel(const T* __restrict__ qkv, const T* __restrict__ rel_h, const T* __restrict__ rel_w,
                                                           float* __restrict__ bias, int Hh, int Ww, int heads, int d) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int dk = (d + 15) & ~15;       // K extent, zero padded to the MMA step
  const int pitch = dk + 8;            // elements; (pitch / 2) % 8 == 4 keeps the fragment loads conflict free
  const int Mp = (Ww + 15) & ~15, Nh = (Hh + 7) & ~7, Nw = (2 * Ww - 1 + 7) & ~7;
  T* qs = reinterpret_cast<T*>(smem_raw);  // [Mp][pitch]
  T* eh = qs + Mp * pitch;                 // [Nh][pitch]  row kh = rel_h[h - kh + Hh - 1]
  T* ew = eh + Nh * pitch;                 // [Nw][pitch]  the whole horizontal table
  const int h = blockIdx.x % Hh;
  const int head = (blockIdx.x / Hh) % heads;
  const int64_t bw = blockIdx.x / (Hh * heads);
  const int HW = Hh * Ww, C = heads * d;
  const int cv = pitch / 8;  // 16-byte chunks per smem row
  const int dv = d / 8;      // ... holding data (d % 8 == 0)
  const T* qbase = qkv + (bw * HW + int64_t(h) * Ww) * 3 * C + head * d;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < (Mp + Nh + Nw) * cv; i += blockDim.x) {
    const int r = i / cv, c = i % cv;
    uint4 val = zero;
    if (c < dv) {
      if (r < Mp) {
        if (r < Ww) val = *reinterpret_cast<const uint4*>(qbase + int64_t(r) * 3 * C + c * 8);
      } else if (r < Mp + Nh) {
        const int kh = r - Mp;
        if (kh < Hh) val = *reinterpret_cast<const uint4*>(rel_h + int64_t(h - kh + Hh - 1) * d + c * 8);
      } else {
        const int j = r - Mp - Nh;
        if (j < 2 * Ww - 1) val = *reinterpret_cast<const uint4*>(rel_w + int64_t(j) * d + c * 8);
      }
    }
    *reinterpret_cast<uint4*>(qs + r * pitch + c * 8) = val;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int MT = Mp / 16, NTh = Nh / 8, NT = NTh + Nw / 8;
  const int K = Hh + Ww, nqt = (HW + 127) / 128;
  const int64_t plane = bw * heads + head;
  for (int tile = warp; tile < MT * NT; tile += nwarps) {
    const int mt = tile % MT, nt = tile / MT;
    const T* arow = qs + (mt * 16 + g) * pitch + 2 * t;
    const T* brow = (nt < NTh ? eh + (nt * 8 + g) * pitch : ew + ((nt - NTh) * 8 + g) * pitch) + 2 * t;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < dk; k0 += 16) {
      uint32_t a[4], b[2];
      a[0] = *reinterpret_cast<const uint32_t*>(arow + k0);
      a[1] = *reinterpret_cast<const uint32_t*>(arow + 8 * pitch + k0);
      a[2] = *reinterpret_cast<const uint32_t*>(arow + k0 + 8);
      a[3] = *reinterpret_cast<const uint32_t*>(arow + 8 * pitch + k0 + 8);
      b[0] = *reinterpret_cast<const uint32_t*>(brow + k0);
      b[1] = *reinterpret_cast<const uint32_t*>(brow + k0 + 8);
      mma_16816(c, a, b, T());
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int w = mt * 16 + g + (e >> 1) * 8;
      const int n = nt * 8 + 2 * t + (e & 1);
      if (w >= Ww) continue;
      const int q = h * Ww + w;
      float* out = bias + ((plane * nqt + q / 128) * K) * 128 + q % 128;
      if (nt < NTh) {
        if (n < Hh) out[n * 128] = c[e];
      } else {
        const int j = n - Nh;
        const int kw = w - j + Ww - 1;
        if (kw >= 0 && kw < Ww) out[(Hh + kw) * 128] = c[e];  // j < 2 Ww - 1 follows from kw >= 0
      }
    }
  }
}

template <typename T>
int launch_rel_bias_mma(cudaStream_t st, const void* qkv, const void* rel_h, const void* rel_w, float* bias, int64_t Bw, int Hh, int Ww,
                        int heads, int d) {
  const int dk = (d + 15) & ~15, pitch = dk + 8;
  const int Mp = (Ww + 15) & ~15, Nh = (Hh + 7) & ~7, Nw = (2 * Ww - 1 + 7) & ~7;
  const size_t smem = size_t(Mp + Nh + Nw) * pitch * sizeof(T);
  if (smem > 48 * 1024) cudaFuncSetAttribute(rel_bias_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  const int tiles = (Mp / 16) * ((Nh + Nw) / 8);
  const int threads = tiles >= 8 ? 256 : 32 * tiles;
  rel_bias_mma_kernel<T><<<unsigned(Bw * heads * Hh), threads, smem, st>>>(static_cast<const T*>(qkv), static_cast<const T*>(rel_h),
                                                                            static_cast<const T*>(rel_w), bias, Hh, Ww, heads, d);
  RB200_CHECK_LAUNCH("sam_rel_bias_mma");
  return 0;
}

bool rel_bias_mma_ok(const void* qkv, const void* rel_h, const void* rel_w, int Hh, int Ww, int d, size_t esz) {
  const int dk = (d + 15) & ~15, pitch = dk + 8;
  const size_t smem = size_t(((Ww + 15) & ~15) + ((Hh + 7) & ~7) + ((2 * Ww - 1 + 7) & ~7)) * pitch * esz;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return d % 8 == 0 && smem <= 200 * 1024 && al(qkv) && al(rel_h) && al(rel_w);
}

}  // namespace

size_t sam_attention_ws(int64_t Bw, int Hh, int Ww, int heads, int d) {
  (void)d;
  const size_t nqt = (size_t(Hh) * Ww + 127) / 128;
  return size_t(Bw) * heads * nqt * (size_t(Hh) + Ww) * 128 * sizeof(float) + 512;
}

int sam_attention_impl(cudaStream_t st, int dtype, const void* qkv, const void* rel_h_emb, const void* rel_w_emb, void* o,
                       int64_t Bw, int Hh, int Ww, int heads, int d, void* ws, size_t ws_bytes) {
  if (Hh < 1 || Ww < 1 || heads < 1 || d < 1 || d > 256) RB200_FAIL(-1, "sam_attention: bad geometry");
  const size_t need = sam_attention_ws(Bw, Hh, Ww, heads, d);
  if (!ws || ws_bytes < need) RB200_FAIL(-1, "sam_attention: workspace %zu < %zu", ws_bytes, need);
  const int64_t HW = int64_t(Hh) * Ww;
  const int64_t nq = Bw * heads * HW;
  if (nq > 2147483647LL) RB200_FAIL(-1, "sam_attention: too many queries for one launch");
  float* bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  const int64_t C = int64_t(heads) * d;
  const size_t esz = dtype_size(dtype);
  SdpaProblem p{};
  p.dtype = dtype;
  p.q = qkv;
  p.k = static_cast<const char*>(qkv) + C * esz;
  p.v = static_cast<const char*>(qkv) + 2 * C * esz;
  p.o = o;
  p.B = Bw; p.H = heads; p.Sq = HW; p.Sk = HW; p.D = d;
  p.q_sb = p.k_sb = p.v_sb = HW * 3 * C;
  p.q_ss = p.k_ss = p.v_ss = 3 * C;
  p.o_sb = HW * C; p.o_ss = C;
  p.scale = 1.0f / sqrtf(float(d));
  p.bias_h = bias; p.bias_w = nullptr; p.bias_H = Hh; p.bias_W = Ww;  // combined table, see rel_bias_kernel
  p.rel_h_emb = rel_h_emb; p.rel_w_emb = rel_w_emb;
  // SAM's 14 x 14 windows: the attention kernel multiplies the queries with the embeddings itself
  if (kernel_mode() != 1 && tc_sdpa_win_fuses_bias(p)) return tc_sdpa_win(st, p);
  int rc = 0;
  const bool mma = kernel_mode() != 1 && dtype != RB200_FP32 && rel_bias_mma_ok(qkv, rel_h_emb, rel_w_emb, Hh, Ww, d, 2);
  switch (dtype) {
    case RB200_BF16:
      rc = mma ? launch_rel_bias_mma<__nv_bfloat16>(st, qkv, rel_h_emb, rel_w_emb, bias, Bw, Hh, Ww, heads, d)
               : launch_rel_bias<__nv_bfloat16>(st, qkv, rel_h_emb, rel_w_emb, bias, Bw, Hh, Ww, heads, d);
      break;
    case RB200_FP16:
      rc = mma ? launch_rel_bias_mma<__half>(st, qkv, rel_h_emb, rel_w_emb, bias, Bw, Hh, Ww, heads, d)
               : launch_rel_bias<__half>(st, qkv, rel_h_emb, rel_w_emb, bias, Bw, Hh, Ww, heads, d);
      break;
    case RB200_FP32: rc = launch_rel_bias<float>(st, qkv, rel_h_emb, rel_w_emb, bias, Bw, Hh, Ww, heads, d); break;
    default: RB200_FAIL(-1, "sam_attention: bad dtype %d", dtype);
  }
  if (rc) return rc;
  // tcgen05 path: head dim 80 runs as two 64-column slabs (TMA zero-fills columns 80..127)
  if (kernel_mode() != 1 && tc_sdpa_win_supported(p)) return tc_sdpa_win(st, p);   // the 14 x 14 windows
  if (kernel_mode() != 1 && tc_sdpa_supported(p)) return tc_sdpa(st, p);
  return simt_sdpa(st, p);
}

}  // namespace rb200
