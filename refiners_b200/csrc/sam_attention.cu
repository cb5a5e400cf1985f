// SAM decomposed relative-position attention
// (foundationals/segment_anything/image_encoder.py:87-143 in the reference):
//   logits[q, (kh, kw)] = (q . k) d^-1/2 + q . R_v[h - kh + H - 1] + q . R_h[w - kw + W - 1]
// Step 1 (this file): the two rank-H / rank-W bias tables per query, fp32, into the workspace:
//   bias_h[bw, head, kh, q] = q . rel_h_emb[h - kh + H - 1],  bias_w[bw, head, kw, q] = q . rel_w_emb[w - kw + W - 1]
// Step 2: flash attention with the bias added to the scores (never materialising HW x HW logits).
#include "common.cuh"

namespace rb200 {
namespace {

// One block per (bw, head, query row h): the Ww queries of that row, the Hh vertical and 2 Ww - 1 horizontal
// embedding rows they need, all staged in shared memory as fp32 with an odd row pitch (bank-conflict free when the
// lanes of a warp walk consecutive rows).  Every thread owns one query w and four table entries k.
// Output tables are [bw, head, k, q] (q fastest) so that both this kernel's stores and the attention kernel's loads
// (one query row per thread) are coalesced.
template <typename T>
__global__ void __launch_bounds__(256) rel_bias_kernel(const T* __restrict__ qkv, const T* __restrict__ rel_h, const T* __restrict__ rel_w,
                                                       float* __restrict__ bias_h, float* __restrict__ bias_w, int Hh, int Ww,
                                                       int heads, int d) {
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
    const int q = h * Ww + w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + i;
      if (k < Hh)
        bias_h[(plane * Hh + k) * HW + q] = acc[i];
      else if (k < K)
        bias_w[(plane * Ww + (k - Hh)) * HW + q] = acc[i];
    }
  }
}

template <typename T>
int launch_rel_bias(cudaStream_t st, const void* qkv, const void* rel_h, const void* rel_w, float* bias_h, float* bias_w, int64_t Bw,
                    int Hh, int Ww, int heads, int d) {
  const size_t smem = size_t(3 * Ww - 1 + Hh) * (d | 1) * sizeof(float);
  if (smem > 200 * 1024) RB200_FAIL(-1, "sam_attention: window %dx%d with head dim %d does not fit shared memory", Hh, Ww, d);
  if (smem > 48 * 1024) cudaFuncSetAttribute(rel_bias_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  const int work = Ww * ((Hh + Ww + 3) / 4);
  const int threads = work >= 256 ? 256 : (work + 31) / 32 * 32;
  rel_bias_kernel<T><<<unsigned(Bw * heads * Hh), threads, smem, st>>>(static_cast<const T*>(qkv), static_cast<const T*>(rel_h),
                                                                        static_cast<const T*>(rel_w), bias_h, bias_w, Hh, Ww, heads, d);
  RB200_CHECK_LAUNCH("sam_rel_bias");
  return 0;
}

}  // namespace

size_t sam_attention_ws(int64_t Bw, int Hh, int Ww, int heads, int d) {
  (void)d;
  return size_t(Bw) * heads * Hh * Ww * (size_t(Hh) + Ww) * sizeof(float) + 512;
}

int sam_attention_impl(cudaStream_t st, int dtype, const void* qkv, const void* rel_h_emb, const void* rel_w_emb, void* o,
                       int64_t Bw, int Hh, int Ww, int heads, int d, void* ws, size_t ws_bytes) {
  if (Hh < 1 || Ww < 1 || heads < 1 || d < 1 || d > 256) RB200_FAIL(-1, "sam_attention: bad geometry");
  const size_t need = sam_attention_ws(Bw, Hh, Ww, heads, d);
  if (!ws || ws_bytes < need) RB200_FAIL(-1, "sam_attention: workspace %zu < %zu", ws_bytes, need);
  const int64_t HW = int64_t(Hh) * Ww;
  const int64_t nq = Bw * heads * HW;
  if (nq > 2147483647LL) RB200_FAIL(-1, "sam_attention: too many queries for one launch");
  float* bias_h = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  float* bias_w = bias_h + nq * Hh;
  int rc = 0;
  switch (dtype) {
    case RB200_BF16: rc = launch_rel_bias<__nv_bfloat16>(st, qkv, rel_h_emb, rel_w_emb, bias_h, bias_w, Bw, Hh, Ww, heads, d); break;
    case RB200_FP16: rc = launch_rel_bias<__half>(st, qkv, rel_h_emb, rel_w_emb, bias_h, bias_w, Bw, Hh, Ww, heads, d); break;
    case RB200_FP32: rc = launch_rel_bias<float>(st, qkv, rel_h_emb, rel_w_emb, bias_h, bias_w, Bw, Hh, Ww, heads, d); break;
    default: RB200_FAIL(-1, "sam_attention: bad dtype %d", dtype);
  }
  if (rc) return rc;
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
  p.bias_h = bias_h; p.bias_w = bias_w; p.bias_H = Hh; p.bias_W = Ww;
  // tcgen05 path: head dim 80 runs as two 64-column slabs (TMA zero-fills columns 80..127)
  if (kernel_mode() != 1 && tc_sdpa_supported(p)) return tc_sdpa(st, p);
  return simt_sdpa(st, p);
}

}  // namespace rb200
