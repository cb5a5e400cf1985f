// SAM decomposed relative-position attention
// (foundationals/segment_anything/image_encoder.py:87-143 in the reference):
//   logits[q, (kh, kw)] = (q . k) d^-1/2 + q . R_v[h - kh + H - 1] + q . R_h[w - kw + W - 1]
// Step 1 (this file): the two rank-H / rank-W bias tables per query, fp32, into the workspace:
//   bias_h[bw, head, q, kh] = q . rel_h_emb[h - kh + H - 1],  bias_w[bw, head, q, kw] = q . rel_w_emb[w - kw + W - 1]
// Step 2: flash attention with the bias added to the scores (never materialising HW x HW logits).
#include "common.cuh"

namespace rb200 {
namespace {

template <typename T>
__global__ void rel_bias_kernel(const T* __restrict__ qkv, const T* __restrict__ rel_h, const T* __restrict__ rel_w,
                                float* __restrict__ bias_h, float* __restrict__ bias_w, int64_t Bw, int Hh, int Ww, int heads,
                                int d) {
  // one block per (bw, head, query); threads over the Hh + Ww outputs
  extern __shared__ float qs[];
  const int64_t idx = blockIdx.x;
  const int HW = Hh * Ww;
  const int qi = int(idx % HW);
  const int head = int((idx / HW) % heads);
  const int64_t bw = idx / (int64_t(HW) * heads);
  const int C = heads * d;
  const T* q = qkv + (bw * HW + qi) * 3 * C + head * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) qs[c] = to_f(q[c]);
  __syncthreads();
  const int h = qi / Ww, w = qi % Ww;
  for (int t = threadIdx.x; t < Hh + Ww; t += blockDim.x) {
    const T* e = (t < Hh) ? rel_h + int64_t(h - t + Hh - 1) * d : rel_w + int64_t(w - (t - Hh) + Ww - 1) * d;
    float acc = 0.f;
    for (int c = 0; c < d; ++c) acc = fmaf(qs[c], to_f(e[c]), acc);
    if (t < Hh)
      bias_h[idx * Hh + t] = acc;
    else
      bias_w[idx * Ww + (t - Hh)] = acc;
  }
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
  const int threads = (Hh + Ww) <= 64 ? 64 : 128;
  switch (dtype) {
    case RB200_BF16:
      rel_bias_kernel<__nv_bfloat16><<<unsigned(nq), threads, d * sizeof(float), st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)rel_h_emb, (const __nv_bfloat16*)rel_w_emb, bias_h, bias_w, Bw, Hh, Ww, heads, d);
      break;
    case RB200_FP16:
      rel_bias_kernel<__half><<<unsigned(nq), threads, d * sizeof(float), st>>>((const __half*)qkv, (const __half*)rel_h_emb, (const __half*)rel_w_emb, bias_h, bias_w, Bw, Hh, Ww, heads, d);
      break;
    case RB200_FP32:
      rel_bias_kernel<float><<<unsigned(nq), threads, d * sizeof(float), st>>>((const float*)qkv, (const float*)rel_h_emb, (const float*)rel_w_emb, bias_h, bias_w, Bw, Hh, Ww, heads, d);
      break;
    default: RB200_FAIL(-1, "sam_attention: bad dtype %d", dtype);
  }
  RB200_CHECK_LAUNCH("sam_rel_bias");
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
