// Attention PROBABILITIES  softmax(q k^T * scale)  materialised as [B, H, Sq, Sk].
//
// The flash kernels never form this matrix; self-attention guidance needs it for ONE layer per UNet pass (the first
// self-attention of the middle block): /root/reference/src/refiners/foundationals/latent_diffusion/
// self_attention_guidance.py:42-47 (SelfAttentionMap.compute_attention_scores: q @ k^T / sqrt(d), softmax).
//
// One CTA = QR query rows of one (batch, head): the scaled logit rows are built in shared memory (fp32) from K rows
// streamed once through registers, then each warp normalises whole rows and writes them with coalesced stores.  The
// output (B*H*Sq*Sk elements, 0.67 GB for SDXL's 32x32 middle block at UNet batch 16) is the algorithmic traffic:
// this kernel is HBM-write bound; Q and K are re-read from L2 (K: Sq/QR times per head).
#include "common.cuh"

namespace rb200 {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxD = 256;

template <typename T> struct Chunk;  // 16 bytes of T as floats
template <> struct Chunk<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* f) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
};
template <> struct Chunk<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <> struct Chunk<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __half* p, float* f) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};

template <typename T, int QR>
__global__ void __launch_bounds__(kThreads) attn_probs_kernel(const T* __restrict__ q, const T* __restrict__ k, T* __restrict__ out, int H,
                                                              int64_t Sq, int64_t Sk, int D, int64_t q_sb, int64_t q_ss, int64_t k_sb,
                                                              int64_t k_ss, float scale) {
  extern __shared__ float smem[];
  float* qs = smem;                 // [QR][D]
  float* logits = smem + QR * D;    // [QR][Sk]
  const int64_t b = blockIdx.z;
  const int h = blockIdx.y;
  const int64_t q0 = int64_t(blockIdx.x) * QR;
  const int rows = (Sq - q0) < QR ? int(Sq - q0) : QR;
  constexpr int N = Chunk<T>::N;

  for (int i = threadIdx.x; i < QR * (D / N); i += kThreads) {
    const int r = i / (D / N), c = i - r * (D / N);
    float f[N];
    if (r < rows) {
      Chunk<T>::load(q + b * q_sb + (q0 + r) * q_ss + int64_t(h) * D + c * N, f);
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) f[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) qs[r * D + c * N + j] = f[j];
  }
  __syncthreads();

  const T* kb = k + b * k_sb + int64_t(h) * D;
  for (int64_t j = threadIdx.x; j < Sk; j += kThreads) {
    float acc[QR];
#pragma unroll
    for (int r = 0; r < QR; ++r) acc[r] = 0.f;
    const T* kr = kb + j * k_ss;
    for (int c = 0; c < D; c += N) {
      float f[N];
      Chunk<T>::load(kr + c, f);
#pragma unroll
      for (int r = 0; r < QR; ++r) {
#pragma unroll
        for (int e = 0; e < N; e += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(qs + r * D + c + e);  // broadcast read
          acc[r] = fmaf(f[e], qv.x, acc[r]);
          acc[r] = fmaf(f[e + 1], qv.y, acc[r]);
          acc[r] = fmaf(f[e + 2], qv.z, acc[r]);
          acc[r] = fmaf(f[e + 3], qv.w, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < QR; ++r) logits[int64_t(r) * Sk + j] = acc[r] * scale;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < rows; r += kThreads / 32) {
    float* row = logits + int64_t(r) * Sk;
    float m = -INFINITY;
    for (int64_t j = lane; j < Sk; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int64_t j = lane; j < Sk; j += 32) {
      const float e = __expf(row[j] - m);
      row[j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    T* dst = out + ((b * H + h) * Sq + q0 + r) * Sk;
    for (int64_t j = lane; j < Sk; j += 32) dst[j] = from_f<T>(row[j] * inv);
  }
}

template <typename T, int QR>
int launch(cudaStream_t st, const void* q, const void* k, void* out, int64_t B, int H, int64_t Sq, int64_t Sk, int D, int64_t q_sb,
           int64_t q_ss, int64_t k_sb, int64_t k_ss, float scale) {
  const size_t smem = (size_t(QR) * D + size_t(QR) * Sk) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(attn_probs_kernel<T, QR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) RB200_FAIL(-2, "attention_probs: cannot reserve shared memory: %s", cudaGetErrorString(e));
  dim3 grid(unsigned(ceil_div(Sq, QR)), unsigned(H), unsigned(B));
  attn_probs_kernel<T, QR><<<grid, kThreads, smem, st>>>(static_cast<const T*>(q), static_cast<const T*>(k), static_cast<T*>(out), H, Sq, Sk,
                                                         D, q_sb, q_ss, k_sb, k_ss, scale);
  RB200_CHECK_LAUNCH("attention_probs");
  return 0;
}

template <typename T>
int dispatch(cudaStream_t st, const void* q, const void* k, void* out, int64_t B, int H, int64_t Sq, int64_t Sk, int D, int64_t q_sb,
             int64_t q_ss, int64_t k_sb, int64_t k_ss, float scale) {
  const size_t budget = 200 * 1024;
  auto fits = [&](int qr) { return (size_t(qr) * D + size_t(qr) * Sk) * sizeof(float) <= budget; };
#define RB200_PROBS(QR) \
  if (fits(QR)) return launch<T, QR>(st, q, k, out, B, H, Sq, Sk, D, q_sb, q_ss, k_sb, k_ss, scale)
  RB200_PROBS(16);
  RB200_PROBS(8);
  RB200_PROBS(4);
  RB200_PROBS(2);
  RB200_PROBS(1);
#undef RB200_PROBS
  RB200_FAIL(-1, "attention_probs: a logit row of %lld keys does not fit in shared memory", (long long)Sk);
}

}  // namespace

int attention_probs_impl(cudaStream_t st, int dtype, const void* q, const void* k, void* out, int64_t B, int H, int64_t Sq, int64_t Sk, int D,
                         int64_t q_sb, int64_t q_ss, int64_t k_sb, int64_t k_ss, float scale) {
  const int n = dtype == RB200_FP32 ? 4 : 8;
  if (D % n != 0 || D > kMaxD) RB200_FAIL(-1, "attention_probs: head dim %d must be a multiple of %d and at most %d", D, n, kMaxD);
  const size_t es = dtype_size(dtype);
  if ((q_sb * es) % 16 || (q_ss * es) % 16 || (k_sb * es) % 16 || (k_ss * es) % 16 || (reinterpret_cast<uintptr_t>(q) & 15) ||
      (reinterpret_cast<uintptr_t>(k) & 15))
    RB200_FAIL(-1, "attention_probs: q and k rows must be 16-byte aligned");
  if (H > 65535 || B > 65535) RB200_FAIL(-1, "attention_probs: batch / heads exceed the grid limits");
  switch (dtype) {
    case RB200_BF16: return dispatch<__nv_bfloat16>(st, q, k, out, B, H, Sq, Sk, D, q_sb, q_ss, k_sb, k_ss, scale);
    case RB200_FP16: return dispatch<__half>(st, q, k, out, B, H, Sq, Sk, D, q_sb, q_ss, k_sb, k_ss, scale);
    default: return dispatch<float>(st, q, k, out, B, H, Sq, Sk, D, q_sb, q_ss, k_sb, k_ss, scale);
  }
}

}  // namespace rb200
