// Generic CUDA-core GEMM / implicit-GEMM conv: the correctness anchor and the path for
// shapes and dtypes the tcgen05 kernel does not take (fp32, K % 8 != 0, tiny Cin ...).
// Same epilogue semantics as the tensor-core kernel (see include/refiners_b200.h):
//   v = acc * colscale[n] + bias[n] + chan_bias[b, n];  v = act(v);  v += residual[m, n]
#include "common.cuh"

namespace rb200 {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct ConvGeom {
  int64_t H, W, Cin, Ho, Wo;
  int S, stride, pad;
};

template <typename T, bool CONV>
__device__ __forceinline__ float load_a(const GemmProblem& p, int64_t m, int64_t k) {
  if (m >= p.M) return 0.f;
  if (k < p.K) {
    if constexpr (CONV) {
      const int64_t hw = p.Ho * p.Wo;
      const int64_t b = m / hw, rem = m - b * hw;
      const int64_t ho = rem / p.Wo, wo = rem - ho * p.Wo;
      const int64_t tap = k / p.Cin, c = k - tap * p.Cin;
      const int r = int(tap / p.S), s = int(tap - int64_t(r) * p.S);
      const int64_t hi = ho * p.stride - p.pad + r, wi = wo * p.stride - p.pad + s;
      if (hi < 0 || hi >= p.H || wi < 0 || wi >= p.W) return 0.f;
      return to_f(static_cast<const T*>(p.a)[((b * p.H + hi) * p.W + wi) * p.Cin + c]);
    } else {
      return to_f(static_cast<const T*>(p.a)[m * p.lda + k]);
    }
  }
  const int64_t k2 = k - p.K;
  if (k2 < p.K2) return to_f(static_cast<const T*>(p.a2)[m * p.lda2 + k2]);
  return 0.f;
}

template <typename T, bool CONV>
__device__ __forceinline__ float load_b(const GemmProblem& p, int64_t n, int64_t k) {
  if (n >= p.N) return 0.f;
  if (k < p.K) {
    if constexpr (CONV) {
      const int64_t tap = k / p.Cin, c = k - tap * p.Cin;
      return to_f(static_cast<const T*>(p.b)[(tap * p.N + n) * p.Cin + c]);
    } else {
      return to_f(static_cast<const T*>(p.b)[n * p.ldb + k]);
    }
  }
  const int64_t k2 = k - p.K;
  if (k2 < p.K2) return to_f(static_cast<const T*>(p.b2)[n * p.ldb2 + k2]);
  return 0.f;
}

template <typename T, bool CONV>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const GemmProblem p) {
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = int64_t(blockIdx.x) * TM, n0 = int64_t(blockIdx.y) * TN;  // row tiles on x: no 65535 limit
  // columns owned by this thread inside the tile: a value pair and its GEGLU gate pair
  const int cg = (tx >> 3) * 32, cj = (tx & 7) * 2;
  const int cols[4] = {cg + cj, cg + cj + 1, cg + 16 + cj, cg + 16 + cj + 1};

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t Ktot = p.K + p.K2;
  for (int64_t k0 = 0; k0 < Ktot; k0 += TK) {
    for (int e = tid; e < TM * TK; e += 256) {
      const int kk = e % TK, mm = e / TK;
      As[kk][mm] = load_a<T, CONV>(p, m0 + mm, k0 + kk);
    }
    for (int e = tid; e < TN * TK; e += 256) {
      const int kk = e % TK, nn = e / TK;
      Bs[kk][nn] = load_b<T, CONV>(p, n0 + nn, k0 + kk);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][cols[j]];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  T* y = static_cast<T*>(p.y);
  const T* res = static_cast<const T*>(p.residual);
  const T* bias = static_cast<const T*>(p.bias);
  const T* cb = static_cast<const T*>(p.chan_bias);
  const int64_t hw = CONV ? p.Ho * p.Wo : 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + cols[j];
      float t = acc[i][j];
      if (n < p.N) {
        if (p.colscale) t *= p.colscale[n];
        if (bias) t += to_f(bias[n]);
        if (cb) t += to_f(cb[(m / hw) * p.N + n]);
      }
      v[j] = t;
    }
    if (p.epilogue == RB200_EPI_GEGLU) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int64_t n = n0 + cols[j];
        if (n >= p.N) continue;
        const int64_t no = (n0 + cg) / 2 + cj + j;
        float o = v[j] * gelu_erf(v[j + 2]);
        if (res) o += to_f(res[m * p.ldr + no]);
        y[m * p.ldy + no] = from_f<T>(o);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + cols[j];
        if (n >= p.N) continue;
        float o = apply_epilogue(v[j], p.epilogue);
        if (res) o += to_f(res[m * p.ldr + n]);
        y[m * p.ldy + n] = from_f<T>(o);
      }
    }
  }
}

template <typename T>
int launch(cudaStream_t st, const GemmProblem& p) {
  if (ceil_div(p.N, TN) > 65535 || ceil_div(p.M, TM) > 2147483647LL) RB200_FAIL(-3, "simt gemm: %lld x %lld is too large for one launch", (long long)p.M, (long long)p.N);
  dim3 grid((unsigned)ceil_div(p.M, TM), (unsigned)ceil_div(p.N, TN));
  if (p.conv)
    simt_gemm_kernel<T, true><<<grid, 256, 0, st>>>(p);
  else
    simt_gemm_kernel<T, false><<<grid, 256, 0, st>>>(p);
  RB200_CHECK_LAUNCH("simt_gemm");
  return 0;
}

}  // namespace

int simt_gemm(cudaStream_t st, const GemmProblem& p) {
  if (p.M <= 0 || p.N <= 0) return 0;
  if (p.epilogue == RB200_EPI_GEGLU && (p.N % 32) != 0) RB200_FAIL(-1, "GEGLU epilogue needs N %% 32 == 0 (N=%lld)", (long long)p.N);
  switch (p.dtype) {
    case RB200_BF16: return launch<__nv_bfloat16>(st, p);
    case RB200_FP16: return launch<__half>(st, p);
    case RB200_FP32: return launch<float>(st, p);
  }
  RB200_FAIL(-1, "simt_gemm: bad dtype %d", p.dtype);
}

}  // namespace rb200
