// Shared helpers for the refiners_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/refiners_b200.h"

namespace rb200 {

// ------------------------------------------------------------------ error plumbing (api.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int kernel_mode();  // 0 auto, 1 force SIMT

#define RB200_FAIL(code, ...)      \
  do {                             \
    ::rb200::set_error(__VA_ARGS__); \
    return (code);                 \
  } while (0)

#define RB200_CHECK_LAUNCH(what)                                                   \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) RB200_FAIL(-2, "%s: %s", what, cudaGetErrorString(e__)); \
    ::rb200::count_launch();                                                       \
  } while (0)

// ------------------------------------------------------------------------ dtype conversion
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); T v[16 / sizeof(T)]; };

template <typename T>
__device__ __forceinline__ Vec16<T> ld16(const T* p) {
  Vec16<T> r;
  *reinterpret_cast<uint4*>(r.v) = *reinterpret_cast<const uint4*>(p);
  return r;
}
template <typename T>
__device__ __forceinline__ void st16(T* p, const Vec16<T>& r) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(r.v);
}

// ------------------------------------------------------------------------------ activations
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// erf GeLU for the 16-bit tensor-core epilogues: Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, two orders below
// bf16 / fp16 output rounding), branch free, 2 MUFU + ~12 FMA-class instructions instead of erff's ~30 with a
// divergent branch.  The negative tail uses erfc directly (1 + erf(-z) = erfc(z)), so there is no cancellation.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float erfc_abs = poly * t * e;                        // 1 - erf(|x| / sqrt 2)
  const float cdf2 = x >= 0.f ? 2.0f - erfc_abs : erfc_abs;   // 1 + erf(x / sqrt 2)
  return 0.5f * x * cdf2;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
// SiLU for 16-bit outputs: ex2.approx + rcp.approx (2 MUFU, ~2 ulp in fp32) instead of an IEEE division
__device__ __forceinline__ float silu_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

__device__ __forceinline__ float apply_epilogue(float v, int epi) {
  if (epi == RB200_EPI_GELU) return gelu_erf(v);
  if (epi == RB200_EPI_SILU) return silu_f(v);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t dtype_size(int dtype) { return dtype == RB200_FP32 ? 4 : 2; }
constexpr int kMaxDevices = 64;
int current_device();  // ordinal of the current CUDA device, clamped to [0, kMaxDevices)
int sm_count();        // of the current device, cached per ordinal (api.cu)

// One-time per-device set-up of a kernel (cudaFuncSetAttribute is per device): `if (once.needed()) {...; once.done();}`
struct PerDeviceOnce {
  bool flag[kMaxDevices] = {};
  bool needed() const { return !flag[current_device()]; }
  void done() { flag[current_device()] = true; }
};

// ---------------------------------------------------------- kernel families (one .cu each)
// GEMM / conv problem handed from api.cu to the SIMT or tcgen05 implementation.
struct GemmProblem {
  int dtype;
  // A operand: plain [M, K] rows (conv == 0) or NHWC gather (conv == 1)
  const void* a; int64_t lda;
  const void* b; int64_t ldb;       // [N, K] K-major rows; conv: [taps][N][Cin]
  const void* a2; int64_t lda2;     // optional second K segment (LoRA up-projection)
  const void* b2; int64_t ldb2; int64_t K2;
  const void* bias;                 // [N] or null
  const float* colscale;            // [N] fp32 or null: acc *= colscale[n] before bias
  const void* chan_bias;            // conv only: [B, N] or null
  const void* residual; int64_t ldr;  // [M, N_out] or null
  void* y; int64_t ldy;
  int64_t M, N, K;
  int epilogue;
  // conv geometry (conv == 1): M = B*Ho*Wo, K = R*S*Cin
  int conv; int64_t B, H, W, Cin, Ho, Wo; int R, S, stride, pad;
};

int simt_gemm(cudaStream_t st, const GemmProblem& p);
bool tc_gemm_supported(const GemmProblem& p);
int tc_gemm(cudaStream_t st, const GemmProblem& p);

struct SdpaProblem {
  int dtype;
  const void *q, *k, *v; void* o;
  int64_t B; int H; int64_t Sq, Sk; int D;
  int64_t q_sb, q_ss, k_sb, k_ss, v_sb, v_ss, o_sb, o_ss;
  float scale; int causal;
  const void *k2, *v2; int64_t Sk2, k2_sb, k2_ss, v2_sb, v2_ss; float scale2;
  // optional decomposed relative-position bias (SAM): bias_h points at the combined fp32 table written by
  // rel_bias_kernel, one [(bias_H + bias_W)][128] block per (b, h, 128-query tile); bias_w is unused (nullptr):
  //   logits[q, kh * bias_W + kw] += blk[kh][q % 128] + blk[bias_H + kw][q % 128]
  const float *bias_h, *bias_w; int bias_H, bias_W;
  // the relative-position embeddings themselves ([2 H - 1, D], [2 W - 1, D], operand dtype): a kernel that computes the
  // bias on its own (tc_sdpa_win_fuses_bias) reads these and ignores the table
  const void *rel_h_emb, *rel_w_emb;
};
int simt_sdpa(cudaStream_t st, const SdpaProblem& p);
bool tc_sdpa_supported(const SdpaProblem& p);
int tc_sdpa(cudaStream_t st, const SdpaProblem& p);
// second-generation kernel (tc_attention2.cu): head dim <= 64, one K/V set, no bias, Sq > 128
bool tc_sdpa2_supported(const SdpaProblem& p);
int tc_sdpa2(cudaStream_t st, const SdpaProblem& p);
// single-pass kernel for short key sequences (tc_attention_short.cu): Sk <= 128, head dim <= 64, one K/V set, no bias
bool tc_sdpa_short_supported(const SdpaProblem& p);
int tc_sdpa_short(cudaStream_t st, const SdpaProblem& p);
// SAM's 14 x 14 windows (tc_attention_win.cu): head dim 65..80, decomposed relative-position bias, 196 queries = keys
bool tc_sdpa_win_supported(const SdpaProblem& p);
bool tc_sdpa_win_fuses_bias(const SdpaProblem& p);   // true: the caller need not fill the bias table
int tc_sdpa_win(cudaStream_t st, const SdpaProblem& p);

}  // namespace rb200
