// Generic flash-style attention on CUDA cores: any head dim <= 512 (the VAE bottleneck's single 512-wide head,
// latent_diffusion/auto_encoder.py:175 of the reference), any sequence lengths, any dtype.  It is the correctness anchor for the tcgen05 kernel and the path for shapes that
// kernel does not take.  Optional second key/value set (IP-Adapter): two independent
// softmaxes, o = A(q,k,v) + scale2 * A(q,k2,v2).
#include "common.cuh"

namespace rb200 {
namespace {

constexpr int QT = 32, KT = 32, NT = 128, DMAX = 512;  // DPT = output columns per thread: 64 (D <= 256) or 128

template <typename T, int DPT>
__global__ void __launch_bounds__(NT) simt_sdpa_kernel(const SdpaProblem p) {
  extern __shared__ float sm[];
  const int D = p.D;
  float* Qs = sm;                  // [QT][D]
  float* Ks = Qs + QT * D;         // [KT][D+1]
  float* Vs = Ks + KT * (D + 1);   // [KT][D]
  float* Ss = Vs + KT * D;         // [QT][KT+1]

  const int tid = threadIdx.x;
  const int r = tid >> 2, part = tid & 3;
  const int dpp = (D + 3) / 4;
  const int d0 = part * dpp;
  const int d1 = (d0 + dpp < D) ? d0 + dpp : D;
  const int64_t q0 = int64_t(blockIdx.x) * QT;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;

  const T* q = static_cast<const T*>(p.q) + b * p.q_sb + int64_t(h) * D;
  for (int e = tid; e < QT * D; e += NT) {
    const int rr = e / D, d = e - rr * D;
    const int64_t qi = q0 + rr;
    Qs[e] = (qi < p.Sq) ? to_f(q[qi * p.q_ss + d]) * p.scale : 0.f;
  }

  float out[DPT];
#pragma unroll
  for (int i = 0; i < DPT; ++i) out[i] = 0.f;

  const int nsets = (p.k2 != nullptr && p.Sk2 > 0) ? 2 : 1;
  for (int set = 0; set < nsets; ++set) {
    const T* k = static_cast<const T*>(set ? p.k2 : p.k) + b * (set ? p.k2_sb : p.k_sb) + int64_t(h) * D;
    const T* v = static_cast<const T*>(set ? p.v2 : p.v) + b * (set ? p.v2_sb : p.v_sb) + int64_t(h) * D;
    const int64_t k_ss = set ? p.k2_ss : p.k_ss, v_ss = set ? p.v2_ss : p.v_ss;
    const int64_t Sk = set ? p.Sk2 : p.Sk;
    const bool causal = p.causal && set == 0;

    float acc[DPT];
#pragma unroll
    for (int i = 0; i < DPT; ++i) acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int64_t kt = 0; kt < Sk; kt += KT) {
      __syncthreads();
      for (int e = tid; e < KT * D; e += NT) {
        const int j = e / D, d = e - j * D;
        const int64_t kj = kt + j;
        const bool ok = kj < Sk;
        Ks[j * (D + 1) + d] = ok ? to_f(k[kj * k_ss + d]) : 0.f;
        Vs[j * D + d] = ok ? to_f(v[kj * v_ss + d]) : 0.f;
      }
      __syncthreads();
      // scores for row r, keys part*8 .. part*8+7
      float s[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) s[jj] = 0.f;
      for (int d = 0; d < D; ++d) {
        const float qv = Qs[r * D + d];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) s[jj] = fmaf(qv, Ks[(part * 8 + jj) * (D + 1) + d], s[jj]);
      }
      if (p.bias_h != nullptr && set == 0 && q0 + r < p.Sq) {
        // one block of (bias_H + bias_W) x 128 floats per (b, head, 128-query tile): [k][q % 128]
        const int64_t qq = q0 + r, nqt = (p.Sq + 127) / 128;
        const float* bh = p.bias_h + (((b * p.H + h) * nqt + qq / 128) * (p.bias_H + p.bias_W)) * 128 + qq % 128;
        const float* bw = bh + p.bias_H * 128;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int64_t kj = kt + part * 8 + jj;
          if (kj < Sk) s[jj] = (s[jj] + bh[(kj / p.bias_W) * 128]) + bw[(kj % p.bias_W) * 128];  // vertical term first, as in the reference
        }
      }
      float tmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int64_t kj = kt + part * 8 + jj;
        const bool masked = (kj >= Sk) || (causal && kj > q0 + r);
        if (masked) s[jj] = -INFINITY;
        tmax = fmaxf(tmax, s[jj]);
      }
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
      const float m_new = fmaxf(m_run, tmax);
      const float alpha = (m_new == -INFINITY) ? 1.f : __expf(m_run - m_new);
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float pv = (s[jj] == -INFINITY) ? 0.f : __expf(s[jj] - m_new);
        psum += pv;
        Ss[r * (KT + 1) + part * 8 + jj] = pv;
      }
      psum += __shfl_xor_sync(0xffffffffu, psum, 1);
      psum += __shfl_xor_sync(0xffffffffu, psum, 2);
      l_run = l_run * alpha + psum;
      m_run = m_new;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < DPT; ++i) acc[i] *= alpha;
      for (int j = 0; j < KT; ++j) {
        const float pv = Ss[r * (KT + 1) + j];
#pragma unroll
        for (int i = 0; i < DPT; ++i) {
          const int d = d0 + i;
          if (d < d1) acc[i] = fmaf(pv, Vs[j * D + d], acc[i]);
        }
      }
    }
    const float w = (set ? p.scale2 : 1.f) * (l_run > 0.f ? 1.f / l_run : 0.f);
#pragma unroll
    for (int i = 0; i < DPT; ++i) out[i] = fmaf(w, acc[i], out[i]);
  }

  const int64_t qi = q0 + r;
  if (qi < p.Sq) {
    T* o = static_cast<T*>(p.o) + b * p.o_sb + qi * p.o_ss + int64_t(h) * D;
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      const int d = d0 + i;
      if (d < d1) o[d] = from_f<T>(out[i]);
    }
  }
}

template <typename T, int DPT>
int launch_dpt(cudaStream_t st, const SdpaProblem& p) {
  const size_t smem = sizeof(float) * (size_t(QT) * p.D + size_t(KT) * (p.D + 1) + size_t(KT) * p.D + size_t(QT) * (KT + 1));
  if (smem > 48 * 1024 && cudaFuncSetAttribute(simt_sdpa_kernel<T, DPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
    RB200_FAIL(-2, "simt_sdpa: cannot reserve %zu bytes of shared memory (head dim %d)", smem, p.D);
  if (p.B > 65535 || p.H > 65535) RB200_FAIL(-1, "sdpa: batch/heads too large for one launch");
  dim3 grid((unsigned)ceil_div(p.Sq, QT), (unsigned)p.H, (unsigned)p.B);
  simt_sdpa_kernel<T, DPT><<<grid, NT, smem, st>>>(p);
  RB200_CHECK_LAUNCH("simt_sdpa");
  return 0;
}

template <typename T>
int launch(cudaStream_t st, const SdpaProblem& p) {
  return p.D <= 256 ? launch_dpt<T, 64>(st, p) : launch_dpt<T, 128>(st, p);
}

}  // namespace

int simt_sdpa(cudaStream_t st, const SdpaProblem& p) {
  if (p.D < 1 || p.D > DMAX) RB200_FAIL(-1, "sdpa: head dim %d unsupported (1..%d)", p.D, DMAX);
  if (p.B <= 0 || p.Sq <= 0) return 0;
  switch (p.dtype) {
    case RB200_BF16: return launch<__nv_bfloat16>(st, p);
    case RB200_FP16: return launch<__half>(st, p);
    case RB200_FP32: return launch<float>(st, p);
  }
  RB200_FAIL(-1, "sdpa: bad dtype %d", p.dtype);
}

}  // namespace rb200
