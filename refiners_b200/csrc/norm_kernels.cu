// HBM-bound kernels: GroupNorm(+SiLU) on NHWC, LayerNorm, unary activations, GEGLU, add,
// plus the one-time weight packers.  All reductions use a fixed order (no atomics), so two
// identical forwards are bit-identical (the reference pins this in
// tests/foundationals/latent_diffusion/test_sd15_unet.py:21-37).
#include "common.cuh"

namespace rb200 {
namespace {

// ------------------------------------------------------------------------------ GroupNorm
// pixels per block: small enough that B * ceil(HW / pix) blocks fill the 148 SMs several times over
inline int gn_pix(int64_t B, int64_t HW) {
  int pix = 128;  // (256 / 512-pixel chunks measured the same on [16, 320, 128, 128], 64 is 7 % slower)
  while (pix > 8 && B * ceil_div(HW, pix) < int64_t(sm_count()) * 6) pix >>= 1;
  return pix;
}

template <typename T, int V>
__global__ void gn_partial_kernel(const T* __restrict__ x, float* __restrict__ part, int64_t HW, int C, int G,
                                  int chunks, int GN_PIX) {
  extern __shared__ float sm[];  // [lanes][C][2]
  const int CV = C / V;
  const int v = threadIdx.x, lane = threadIdx.y, lanes = blockDim.y;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int64_t p0 = int64_t(chunk) * GN_PIX;
  const int64_t p1 = (p0 + GN_PIX < HW) ? p0 + GN_PIX : HW;
  float s[V], q[V];
#pragma unroll
  for (int e = 0; e < V; ++e) s[e] = q[e] = 0.f;
  const T* base = x + (int64_t(b) * HW) * C + int64_t(v) * V;
  constexpr int U = 4;  // pixels in flight per thread: four independent 16-byte loads hide the HBM latency
  for (int64_t p = p0 + lane; p < p1; p += int64_t(lanes) * U) {
    T vals[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = p + int64_t(u) * lanes;
      if (pu < p1) {
        if constexpr (V * sizeof(T) == 16) {
          *reinterpret_cast<uint4*>(vals[u]) = *reinterpret_cast<const uint4*>(base + pu * C);
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) vals[u][e] = base[pu * C + e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) vals[u][e] = from_f<T>(0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {  // fixed order u = 0..3: the sums are run-to-run deterministic
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float f = to_f(vals[u][e]);
        s[e] += f;
        q[e] = fmaf(f, f, q[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) {
    sm[(lane * C + v * V + e) * 2 + 0] = s[e];
    sm[(lane * C + v * V + e) * 2 + 1] = q[e];
  }
  __syncthreads();
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
  if (t < G) {
    const int cpg = C / G;
    float ss = 0.f, qq = 0.f;
    for (int l = 0; l < lanes; ++l)
      for (int c = t * cpg; c < (t + 1) * cpg; ++c) {
        ss += sm[(l * C + c) * 2 + 0];
        qq += sm[(l * C + c) * 2 + 1];
      }
    float* out = part + ((int64_t(b) * chunks + chunk) * G + t) * 2;
    out[0] = ss;
    out[1] = qq;
  }
  (void)CV;
}

// one warp per (sample, group): fixed-order fp64 combination of the per-chunk partial sums
__global__ void gn_finalize_kernel(const float* __restrict__ part, float* __restrict__ stats, int64_t HW, int C, int G,
                                   int chunks, float eps, int total) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= total) return;
  const int b = gw / G, g = gw - b * G;
  const float* src = part + (int64_t(b) * chunks * G + g) * 2;
  double ss = 0.0, qq = 0.0;
  for (int c = lane; c < chunks; c += 32) {
    ss += double(src[int64_t(c) * G * 2 + 0]);
    qq += double(src[int64_t(c) * G * 2 + 1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    qq += __shfl_xor_sync(0xffffffffu, qq, o);
  }
  if (lane == 0) {
    const double n = double(HW) * double(C / G);
    const double mean = ss / n;
    double var = qq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(int64_t(b) * G + g) * 2 + 0] = float(mean);
    stats[(int64_t(b) * G + g) * 2 + 1] = float(1.0 / sqrt(var + double(eps)));
  }
}

template <typename T, int V>
__global__ void gn_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ stats,
                                const T* __restrict__ gamma, const T* __restrict__ beta, int64_t HW, int C, int G,
                                int GN_PIX, int silu) {
  extern __shared__ float sm[];  // mean[G], rstd[G]
  const int v = threadIdx.x, lane = threadIdx.y, lanes = blockDim.y;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthreads = blockDim.x * blockDim.y;
  const int cpg = C / G;
  for (int g = t; g < G; g += nthreads) {
    sm[g] = stats[(int64_t(b) * G + g) * 2 + 0];
    sm[G + g] = stats[(int64_t(b) * G + g) * 2 + 1];
  }
  __syncthreads();
  float sc[V], be[V], mu[V];  // y = (x - mean) * (rstd * gamma) + beta
#pragma unroll
  for (int e = 0; e < V; ++e) {
    const int c = v * V + e;
    sc[e] = sm[G + c / cpg] * to_f(gamma[c]);
    be[e] = to_f(beta[c]);
    mu[e] = sm[c / cpg];
  }
  const int64_t p0 = int64_t(chunk) * GN_PIX;
  const int64_t p1 = (p0 + GN_PIX < HW) ? p0 + GN_PIX : HW;
  const int64_t base = (int64_t(b) * HW) * C + int64_t(v) * V;
  constexpr int U = 4;  // pixels in flight per thread
  for (int64_t p = p0 + lane; p < p1; p += int64_t(lanes) * U) {
    T vals[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = p + int64_t(u) * lanes;
      if (pu < p1) {
        if constexpr (V * sizeof(T) == 16) {
          *reinterpret_cast<uint4*>(vals[u]) = *reinterpret_cast<const uint4*>(x + base + pu * C);
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) vals[u][e] = x[base + pu * C + e];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = p + int64_t(u) * lanes;
      if (pu >= p1) break;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float f = fmaf(to_f(vals[u][e]) - mu[e], sc[e], be[e]);
        if (silu) f = sizeof(T) == 2 ? silu_fast(f) : silu_f(f);
        vals[u][e] = from_f<T>(f);
      }
      if constexpr (V * sizeof(T) == 16) {
        *reinterpret_cast<uint4*>(y + base + pu * C) = *reinterpret_cast<const uint4*>(vals[u]);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) y[base + pu * C + e] = vals[u][e];
      }
    }
  }
}

template <typename T, int V>
int gn_launch(cudaStream_t st, const T* x, T* y, int64_t B, int64_t HW, int C, int G, float eps, const T* gamma,
              const T* beta, int silu, float* part, float* fixed_stats, int frozen) {
  const int CV = C / V;
  if (CV > 1024) RB200_FAIL(-1, "group_norm: C=%d too wide for this layout", C);
  const int pix = gn_pix(B, HW);
  int lanes = 256 / CV;
  if (lanes < 1) lanes = 1;
  if (lanes > pix) lanes = pix;
  while (CV * lanes < G) ++lanes;  // need at least G threads for the group reduce
  const int chunks = int(ceil_div(HW, pix));
  dim3 block(CV, lanes), grid(chunks, (unsigned)B);
  const size_t sm1 = size_t(lanes) * C * 2 * sizeof(float);
  if (sm1 > 200 * 1024) RB200_FAIL(-1, "group_norm: shared memory need %zu too large", sm1);
  if (sm1 > 48 * 1024) cudaFuncSetAttribute(gn_partial_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm1));
  // (mean, rstd) per (sample, group): the workspace's own slot, or the caller's buffer - which is either filled here
  // (first pass of a FixedGroupNorm) or, when `frozen`, read as it is: no statistics pass at all
  float* stats = fixed_stats ? fixed_stats : part + size_t(B) * chunks * G * 2;
  if (!frozen) {
    gn_partial_kernel<T, V><<<grid, block, sm1, st>>>(x, part, HW, C, G, chunks, pix);
    RB200_CHECK_LAUNCH("gn_partial");
    const int total = int(B) * G;
    gn_finalize_kernel<<<unsigned(ceil_div(int64_t(total) * 32, 256)), 256, 0, st>>>(part, stats, HW, C, G, chunks, eps, total);
    RB200_CHECK_LAUNCH("gn_finalize");
  }
  gn_apply_kernel<T, V><<<grid, block, 2 * G * sizeof(float), st>>>(x, y, stats, gamma, beta, HW, C, G, pix, silu);
  RB200_CHECK_LAUNCH("gn_apply");
  return 0;
}

template <typename T>
int gn_dispatch(cudaStream_t st, const void* x, void* y, int64_t B, int64_t HW, int64_t C, int G, float eps,
                const void* gamma, const void* beta, int silu, float* part, float* fixed_stats, int frozen) {
  constexpr int VMAX = 16 / sizeof(T);
  const bool aligned = (C % VMAX == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0);
  if (aligned)
    return gn_launch<T, VMAX>(st, (const T*)x, (T*)y, B, HW, int(C), G, eps, (const T*)gamma, (const T*)beta, silu, part, fixed_stats, frozen);
  return gn_launch<T, 1>(st, (const T*)x, (T*)y, B, HW, int(C), G, eps, (const T*)gamma, (const T*)beta, silu, part, fixed_stats, frozen);
}

// ------------------------------------------------------------------------------ LayerNorm
template <typename T>
__global__ void __launch_bounds__(256) layer_norm_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int C,
                                                         float eps, const T* __restrict__ gamma,
                                                         const T* __restrict__ beta, int vec_ok) {
  constexpr int V = 16 / sizeof(T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const T* xr = x + row * C;
  T* yr = y + row * C;
  float s = 0.f;
  if (vec_ok) {
    for (int c = lane * V; c < C; c += 32 * V) {
      Vec16<T> r = ld16(xr + c);
#pragma unroll
      for (int e = 0; e < V; ++e) s += to_f(r.v[e]);
    }
  } else {
    for (int c = lane; c < C; c += 32) s += to_f(xr[c]);
  }
  const float mean = warp_sum(s) / float(C);
  float q = 0.f;
  if (vec_ok) {
    for (int c = lane * V; c < C; c += 32 * V) {
      Vec16<T> r = ld16(xr + c);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float d = to_f(r.v[e]) - mean;
        q = fmaf(d, d, q);
      }
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      const float d = to_f(xr[c]) - mean;
      q = fmaf(d, d, q);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / float(C) + eps);
  if (vec_ok) {
    for (int c = lane * V; c < C; c += 32 * V) {
      Vec16<T> r = ld16(xr + c), g = ld16(gamma + c), bt = ld16(beta + c);
#pragma unroll
      for (int e = 0; e < V; ++e) r.v[e] = from_f<T>((to_f(r.v[e]) - mean) * rstd * to_f(g.v[e]) + to_f(bt.v[e]));
      st16(yr + c, r);
    }
  } else {
    for (int c = lane; c < C; c += 32) yr[c] = from_f<T>((to_f(xr[c]) - mean) * rstd * to_f(gamma[c]) + to_f(beta[c]));
  }
}

// Row kept in registers: one global read, two reductions, one write (VPL 16-byte vectors per lane).
template <typename T, int VPL>
__global__ void __launch_bounds__(256) layer_norm_reg_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int C,
                                                             float eps, const T* __restrict__ gamma,
                                                             const T* __restrict__ beta) {
  constexpr int V = 16 / sizeof(T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const T* xr = x + row * C;
  T* yr = y + row * C;
  const int nvec = C / V;
  Vec16<T> r[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      r[i] = ld16(xr + vi * V);
#pragma unroll
      for (int e = 0; e < V; ++e) s += to_f(r[i].v[e]);
    }
  }
  const float mean = warp_sum(s) / float(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float d = to_f(r[i].v[e]) - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / float(C) + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const Vec16<T> g = ld16(gamma + vi * V), bt = ld16(beta + vi * V);
      Vec16<T> o;
#pragma unroll
      for (int e = 0; e < V; ++e) o.v[e] = from_f<T>((to_f(r[i].v[e]) - mean) * rstd * to_f(g.v[e]) + to_f(bt.v[e]));
      st16(yr + vi * V, o);
    }
  }
}

template <typename T>
void ln_launch(cudaStream_t st, unsigned grid, int warps, const T* x, T* y, int64_t rows, int C, float eps, const T* gamma,
               const T* beta, int vec_ok) {
  constexpr int V = 16 / sizeof(T);
  const int vpl = vec_ok ? int(ceil_div(C / V, 32)) : 99;
  const int threads = warps * 32;
  if (vpl <= 1) layer_norm_reg_kernel<T, 1><<<grid, threads, 0, st>>>(x, y, rows, C, eps, gamma, beta);
  else if (vpl == 2) layer_norm_reg_kernel<T, 2><<<grid, threads, 0, st>>>(x, y, rows, C, eps, gamma, beta);
  else if (vpl == 3) layer_norm_reg_kernel<T, 3><<<grid, threads, 0, st>>>(x, y, rows, C, eps, gamma, beta);
  else if (vpl <= 5) layer_norm_reg_kernel<T, 5><<<grid, threads, 0, st>>>(x, y, rows, C, eps, gamma, beta);
  else if (vpl <= 8) layer_norm_reg_kernel<T, 8><<<grid, threads, 0, st>>>(x, y, rows, C, eps, gamma, beta);
  else layer_norm_kernel<T><<<grid, threads, 0, st>>>(x, y, rows, C, eps, gamma, beta, vec_ok);
}

// ---------------------------------------------------------------------------- elementwise
__device__ __forceinline__ float unary_f(float x, int op) {
  switch (op) {
    case RB200_UNARY_SILU: return silu_f(x);
    case RB200_UNARY_GELU: return gelu_erf(x);
    case RB200_UNARY_GELU_TANH: return gelu_tanh(x);
    case RB200_UNARY_GELU_SIGMOID: return x * sigmoid_f(1.702f * x);
    case RB200_UNARY_RELU: return fmaxf(x, 0.f);
    default: return sigmoid_f(x);
  }
}

template <typename T>
__global__ void unary_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, int op, int vec_ok) {
  constexpr int V = 16 / sizeof(T);
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t nv = n / V;
    for (; i < nv; i += stride) {
      Vec16<T> r = ld16(x + i * V);
#pragma unroll
      for (int e = 0; e < V; ++e) r.v[e] = from_f<T>(unary_f(to_f(r.v[e]), op));
      st16(y + i * V, r);
    }
    for (int64_t j = nv * V + (int64_t(blockIdx.x) * blockDim.x + threadIdx.x); j < n; j += stride)
      y[j] = from_f<T>(unary_f(to_f(x[j]), op));
  } else {
    for (; i < n; i += stride) y[i] = from_f<T>(unary_f(to_f(x[i]), op));
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t n, float alpha,
                           int vec_ok) {
  constexpr int V = 16 / sizeof(T);
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t nv = n / V;
    for (; i < nv; i += stride) {
      Vec16<T> r = ld16(a + i * V), s = ld16(b + i * V);
#pragma unroll
      for (int e = 0; e < V; ++e) r.v[e] = from_f<T>(fmaf(alpha, to_f(s.v[e]), to_f(r.v[e])));
      st16(y + i * V, r);
    }
    for (int64_t j = nv * V + (int64_t(blockIdx.x) * blockDim.x + threadIdx.x); j < n; j += stride)
      y[j] = from_f<T>(fmaf(alpha, to_f(b[j]), to_f(a[j])));
  } else {
    for (; i < n; i += stride) y[i] = from_f<T>(fmaf(alpha, to_f(b[i]), to_f(a[i])));
  }
}

// Channel concatenation of up to four NHWC maps (the UNet's skip connections, latent_diffusion/unet.py:66-79) and
// nearest-neighbour resize (layers/sampling.py:13-38), both as one pass of 16-byte vector copies.
struct CatSources {
  const void* src[4];
  int channels[4];
  int n;
};

template <typename T>
__global__ void concat_channels_kernel(const CatSources cs, T* __restrict__ y, int64_t pixels, int Ct) {
  constexpr int V = 16 / sizeof(T);
  const int cv = Ct / V;
  const int64_t total = pixels * cv;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int c = int(idx % cv) * V;
    const int64_t pix = idx / cv;
    int which = 0;
    while (which < cs.n - 1 && c >= cs.channels[which]) {
      c -= cs.channels[which];
      ++which;
    }
    const T* src = static_cast<const T*>(cs.src[which]) + pix * cs.channels[which] + c;
    st16(y + idx * V, ld16(src));
  }
}

template <typename T>
__global__ void resize_nearest_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t B, int H, int W, int C, int Ho, int Wo,
                                      float scale_h, float scale_w) {
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V;
  const int64_t total = B * Ho * Wo * cv;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int v = int(idx % cv);
    int64_t t = idx / cv;
    const int wo = int(t % Wo);
    t /= Wo;
    const int ho = int(t % Ho);
    const int64_t b = t / Ho;
    // ATen's legacy "nearest": src = min(floor(dst * in / out), in - 1), computed in float
    int hi = int(floorf(float(ho) * scale_h)), wi = int(floorf(float(wo) * scale_w));
    hi = hi < H - 1 ? hi : H - 1;
    wi = wi < W - 1 ? wi : W - 1;
    st16(y + idx * V, ld16(x + ((b * H + hi) * W + wi) * C + v * V));
  }
}

// k x k average pooling with stride k of a dense NHWC map (T2I-Adapter's Downsample2d); fp32 sum in raster order of the
// window, one rounding.  One thread per output pixel-vector of 16 bytes.
template <typename T>
__global__ void avg_pool_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t B, int H, int W, int C, int k) {
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V, Ho = H / k, Wo = W / k;
  const int64_t total = B * Ho * Wo * cv;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const float inv = 1.0f / float(k * k);
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int v = int(idx % cv);
    int64_t t = idx / cv;
    const int wo = int(t % Wo);
    t /= Wo;
    const int ho = int(t % Ho);
    const int64_t b = t / Ho;
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        const Vec16<T> in = ld16(x + ((b * H + ho * k + dy) * W + wo * k + dx) * C + v * V);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += to_f(in.v[e]);
      }
    Vec16<T> out;
#pragma unroll
    for (int e = 0; e < V; ++e) out.v[e] = from_f<T>(acc[e] * inv);
    st16(y + idx * V, out);
  }
}

// Channel padding for the tensor-core conv (needs Cin % 8 == 0): y[b, h, w, 0..Cp) = x[b, 0..C, h, w] then zeros.
// x is addressed through element strides (NCHW or channels-last), y is dense NHWC; one thread per output pixel-vector.
template <typename T>
__global__ void pad_channels_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t pixels, int H, int W, int C, int Cp, int64_t sb,
                                    int64_t sc, int64_t sh, int64_t sw) {
  constexpr int V = 16 / sizeof(T);
  const int cv = Cp / V;
  const int64_t total = pixels * cv;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int v = int(idx % cv);
    const int64_t pix = idx / cv;
    const int w = int(pix % W);
    const int h = int((pix / W) % H);
    const int64_t b = pix / (int64_t(W) * H);
    const T* src = x + b * sb + int64_t(h) * sh + int64_t(w) * sw;
    Vec16<T> val;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const int c = v * V + e;
      val.v[e] = c < C ? src[int64_t(c) * sc] : from_f<T>(0.f);
    }
    st16(y + pix * Cp + v * V, val);
  }
}

// Non-overlapping P x P patches of an image as GEMM rows: y[(b, ho, wo), (r, s, c)] = x[b, c, ho P + r, wo P + s]
// (x addressed through element strides, so NCHW and channels-last inputs both work without a layout copy).
template <typename T>
__global__ void patchify_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t total, int Ho, int Wo, int C, int P, int64_t sb,
                                int64_t sc, int64_t sh, int64_t sw) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = int(i % C);
    int64_t t = i / C;
    const int s = int(t % P);
    t /= P;
    const int r = int(t % P);
    t /= P;
    const int wo = int(t % Wo);
    t /= Wo;
    const int ho = int(t % Ho);
    const int64_t b = t / Ho;
    y[i] = x[b * sb + c * sc + (int64_t(ho) * P + r) * sh + (int64_t(wo) * P + s) * sw];
  }
}

// Window partition (zero padded up to multiples of ws) and its inverse, on channels-last rows of C elements:
//   partition: y[(b, wh, ww), i, j, :] = x[b, wh ws + i, ww ws + j, :]   (0 outside the H x W map)
//   merge:     y[b, h, w, :]           = x[(b, h / ws, w / ws), h % ws, w % ws, :]
template <typename T, bool MERGE>
__global__ void window_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t B, int H, int W, int C, int ws, int nH, int nW) {
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V;
  const int64_t rows = MERGE ? B * H * W : B * nH * nW * ws * ws;
  const int64_t total = rows * cv;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int v = int(idx % cv);
    int64_t row = idx / cv;
    if (MERGE) {
      const int w = int(row % W);
      const int h = int((row / W) % H);
      const int64_t b = row / (int64_t(W) * H);
      const int64_t src = (((b * nH + h / ws) * nW + w / ws) * ws + h % ws) * ws + w % ws;
      st16(y + row * C + v * V, ld16(x + src * C + v * V));
    } else {
      const int j = int(row % ws);
      int64_t t = row / ws;
      const int i = int(t % ws);
      t /= ws;
      const int ww = int(t % nW);
      t /= nW;
      const int wh = int(t % nH);
      const int64_t b = t / nH;
      const int h = wh * ws + i, w = ww * ws + j;
      Vec16<T> val;
      if (h < H && w < W) {
        val = ld16(x + ((b * H + h) * W + w) * C + v * V);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) val.v[e] = from_f<T>(0.f);
      }
      st16(y + row * C + v * V, val);
    }
  }
}

template <typename T>
__global__ void geglu_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int64_t F, int vec_ok) {
  constexpr int V = 16 / sizeof(T);
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t fv = F / V, total = rows * fv;
    for (; i < total; i += stride) {
      const int64_t r = i / fv, c = (i - r * fv) * V;
      Vec16<T> a = ld16(x + r * 2 * F + c), g = ld16(x + r * 2 * F + F + c);
#pragma unroll
      for (int e = 0; e < V; ++e) a.v[e] = from_f<T>(to_f(a.v[e]) * gelu_erf(to_f(g.v[e])));
      st16(y + r * F + c, a);
    }
  } else {
    const int64_t total = rows * F;
    for (; i < total; i += stride) {
      const int64_t r = i / F, c = i - r * F;
      y[i] = from_f<T>(to_f(x[r * 2 * F + c]) * gelu_erf(to_f(x[r * 2 * F + F + c])));
    }
  }
}

// -------------------------------------------------------------------------------- packers
template <typename T>
__global__ void conv_pack_kernel(const T* __restrict__ w, T* __restrict__ out, int64_t Cout, int64_t Cin, int RS) {
  // out[t][n][c] = w[n][c][t]
  const int64_t total = Cout * Cin * RS;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t c = i % Cin, n = (i / Cin) % Cout, t = i / (Cin * Cout);
    out[i] = w[(n * Cin + c) * RS + t];
  }
}

template <typename T>
__global__ void geglu_pack_kernel(const T* __restrict__ w, const T* __restrict__ bias, T* __restrict__ wp,
                                  T* __restrict__ bp, int64_t F, int64_t K) {
  // packed row p: block = p / 32, j = p % 32; j < 16 -> value row block*16 + j, else gate row F + block*16 + j - 16
  const int64_t total = 2 * F * K;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t p = i / K, k = i - p * K;
    const int64_t blk = p / 32, j = p % 32;
    const int64_t src = (j < 16) ? blk * 16 + j : F + blk * 16 + (j - 16);
    wp[i] = w[src * K + k];
    if (bias != nullptr && k == 0) bp[p] = bias[src];
  }
}

struct LoraList { rb200_lora l[8]; };

template <typename T>
__global__ void lora_pack_kernel(const LoraList list, int n_lora, int64_t N, int64_t K,
                                 T* __restrict__ down_cat, T* __restrict__ up_cat, float* __restrict__ colscale,
                                 int r_pad) {
  const rb200_lora* loras = list.l;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int64_t t0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  // down_cat [r_pad, K]
  for (int64_t i = t0; i < int64_t(r_pad) * K; i += stride) {
    const int64_t r = i / K, k = i - r * K;
    int64_t off = 0;
    T val = from_f<T>(0.f);
    for (int l = 0; l < n_lora; ++l) {
      if (r < off + loras[l].rank) {
        val = static_cast<const T*>(loras[l].down)[(r - off) * K + k];
        break;
      }
      off += loras[l].rank;
    }
    down_cat[i] = val;
  }
  // up_cat [N, r_pad]
  for (int64_t i = t0; i < N * r_pad; i += stride) {
    const int64_t n = i / r_pad, r = i - n * r_pad;
    int64_t off = 0;
    T val = from_f<T>(0.f);
    for (int l = 0; l < n_lora; ++l) {
      if (r < off + loras[l].rank) {
        val = static_cast<const T*>(loras[l].up)[n * loras[l].rank + (r - off)];
        break;
      }
      off += loras[l].rank;
    }
    up_cat[i] = val;
  }
  for (int64_t r = t0; r < r_pad; r += stride) {
    int64_t off = 0;
    float sc = 0.f;
    for (int l = 0; l < n_lora; ++l) {
      if (r < off + loras[l].rank) {
        sc = loras[l].scale;
        break;
      }
      off += loras[l].rank;
    }
    colscale[r] = sc;
  }
}

inline int ew_grid(int64_t work_items) {
  int64_t blocks = ceil_div(work_items, 256);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}

inline bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

}  // namespace

#define DISPATCH_T(dtype, ...)                                             \
  switch (dtype) {                                                         \
    case RB200_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }      \
    case RB200_FP16: { using T = __half; __VA_ARGS__; break; }             \
    case RB200_FP32: { using T = float; __VA_ARGS__; break; }              \
    default: RB200_FAIL(-1, "bad dtype %d", dtype);                        \
  }

size_t group_norm_ws(int64_t B, int64_t HW, int G);

int group_norm_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t B, int64_t HW, int64_t C, int G,
                    float eps, const void* gamma, const void* beta, int silu, void* ws, size_t ws_bytes, float* fixed_stats,
                    int frozen) {
  if (G <= 0 || C % G != 0) RB200_FAIL(-1, "group_norm: C=%lld not divisible by G=%d", (long long)C, G);
  const size_t need = group_norm_ws(B, HW, G);
  if (ws_bytes < need || ws == nullptr) RB200_FAIL(-1, "group_norm: workspace %zu < %zu", ws_bytes, need);
  if (B > 65535) RB200_FAIL(-1, "group_norm: batch %lld too large", (long long)B);
  DISPATCH_T(dtype, return gn_dispatch<T>(st, x, y, B, HW, C, G, eps, gamma, beta, silu, static_cast<float*>(ws), fixed_stats, frozen));
  return 0;
}

size_t group_norm_ws(int64_t B, int64_t HW, int G) {
  // per-chunk partial sums + final (mean, rstd) per (sample, group)
  return (size_t(B) * size_t(ceil_div(HW, gn_pix(B, HW))) + size_t(B)) * size_t(G) * 2 * sizeof(float);
}

int layer_norm_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t rows, int64_t C, float eps,
                    const void* gamma, const void* beta) {
  const int warps = 8;
  const unsigned grid = (unsigned)ceil_div(rows, warps);
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    const int vec_ok = (C % V == 0) && aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta);
    ln_launch<T>(st, grid, warps, (const T*)x, (T*)y, rows, int(C), eps, (const T*)gamma, (const T*)beta, vec_ok);
  });
  RB200_CHECK_LAUNCH("layer_norm");
  return 0;
}

int unary_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t n, int op) {
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    const int vec_ok = aligned16(x) && aligned16(y);
    unary_kernel<T><<<ew_grid(n / V + 1), 256, 0, st>>>((const T*)x, (T*)y, n, op, vec_ok);
  });
  RB200_CHECK_LAUNCH("unary");
  return 0;
}

int add_impl(cudaStream_t st, int dtype, const void* a, const void* b, void* y, int64_t n, float alpha) {
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    const int vec_ok = aligned16(a) && aligned16(b) && aligned16(y);
    add_kernel<T><<<ew_grid(n / V + 1), 256, 0, st>>>((const T*)a, (const T*)b, (T*)y, n, alpha, vec_ok);
  });
  RB200_CHECK_LAUNCH("add");
  return 0;
}

// ------------------------------------------------------------------------------ StyleAligned (shared attention)
// foundationals/latent_diffusion/style_aligned.py:13-207 of the reference, for one of q / k / v [B, S, C] of a guidance batch
// (two halves of B / 2 images; ref(b) = first image of b's half):
//   y[b, s < S]      = adain ? (x[b, s] - mean[b]) / (std[b] + eps) * std[ref(b)] + mean[ref(b)] : x[b, s]
//   y[b, S + s]      = x[ref(b), s] * (b == ref(b) ? 1 : scale)                                   (only when concatenating)
// with mean / std per (image, channel) over the S tokens (std unbiased, as torch.std).  Statistics in fp32, two passes over
// rows that stay in L2; the apply pass rounds once.
template <typename T>
__global__ void __launch_bounds__(256) token_stats_kernel(const T* __restrict__ x, float* __restrict__ stats, int64_t S, int C,
                                                          int64_t x_sb, int64_t x_ss) {
  constexpr int V = 16 / sizeof(T);
  constexpr int CH_THREADS = 8;             // 8 threads x V channels per row chunk
  constexpr int ROWS = 256 / CH_THREADS;    // 32 row lanes
  __shared__ float red[ROWS][CH_THREADS * V];
  __shared__ float mean_s[CH_THREADS * V];
  const int ct = threadIdx.x % CH_THREADS, rl = threadIdx.x / CH_THREADS;
  const int c0 = (blockIdx.x * CH_THREADS + ct) * V;
  const int64_t b = blockIdx.y;
  const bool live = c0 < C;
  const T* base = x + b * x_sb + c0;
  float acc[V];
  auto reduce_rows = [&](float (&a)[V]) {  // sum over the 32 row lanes, result in red[0]
#pragma unroll
    for (int e = 0; e < V; ++e) red[rl][ct * V + e] = a[e];
    __syncthreads();
    for (int h = ROWS / 2; h > 0; h >>= 1) {
      if (rl < h) {
#pragma unroll
        for (int e = 0; e < V; ++e) red[rl][ct * V + e] += red[rl + h][ct * V + e];
      }
      __syncthreads();
    }
  };
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.f;
  if (live)
    for (int64_t r = rl; r < S; r += ROWS) {
      const Vec16<T> v = ld16(base + r * x_ss);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] += to_f(v.v[e]);
    }
  reduce_rows(acc);
  if (rl == 0) {
#pragma unroll
    for (int e = 0; e < V; ++e) mean_s[ct * V + e] = red[0][ct * V + e] / float(S);
  }
  __syncthreads();
  float mu[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    mu[e] = mean_s[ct * V + e];
    acc[e] = 0.f;
  }
  if (live)
    for (int64_t r = rl; r < S; r += ROWS) {
      const Vec16<T> v = ld16(base + r * x_ss);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float d = to_f(v.v[e]) - mu[e];
        acc[e] = fmaf(d, d, acc[e]);
      }
    }
  reduce_rows(acc);
  if (rl == 0 && live) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      stats[(b * C + c0 + e) * 2 + 0] = mu[e];
      stats[(b * C + c0 + e) * 2 + 1] = sqrtf(red[0][ct * V + e] / float(S > 1 ? S - 1 : 1));
    }
  }
}

template <typename T>
__global__ void style_aligned_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ stats, int64_t B, int64_t S,
                                           int C, int64_t x_sb, int64_t x_ss, int adain, int concatenate, float scale, float eps) {
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V;
  const int64_t S_out = concatenate ? 2 * S : S, half = B / 2;
  const int64_t total = B * S_out * cv;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int v = int(idx % cv);
    int64_t t = idx / cv;
    const int64_t so = t % S_out, b = t / S_out;
    const int64_t ref = (b / half) * half;
    Vec16<T> out;
    if (so < S) {
      const Vec16<T> in = ld16(x + b * x_sb + so * x_ss + v * V);
      if (adain) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const int c = v * V + e;
          const float m = stats[(b * C + c) * 2], sd = stats[(b * C + c) * 2 + 1];
          const float rm = stats[(ref * C + c) * 2], rs = stats[(ref * C + c) * 2 + 1];
          out.v[e] = from_f<T>((to_f(in.v[e]) - m) / (sd + eps) * rs + rm);
        }
      } else {
        out = in;
      }
    } else {
      const Vec16<T> in = ld16(x + ref * x_sb + (so - S) * x_ss + v * V);
      const float k = b == ref ? 1.0f : scale;
#pragma unroll
      for (int e = 0; e < V; ++e) out.v[e] = from_f<T>(to_f(in.v[e]) * k);
    }
    st16(y + idx * V, out);
  }
}

// ---------------------------------------------------------------- denoising-step glue (CFG + Euler)
// The reference evaluates these with one ATen kernel per arithmetic operator, each rounding its result to the tensor
// dtype (model.py:137-159, solvers/euler.py:63-100).  The fused kernels reproduce that rounding sequence (`rn<T>`), so
// a step through them is bit-identical to the operator-by-operator evaluation.
template <typename T> __device__ __forceinline__ float rn(float v) { return to_f(from_f<T>(v)); }

// y[i] (and y[n + i] when `twice`: the unconditional / conditional halves of classifier-free guidance are the same latents)
//   = x[i] / ((sigma^2 + 1) ^ 0.5),  sigma = sigmas[0]
template <typename T>
__global__ void cfg_scale_input_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, const T* __restrict__ sigmas, int twice) {
  const float sigma = to_f(sigmas[0]);
  // __f*_rn: no FMA contraction - the operator sequence rounds after every multiply and every add, also in fp32
  const float denom = rn<T>(__fsqrt_rn(rn<T>(__fadd_rn(rn<T>(__fmul_rn(sigma, sigma)), 1.0f))));
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const T v = from_f<T>(__fdiv_rn(to_f(x[i]), denom));
    y[i] = v;
    if (twice) y[n + i] = v;
  }
}

// eps: [2n] = (unconditional | conditional) when `guided`, else [n].
//   noise = u + scale * (c - u);   y = x + noise * (sigmas[1] - sigmas[0])
template <typename T>
__global__ void cfg_euler_kernel(const T* __restrict__ x, const T* __restrict__ eps, T* __restrict__ y, int64_t n,
                                 const T* __restrict__ sigmas, float scale, int guided) {
  const float dsigma = rn<T>(__fsub_rn(to_f(sigmas[1]), to_f(sigmas[0])));
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    float noise = to_f(eps[i]);
    if (guided) {
      const float u = noise, c = to_f(eps[n + i]);
      noise = rn<T>(__fadd_rn(u, rn<T>(__fmul_rn(scale, rn<T>(__fsub_rn(c, u))))));
    }
    y[i] = from_f<T>(__fadd_rn(to_f(x[i]), rn<T>(__fmul_rn(noise, dsigma))));
  }
}

int cfg_scale_input_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t n, const void* sigmas, int twice) {
  DISPATCH_T(dtype, { cfg_scale_input_kernel<T><<<ew_grid(n), 256, 0, st>>>((const T*)x, (T*)y, n, (const T*)sigmas, twice); });
  RB200_CHECK_LAUNCH("cfg_scale_input");
  return 0;
}

int cfg_euler_impl(cudaStream_t st, int dtype, const void* x, const void* eps, void* y, int64_t n, const void* sigmas, float scale, int guided) {
  DISPATCH_T(dtype, { cfg_euler_kernel<T><<<ew_grid(n), 256, 0, st>>>((const T*)x, (const T*)eps, (T*)y, n, (const T*)sigmas, scale, guided); });
  RB200_CHECK_LAUNCH("cfg_euler");
  return 0;
}

int geglu_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t rows, int64_t F) {
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    const int vec_ok = (F % V == 0) && aligned16(x) && aligned16(y);
    geglu_kernel<T><<<ew_grid(rows * F / V + 1), 256, 0, st>>>((const T*)x, (T*)y, rows, F, vec_ok);
  });
  RB200_CHECK_LAUNCH("geglu");
  return 0;
}

int concat_channels_impl(cudaStream_t st, int dtype, int n, const void* const* srcs, const int* channels, void* y, int64_t pixels) {
  if (n < 1 || n > 4) RB200_FAIL(-1, "concat_channels: 1..4 sources, got %d", n);
  CatSources cs{};
  cs.n = n;
  int ct = 0;
  for (int i = 0; i < n; ++i) {
    if (!srcs[i] || !aligned16(srcs[i])) RB200_FAIL(-1, "concat_channels: source %d is null or not 16-byte aligned", i);
    cs.src[i] = srcs[i];
    cs.channels[i] = channels[i];
    ct += channels[i];
  }
  if (!aligned16(y)) RB200_FAIL(-1, "concat_channels: output must be 16-byte aligned");
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    for (int i = 0; i < n; ++i)
      if (channels[i] % V != 0) RB200_FAIL(-1, "concat_channels: channel count %d is not a multiple of %d", channels[i], V);
    concat_channels_kernel<T><<<ew_grid(pixels * (ct / V)), 256, 0, st>>>(cs, (T*)y, pixels, ct);
  });
  RB200_CHECK_LAUNCH("concat_channels");
  return 0;
}

int resize_nearest_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t B, int H, int W, int C, int Ho, int Wo) {
  if (!aligned16(x) || !aligned16(y)) RB200_FAIL(-1, "resize_nearest: buffers must be 16-byte aligned");
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    if (C % V != 0) RB200_FAIL(-1, "resize_nearest: C=%d must be a multiple of %d", C, V);
    resize_nearest_kernel<T><<<ew_grid(B * Ho * Wo * (C / V)), 256, 0, st>>>((const T*)x, (T*)y, B, H, W, C, Ho, Wo, float(H) / float(Ho),
                                                                             float(W) / float(Wo));
  });
  RB200_CHECK_LAUNCH("resize_nearest");
  return 0;
}

int style_aligned_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t B, int64_t S, int64_t C, int64_t x_sb, int64_t x_ss, int adain,
                       int concatenate, float scale, float eps, float* stats) {
  if (!aligned16(x) || !aligned16(y)) RB200_FAIL(-1, "style_aligned: buffers must be 16-byte aligned");
  if (B > 65535) RB200_FAIL(-1, "style_aligned: batch %lld too large", (long long)B);
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    if (C % V != 0 || (x_sb * sizeof(T)) % 16 != 0 || (x_ss * sizeof(T)) % 16 != 0)
      RB200_FAIL(-1, "style_aligned: channels and strides must be whole 16-byte vectors");
    if (adain) {
      dim3 grid(unsigned(ceil_div(C, 8 * V)), unsigned(B));
      token_stats_kernel<T><<<grid, 256, 0, st>>>((const T*)x, stats, S, int(C), x_sb, x_ss);
    }
    const int64_t S_out = concatenate ? 2 * S : S;
    style_aligned_apply_kernel<T><<<ew_grid(B * S_out * (C / V)), 256, 0, st>>>((const T*)x, (T*)y, stats, B, S, int(C), x_sb, x_ss, adain,
                                                                               concatenate, scale, eps);
  });
  RB200_CHECK_LAUNCH("style_aligned");
  if (adain) count_launch();
  return 0;
}

int avg_pool_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t B, int H, int W, int C, int k) {
  if (!aligned16(x) || !aligned16(y)) RB200_FAIL(-1, "avg_pool2d: buffers must be 16-byte aligned");
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    if (C % V != 0) RB200_FAIL(-1, "avg_pool2d: C=%d must be a multiple of %d", C, V);
    avg_pool_kernel<T><<<ew_grid(B * (H / k) * (W / k) * (C / V)), 256, 0, st>>>((const T*)x, (T*)y, B, H, W, C, k);
  });
  RB200_CHECK_LAUNCH("avg_pool2d");
  return 0;
}

int pad_channels_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t B, int H, int W, int C, int Cp, int64_t sb, int64_t sc,
                      int64_t sh, int64_t sw) {
  if (!aligned16(y)) RB200_FAIL(-1, "pad_channels: output must be 16-byte aligned");
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    if (Cp % V != 0 || Cp < C) RB200_FAIL(-1, "pad_channels: Cp=%d must be >= C=%d and a multiple of %d", Cp, C, V);
    const int64_t pixels = B * H * W;
    pad_channels_kernel<T><<<ew_grid(pixels * (Cp / V)), 256, 0, st>>>((const T*)x, (T*)y, pixels, H, W, C, Cp, sb, sc, sh, sw);
  });
  RB200_CHECK_LAUNCH("pad_channels");
  return 0;
}

int patchify_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t B, int64_t H, int64_t W, int64_t C, int P, int64_t sb,
                  int64_t sc, int64_t sh, int64_t sw) {
  const int64_t Ho = H / P, Wo = W / P, total = B * Ho * Wo * P * P * C;
  DISPATCH_T(dtype, patchify_kernel<T><<<ew_grid(total), 256, 0, st>>>((const T*)x, (T*)y, total, int(Ho), int(Wo), int(C), P, sb, sc, sh, sw));
  RB200_CHECK_LAUNCH("patchify");
  return 0;
}

int window_impl(cudaStream_t st, int dtype, const void* x, void* y, int64_t B, int H, int W, int C, int ws, int merge) {
  const int nH = (H + ws - 1) / ws, nW = (W + ws - 1) / ws;
  if (!aligned16(x) || !aligned16(y)) RB200_FAIL(-1, "window: buffers must be 16-byte aligned");
  DISPATCH_T(dtype, {
    constexpr int V = 16 / sizeof(T);
    if (C % V != 0) RB200_FAIL(-1, "window: C=%d must be a multiple of %d", C, V);
    const int64_t rows = merge ? B * H * W : B * nH * nW * ws * ws;
    if (merge)
      window_kernel<T, true><<<ew_grid(rows * (C / V)), 256, 0, st>>>((const T*)x, (T*)y, B, H, W, C, ws, nH, nW);
    else
      window_kernel<T, false><<<ew_grid(rows * (C / V)), 256, 0, st>>>((const T*)x, (T*)y, B, H, W, C, ws, nH, nW);
  });
  RB200_CHECK_LAUNCH("window");
  return 0;
}

int conv_pack_impl(cudaStream_t st, int dtype, const void* w, void* out, int64_t Cout, int64_t Cin, int R, int S) {
  DISPATCH_T(dtype, conv_pack_kernel<T><<<ew_grid(Cout * Cin * R * S), 256, 0, st>>>((const T*)w, (T*)out, Cout, Cin, R * S));
  RB200_CHECK_LAUNCH("conv_pack");
  return 0;
}

int geglu_pack_impl(cudaStream_t st, int dtype, const void* w, const void* bias, void* wp, void* bp, int64_t F, int64_t K) {
  if (F % 16 != 0) RB200_FAIL(-1, "geglu_pack: F=%lld must be a multiple of 16", (long long)F);
  DISPATCH_T(dtype, geglu_pack_kernel<T><<<ew_grid(2 * F * K), 256, 0, st>>>((const T*)w, (const T*)bias, (T*)wp, (T*)bp, F, K));
  RB200_CHECK_LAUNCH("geglu_pack");
  return 0;
}

int lora_pack_impl(cudaStream_t st, int dtype, int n_lora, const rb200_lora* descs, int64_t N, int64_t K,
                   void* down_cat, void* up_cat, float* colscale, int r_pad) {
  if (n_lora < 1 || n_lora > 8) RB200_FAIL(-1, "lora_pack: 1..8 LoRAs per layer supported, got %d", n_lora);
  LoraList list;
  int total = 0;
  for (int i = 0; i < n_lora; ++i) {
    list.l[i] = descs[i];
    total += descs[i].rank;
  }
  if (total > r_pad || r_pad % 64 != 0) RB200_FAIL(-1, "lora_pack: r_pad=%d must be a multiple of 64 and >= %d", r_pad, total);
  const int64_t work = (N > K ? N : K) * r_pad;
  DISPATCH_T(dtype, lora_pack_kernel<T><<<ew_grid(work), 256, 0, st>>>(list, n_lora, N, K, (T*)down_cat, (T*)up_cat, colscale, r_pad));
  RB200_CHECK_LAUNCH("lora_pack");
  return 0;
}

}  // namespace rb200
