// extern "C" entry points of librefiners_b200.so (declared in include/refiners_b200.h):
// argument checking, kernel-family dispatch (tcgen05 where the shape allows, CUDA-core
// kernels otherwise), error text, launch accounting.  No entry point allocates or synchronises.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace rb200 {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_mode{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int kernel_mode() { return g_mode.load(std::memory_order_relaxed); }

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
  return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

// SM count of the CURRENT device (cached per ordinal: a process may drive several GPUs)
int sm_count() {
  static std::atomic<int> cached[kMaxDevices];
  const int dev = current_device();
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  cached[dev].store(n, std::memory_order_relaxed);
  return n;
}

// implemented in norm_kernels.cu
int group_norm_impl(cudaStream_t, int, const void*, void*, int64_t, int64_t, int64_t, int, float, const void*, const void*, int, void*, size_t, float*, int);
size_t group_norm_ws(int64_t, int64_t, int);
int layer_norm_impl(cudaStream_t, int, const void*, void*, int64_t, int64_t, float, const void*, const void*);
int unary_impl(cudaStream_t, int, const void*, void*, int64_t, int);
int add_impl(cudaStream_t, int, const void*, const void*, void*, int64_t, float);
int geglu_impl(cudaStream_t, int, const void*, void*, int64_t, int64_t);
int cfg_scale_input_impl(cudaStream_t, int, const void*, void*, int64_t, const void*, int);
int cfg_euler_impl(cudaStream_t, int, const void*, const void*, void*, int64_t, const void*, float, int);
int patchify_impl(cudaStream_t, int, const void*, void*, int64_t, int64_t, int64_t, int64_t, int, int64_t, int64_t, int64_t, int64_t);
int window_impl(cudaStream_t, int, const void*, void*, int64_t, int, int, int, int, int);
int pad_channels_impl(cudaStream_t, int, const void*, void*, int64_t, int, int, int, int, int64_t, int64_t, int64_t, int64_t);
int concat_channels_impl(cudaStream_t, int, int, const void* const*, const int*, void*, int64_t);
int resize_nearest_impl(cudaStream_t, int, const void*, void*, int64_t, int, int, int, int, int);
int avg_pool_impl(cudaStream_t, int, const void*, void*, int64_t, int, int, int, int);
int style_aligned_impl(cudaStream_t, int, const void*, void*, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, float, float, float*);
int conv_pack_impl(cudaStream_t, int, const void*, void*, int64_t, int64_t, int, int);
int geglu_pack_impl(cudaStream_t, int, const void*, const void*, void*, void*, int64_t, int64_t);
int lora_pack_impl(cudaStream_t, int, int, const rb200_lora*, int64_t, int64_t, void*, void*, float*, int);
// implemented in attn_probs.cu
int attention_probs_impl(cudaStream_t, int, const void*, const void*, void*, int64_t, int, int64_t, int64_t, int, int64_t, int64_t, int64_t, int64_t,
                         float);
// implemented in sam_attention.cu
int sam_attention_impl(cudaStream_t, int, const void*, const void*, const void*, void*, int64_t, int, int, int, int, void*, size_t);
size_t sam_attention_ws(int64_t, int, int, int, int);

static int run_gemm(cudaStream_t st, const GemmProblem& p) {
  if (kernel_mode() != 1 && tc_gemm_supported(p)) return tc_gemm(st, p);
  return simt_gemm(st, p);
}

static bool bad_dtype(int dtype) { return dtype != RB200_BF16 && dtype != RB200_FP16 && dtype != RB200_FP32; }

}  // namespace rb200

using namespace rb200;

extern "C" {

int rb200_abi_version(void) { return RB200_ABI_VERSION; }
const char* rb200_last_error(void) { return g_error; }
int64_t rb200_launch_count(void) { return g_launches.load(); }
int rb200_set_kernel_mode(int mode) { return g_mode.exchange(mode); }

int rb200_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) RB200_FAIL(-2, "no CUDA device");
  cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev);
  return 0;
}

size_t rb200_linear_workspace_bytes(int64_t M, int r_pad, int dtype) {
  if (r_pad <= 0) return 0;
  return size_t(M) * size_t(r_pad) * dtype_size(dtype) + 256;
}

int rb200_linear(void* stream, int dtype, const void* x, int64_t ldx, const void* w, int64_t ldw, const void* bias, void* y,
                 int64_t ldy, int64_t M, int64_t N, int64_t K, int r_pad, const void* lora_down_cat, const void* lora_up_cat,
                 const float* lora_colscale, const void* residual, int64_t ldr, int epilogue, void* ws, size_t ws_bytes) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bad_dtype(dtype)) RB200_FAIL(-1, "linear: bad dtype %d", dtype);
  if (M < 0 || N <= 0 || K <= 0) RB200_FAIL(-1, "linear: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  if (!x || !w || !y) RB200_FAIL(-1, "linear: null operand");
  if (epilogue < RB200_EPI_NONE || epilogue > RB200_EPI_SILU) RB200_FAIL(-1, "linear: bad epilogue %d", epilogue);
  if (M == 0) return 0;
  GemmProblem p{};
  p.dtype = dtype;
  p.a = x; p.lda = ldx;
  p.b = w; p.ldb = ldw;
  p.bias = bias;
  p.residual = residual; p.ldr = ldr;
  p.y = y; p.ldy = ldy;
  p.M = M; p.N = N; p.K = K;
  p.epilogue = epilogue;
  if (r_pad > 0) {
    if (!lora_down_cat || !lora_up_cat || !lora_colscale) RB200_FAIL(-1, "linear: LoRA operands missing");
    if (r_pad % 64 != 0) RB200_FAIL(-1, "linear: r_pad=%d must be a multiple of 64", r_pad);
    const size_t need = rb200_linear_workspace_bytes(M, r_pad, dtype);
    if (!ws || ws_bytes < need) RB200_FAIL(-1, "linear: workspace %zu < %zu", ws_bytes, need);
    void* t = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
    // T[M, r_pad] = (X down_cat^T) * colscale   (fp32 accumulate, stored in the activation dtype)
    GemmProblem d{};
    d.dtype = dtype;
    d.a = x; d.lda = ldx;
    d.b = lora_down_cat; d.ldb = K;
    d.colscale = lora_colscale;
    d.y = t; d.ldy = r_pad;
    d.M = M; d.N = r_pad; d.K = K;
    d.epilogue = RB200_EPI_NONE;
    if (int rc = run_gemm(st, d)) return rc;
    p.a2 = t; p.lda2 = r_pad;
    p.b2 = lora_up_cat; p.ldb2 = r_pad;
    p.K2 = r_pad;
  }
  return run_gemm(st, p);
}

int rb200_lora_pack(void* stream, int dtype, int n_lora, const rb200_lora* loras, int64_t N, int64_t K, void* down_cat,
                    void* up_cat, float* colscale, int r_pad) {
  if (bad_dtype(dtype) || !loras || !down_cat || !up_cat || !colscale) RB200_FAIL(-1, "lora_pack: bad arguments");
  return lora_pack_impl(static_cast<cudaStream_t>(stream), dtype, n_lora, loras, N, K, down_cat, up_cat, colscale, r_pad);
}

int rb200_geglu_pack(void* stream, int dtype, const void* w, const void* bias, void* w_packed, void* bias_packed, int64_t F,
                     int64_t K) {
  if (bad_dtype(dtype) || !w || !w_packed || (bias && !bias_packed)) RB200_FAIL(-1, "geglu_pack: bad arguments");
  return geglu_pack_impl(static_cast<cudaStream_t>(stream), dtype, w, bias, w_packed, bias_packed, F, K);
}

int rb200_conv2d_pack_weight(void* stream, int dtype, const void* w, void* w_packed, int64_t Cout, int64_t Cin, int R, int S) {
  if (bad_dtype(dtype) || !w || !w_packed || R < 1 || S < 1) RB200_FAIL(-1, "conv2d_pack_weight: bad arguments");
  return conv_pack_impl(static_cast<cudaStream_t>(stream), dtype, w, w_packed, Cout, Cin, R, S);
}

int rb200_conv2d(void* stream, int dtype, const void* x, const void* w_packed, const void* bias, const void* chan_bias,
                 const void* residual, void* y, int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int R, int S,
                 int stride, int pad, int epilogue) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bad_dtype(dtype)) RB200_FAIL(-1, "conv2d: bad dtype %d", dtype);
  if (!x || !w_packed || !y) RB200_FAIL(-1, "conv2d: null operand");
  if (B < 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || R < 1 || S < 1 || stride < 1 || pad < 0)
    RB200_FAIL(-1, "conv2d: bad geometry");
  if (epilogue != RB200_EPI_NONE && epilogue != RB200_EPI_GELU && epilogue != RB200_EPI_SILU) RB200_FAIL(-1, "conv2d: bad epilogue %d", epilogue);
  const int64_t Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  if (Ho <= 0 || Wo <= 0) RB200_FAIL(-1, "conv2d: empty output");
  if (B == 0) return 0;
  GemmProblem p{};
  p.dtype = dtype;
  p.conv = 1;
  p.a = x;
  p.b = w_packed;
  p.bias = bias;
  p.chan_bias = chan_bias;
  p.residual = residual; p.ldr = Cout;
  p.y = y; p.ldy = Cout;
  p.M = B * Ho * Wo; p.N = Cout; p.K = int64_t(R) * S * Cin;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Ho = Ho; p.Wo = Wo;
  p.R = R; p.S = S; p.stride = stride; p.pad = pad;
  p.epilogue = epilogue;
  return run_gemm(st, p);
}

size_t rb200_group_norm_workspace_bytes(int64_t B, int64_t HW, int G) { return group_norm_ws(B, HW, G) + 256; }

int rb200_group_norm(void* stream, int dtype, const void* x, void* y, int64_t B, int64_t HW, int64_t C, int G, float eps,
                     const void* gamma, const void* beta, int silu, void* ws, size_t ws_bytes) {
  if (bad_dtype(dtype) || !x || !y || !gamma || !beta) RB200_FAIL(-1, "group_norm: bad arguments");
  if (B <= 0 || HW <= 0 || C <= 0) return 0;
  return group_norm_impl(static_cast<cudaStream_t>(stream), dtype, x, y, B, HW, C, G, eps, gamma, beta, silu, ws, ws_bytes, nullptr, 0);
}

int rb200_group_norm_fixed(void* stream, int dtype, const void* x, void* y, int64_t B, int64_t HW, int64_t C, int G, float eps,
                           const void* gamma, const void* beta, int silu, void* ws, size_t ws_bytes, float* stats, int frozen) {
  if (bad_dtype(dtype) || !x || !y || !gamma || !beta || !stats) RB200_FAIL(-1, "group_norm_fixed: bad arguments");
  if (B <= 0 || HW <= 0 || C <= 0) return 0;
  return group_norm_impl(static_cast<cudaStream_t>(stream), dtype, x, y, B, HW, C, G, eps, gamma, beta, silu, ws, ws_bytes, stats,
                         frozen != 0);
}

int rb200_layer_norm(void* stream, int dtype, const void* x, void* y, int64_t rows, int64_t C, float eps, const void* gamma,
                     const void* beta) {
  if (bad_dtype(dtype) || !x || !y || !gamma || !beta) RB200_FAIL(-1, "layer_norm: bad arguments");
  if (rows <= 0 || C <= 0) return 0;
  return layer_norm_impl(static_cast<cudaStream_t>(stream), dtype, x, y, rows, C, eps, gamma, beta);
}

int rb200_unary(void* stream, int dtype, const void* x, void* y, int64_t n, int op) {
  if (bad_dtype(dtype) || !x || !y || op < 0 || op > RB200_UNARY_SIGMOID) RB200_FAIL(-1, "unary: bad arguments");
  if (n <= 0) return 0;
  return unary_impl(static_cast<cudaStream_t>(stream), dtype, x, y, n, op);
}

int rb200_geglu(void* stream, int dtype, const void* x, void* y, int64_t rows, int64_t F) {
  if (bad_dtype(dtype) || !x || !y) RB200_FAIL(-1, "geglu: bad arguments");
  if (rows <= 0 || F <= 0) return 0;
  return geglu_impl(static_cast<cudaStream_t>(stream), dtype, x, y, rows, F);
}

int rb200_cfg_scale_input(void* stream, int dtype, const void* x, void* y, int64_t n, const void* sigma, int twice) {
  if (bad_dtype(dtype) || !x || !y || !sigma) RB200_FAIL(-1, "cfg_scale_input: bad arguments");
  if (n <= 0) return 0;
  return cfg_scale_input_impl(static_cast<cudaStream_t>(stream), dtype, x, y, n, sigma, twice);
}

int rb200_cfg_euler(void* stream, int dtype, const void* x, const void* eps, void* y, int64_t n, const void* sigmas, float condition_scale,
                    int guided) {
  if (bad_dtype(dtype) || !x || !eps || !y || !sigmas) RB200_FAIL(-1, "cfg_euler: bad arguments");
  if (n <= 0) return 0;
  return cfg_euler_impl(static_cast<cudaStream_t>(stream), dtype, x, eps, y, n, sigmas, condition_scale, guided);
}

int rb200_add(void* stream, int dtype, const void* a, const void* b, void* y, int64_t n, float alpha) {
  if (bad_dtype(dtype) || !a || !b || !y) RB200_FAIL(-1, "add: bad arguments");
  if (n <= 0) return 0;
  return add_impl(static_cast<cudaStream_t>(stream), dtype, a, b, y, n, alpha);
}

int rb200_pad_channels(void* stream, int dtype, const void* x, void* y, int64_t B, int H, int W, int C, int Cp, int64_t sb, int64_t sc,
                       int64_t sh, int64_t sw) {
  if (bad_dtype(dtype) || !x || !y) RB200_FAIL(-1, "pad_channels: bad arguments");
  if (H < 1 || W < 1 || C < 1) RB200_FAIL(-1, "pad_channels: bad geometry");
  if (B <= 0) return 0;
  return pad_channels_impl(static_cast<cudaStream_t>(stream), dtype, x, y, B, H, W, C, Cp, sb, sc, sh, sw);
}

int rb200_concat_channels(void* stream, int dtype, int n, const void* const* srcs, const int* channels, void* y, int64_t pixels) {
  if (bad_dtype(dtype) || !srcs || !channels || !y) RB200_FAIL(-1, "concat_channels: bad arguments");
  if (pixels <= 0) return 0;
  return concat_channels_impl(static_cast<cudaStream_t>(stream), dtype, n, srcs, channels, y, pixels);
}

size_t rb200_style_aligned_workspace_bytes(int64_t B, int64_t C) { return size_t(B) * size_t(C) * 2 * sizeof(float); }

int rb200_style_aligned(void* stream, int dtype, const void* x, void* y, int64_t B, int64_t S, int64_t C, int64_t x_sb, int64_t x_ss, int adain,
                        int concatenate, float scale, float eps, void* ws, size_t ws_bytes) {
  if (bad_dtype(dtype) || !x || !y) RB200_FAIL(-1, "style_aligned: bad arguments");
  if (B <= 0 || S <= 0 || C <= 0) return 0;
  if (B % 2 != 0) RB200_FAIL(-1, "style_aligned: the batch (%lld) must be the two halves of a classifier-free-guidance batch", (long long)B);
  if (adain && (!ws || ws_bytes < rb200_style_aligned_workspace_bytes(B, C))) RB200_FAIL(-1, "style_aligned: workspace too small");
  return style_aligned_impl(static_cast<cudaStream_t>(stream), dtype, x, y, B, S, C, x_sb, x_ss, adain, concatenate, scale, eps,
                            static_cast<float*>(ws));
}

int rb200_avg_pool2d(void* stream, int dtype, const void* x, void* y, int64_t B, int H, int W, int C, int k) {
  if (bad_dtype(dtype) || !x || !y) RB200_FAIL(-1, "avg_pool2d: bad arguments");
  if (k < 1 || H < k || W < k || C < 1) RB200_FAIL(-1, "avg_pool2d: bad geometry");
  if (B <= 0) return 0;
  return avg_pool_impl(static_cast<cudaStream_t>(stream), dtype, x, y, B, H, W, C, k);
}

int rb200_resize_nearest(void* stream, int dtype, const void* x, void* y, int64_t B, int H, int W, int C, int Ho, int Wo) {
  if (bad_dtype(dtype) || !x || !y) RB200_FAIL(-1, "resize_nearest: bad arguments");
  if (H < 1 || W < 1 || C < 1 || Ho < 1 || Wo < 1) RB200_FAIL(-1, "resize_nearest: bad geometry");
  if (B <= 0) return 0;
  return resize_nearest_impl(static_cast<cudaStream_t>(stream), dtype, x, y, B, H, W, C, Ho, Wo);
}

int rb200_patchify(void* stream, int dtype, const void* x, void* y, int64_t B, int64_t H, int64_t W, int64_t C, int P, int64_t sb,
                   int64_t sc, int64_t sh, int64_t sw) {
  if (bad_dtype(dtype) || !x || !y) RB200_FAIL(-1, "patchify: bad arguments");
  if (P < 1 || H % P != 0 || W % P != 0 || C < 1) RB200_FAIL(-1, "patchify: %lldx%lld is not a multiple of the patch size %d", (long long)H, (long long)W, P);
  if (B <= 0) return 0;
  return patchify_impl(static_cast<cudaStream_t>(stream), dtype, x, y, B, H, W, C, P, sb, sc, sh, sw);
}

int rb200_window_partition(void* stream, int dtype, const void* x, void* y, int64_t B, int H, int W, int C, int window, int merge) {
  if (bad_dtype(dtype) || !x || !y) RB200_FAIL(-1, "window_partition: bad arguments");
  if (window < 1 || H < 1 || W < 1 || C < 1) RB200_FAIL(-1, "window_partition: bad geometry");
  if (B <= 0) return 0;
  return window_impl(static_cast<cudaStream_t>(stream), dtype, x, y, B, H, W, C, window, merge);
}

int rb200_sdpa(void* stream, int dtype, const void* q, const void* k, const void* v, void* o, int64_t B, int H, int64_t Sq,
               int64_t Sk, int D, int64_t q_sb, int64_t q_ss, int64_t k_sb, int64_t k_ss, int64_t v_sb, int64_t v_ss,
               int64_t o_sb, int64_t o_ss, float scale, int is_causal, const void* k2, const void* v2, int64_t Sk2,
               int64_t k2_sb, int64_t k2_ss, int64_t v2_sb, int64_t v2_ss, float scale2) {
  if (bad_dtype(dtype) || !q || !k || !v || !o) RB200_FAIL(-1, "sdpa: bad arguments");
  if ((k2 == nullptr) != (v2 == nullptr)) RB200_FAIL(-1, "sdpa: k2 and v2 must come together");
  if (H <= 0 || D <= 0 || Sk < 0) RB200_FAIL(-1, "sdpa: bad shape");
  SdpaProblem p{};
  p.dtype = dtype;
  p.q = q; p.k = k; p.v = v; p.o = o;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.D = D;
  p.q_sb = q_sb; p.q_ss = q_ss; p.k_sb = k_sb; p.k_ss = k_ss; p.v_sb = v_sb; p.v_ss = v_ss; p.o_sb = o_sb; p.o_ss = o_ss;
  p.scale = scale; p.causal = is_causal;
  p.k2 = k2; p.v2 = v2; p.Sk2 = Sk2; p.k2_sb = k2_sb; p.k2_ss = k2_ss; p.v2_sb = v2_sb; p.v2_ss = v2_ss; p.scale2 = scale2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kernel_mode() != 1) {
    if (tc_sdpa_short_supported(p)) return tc_sdpa_short(st, p);
    if (tc_sdpa2_supported(p)) return tc_sdpa2(st, p);
    if (tc_sdpa_supported(p)) return tc_sdpa(st, p);
  }
  return simt_sdpa(st, p);
}

int rb200_attention_probs(void* stream, int dtype, const void* q, const void* k, void* probs, int64_t B, int H, int64_t Sq, int64_t Sk, int D,
                          int64_t q_sb, int64_t q_ss, int64_t k_sb, int64_t k_ss, float scale) {
  if (bad_dtype(dtype) || !q || !k || !probs) RB200_FAIL(-1, "attention_probs: bad arguments");
  if (H <= 0 || D <= 0 || Sk <= 0 || Sq < 0) RB200_FAIL(-1, "attention_probs: bad shape");
  if (B <= 0 || Sq == 0) return 0;
  return attention_probs_impl(static_cast<cudaStream_t>(stream), dtype, q, k, probs, B, H, Sq, Sk, D, q_sb, q_ss, k_sb, k_ss, scale);
}

size_t rb200_sam_attention_workspace_bytes(int64_t Bw, int Hh, int Ww, int heads, int d) { return sam_attention_ws(Bw, Hh, Ww, heads, d); }

int rb200_sam_attention(void* stream, int dtype, const void* qkv, const void* rel_h_emb, const void* rel_w_emb, void* o,
                        int64_t Bw, int Hh, int Ww, int heads, int d, void* ws, size_t ws_bytes) {
  if (bad_dtype(dtype) || !qkv || !rel_h_emb || !rel_w_emb || !o) RB200_FAIL(-1, "sam_attention: bad arguments");
  if (Bw <= 0) return 0;
  return sam_attention_impl(static_cast<cudaStream_t>(stream), dtype, qkv, rel_h_emb, rel_w_emb, o, Bw, Hh, Ww, heads, d, ws, ws_bytes);
}

}  // extern "C"
