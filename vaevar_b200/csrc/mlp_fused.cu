// Host side of the fused tower MLP (mlp_fused.cuh): tensor maps, dispatch over the token width, kernel-level test hooks.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "engine.h"
#include "mlp_fused.cuh"

namespace vv {

static unsigned long long* g_mlp_trace = nullptr;

bool mlp_fused_supported(int D, int rows) {
  return (D == 64 || D == 96 || D == 128 || D == 192) && rows > 0 && rows % 128 == 0;
}

// W1: [batch][4D][D] (N x K, K-major) -- forward fc1 (gamma folded in), backward W2^T;  W2: [batch][D][4D] -- forward fc2, backward W1^T.
// u: [batch][rows][4D] saved gelu'(u) (written by the forward kernel, read by the backward one);  dy16: backward only, [batch][rows][D] bf16.
const char* make_mlp_desc(MlpDesc* d, int D, bool bwd, const bf16* W1, const bf16* W2, const bf16* u, const bf16* dy16, const MlpArgs& args) {
  if (!mlp_fused_supported(D, args.rows)) return "fused MLP: token width / row count not supported";
  d->D = D; d->bwd = bwd ? 1 : 0; d->mode = bwd ? MLP_BWD : MLP_FWD; d->a = args;
  d->a.trace = g_mlp_trace;
  const unsigned long long wb = 4ull * D * D * args.batch * 2;                     // bytes of one of the two weight blocks
  const bool adj = reinterpret_cast<const char*>(W2) == reinterpret_cast<const char*>(W1) + wb;
  d->w_ptr = adj ? W1 : nullptr; d->w_bytes = adj ? 2 * wb : 0;
  const long long rows = args.rows, B = args.batch;
  const char* e;
  if ((e = encode_tma_3d_16(&d->tmW1, W1, D, 4LL * D, B, D, 4LL * D * D, 64, MLP_HC))) return e;
  if ((e = encode_tma_3d_16(&d->tmW2, W2, 4LL * D, D, B, 4LL * D, 4LL * D * D, 64, D))) return e;
  if (bwd) {
    if ((e = encode_tma_3d_16(&d->tmA, dy16, D, rows, B, D, rows * D, 64, 128))) return e;
    if ((e = encode_tma_3d_16(&d->tmU, u, 4LL * D, rows, B, 4LL * D, rows * 4 * D, 64, 128))) return e;
  } else {
    d->tmA = d->tmW1;                                                             // unused
    if ((e = encode_tma_3d_16(&d->tmU, u, 4LL * D, rows, B, 4LL * D, rows * 4 * D, 16, 32))) return e;   // per-warp 32 x 16 store boxes, dense rows
  }
  return nullptr;
}

template <int D, int MODE, bool F16>
static void launch_mlp_t(const MlpDesc& d, cudaStream_t s) {
  using L = MlpSmem<D, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mlp_fused_kernel<D, MODE, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    attr_set = true;
  }
  const int tiles = (d.a.rows / 128) * d.a.batch;
  launch_kernel(mlp_fused_kernel<D, MODE, F16>, dim3(std::min(tiles, num_sms())), dim3(MLP_THREADS), L::TOTAL, s, d.tmW1, d.tmW2, d.tmA, d.tmU, d.a);
}

template <int D>
static void launch_mlp_d(const MlpDesc& d, cudaStream_t s) {
  if (d.mode == MLP_BWD) launch_mlp_t<D, MLP_BWD, false>(d, s);
  else if (d.mode == MLP_LIN) { if (d.a.f16) launch_mlp_t<D, MLP_LIN, true>(d, s); else launch_mlp_t<D, MLP_LIN, false>(d, s); }
  else if (d.a.f16) launch_mlp_t<D, MLP_FWD, true>(d, s);
  else launch_mlp_t<D, MLP_FWD, false>(d, s);
}

void launch_mlp(const MlpDesc& d, cudaStream_t s) {
  switch (d.D) {
    case 64: launch_mlp_d<64>(d, s); break;
    case 96: launch_mlp_d<96>(d, s); break;
    case 128: launch_mlp_d<128>(d, s); break;
    case 192: launch_mlp_d<192>(d, s); break;
    default: break;
  }
}

// LayerNorm + ONE Linear of a tower block (norm1 -> qkv, swinblock.py:268-269 + :139) on the same kernel: out16 [batch][rows][n_out] =
// 16bit(normalise(x) W^T + bias) with W (batch, n_out, D) carrying gamma and bias the folded beta.
const char* make_lin_desc(MlpDesc* d, int D, const bf16* W, bf16* out16, long long ld_out, long long out_bs, const MlpArgs& args) {
  if (!mlp_fused_supported(D, args.rows)) return "fused LayerNorm + Linear: token width / row count not supported";
  if (args.n_out <= 0 || args.n_out % 16 || args.n_out > 5 * D) return "fused LayerNorm + Linear: n_out must be a multiple of 16, at most 5 D";
  d->D = D; d->bwd = 0; d->mode = MLP_LIN; d->a = args;
  d->a.trace = nullptr;
  const long long rows = args.rows, B = args.batch;
  const char* e;
  if ((e = encode_tma_3d_16(&d->tmW1, W, D, args.n_out, B, D, (long long)args.n_out * D, 64, MLP_HC))) return e;
  d->tmW2 = d->tmW1; d->tmA = d->tmW1;                                                  // unused
  if ((e = encode_tma_3d_16(&d->tmU, out16, args.n_out, rows, B, ld_out, out_bs, 16, 32))) return e;
  d->w_ptr = W; d->w_bytes = (unsigned long long)args.n_out * D * B * 2;
  return nullptr;
}

}  // namespace vv

using namespace vv;

extern "C" {

// Debug: fused-MLP launches built after this call stamp clock64 values of CTA 0 into trace_dev (128 x uint64; tools/mlp_trace.py); null = off.
VV_API int vv_debug_mlp_trace(void* trace_dev) {
  g_mlp_trace = (unsigned long long*)trace_dev;
  return 0;
}

// Kernel-level hook: out = x1 + fc2(gelu(fc1(normalise(x1)))) with W1 (4D x D) / b1 standing for the gamma- / beta-folded fc1; u_out
// receives gelu'(u).  16-bit buffers are fp16 when f16 != 0, else bf16.
VV_API int vv_test_mlp_fwd(const float* x1, const void* W1, const void* W2, const float* b1, const float* b2, int rows, int batch, int D, int f16,
                           float eps, void* u_out, float* out_f32, void* out16, const float* shift, float* stats_out, void* stream) {
  if (!mlp_fused_supported(D, rows)) { set_error("vv_test_mlp_fwd: D must be 64 / 96 / 128 / 192 and rows a multiple of 128"); return -2; }
  MlpArgs a{};
  a.rows = rows; a.batch = batch; a.f16 = f16; a.eps = eps; a.x1 = x1; a.b1 = b1; a.b2 = b2; a.out_f32 = out_f32;
  a.out16 = (bf16*)out16; a.ld16 = D; a.bs16 = (long long)rows * D; a.shift = shift; a.stats_out = stats_out; a.u_out = (bf16*)u_out;
  MlpDesc d;
  const char* er = make_mlp_desc(&d, D, false, (const bf16*)W1, (const bf16*)W2, (const bf16*)u_out, nullptr, a);
  if (er) { set_error("%s", er); return -2; }
  launch_mlp(d, (cudaStream_t)stream);
  if (cudaGetLastError() != cudaSuccess) { set_error("vv_test_mlp_fwd: launch failed"); return -1; }
  return 0;
}

// Kernel-level hook: dx = LN^T((dy W2T^T . u) W1T^T) + dres with W2T (4D x D) = fc2.weight^T, W1T (D x 4D) = fc1.weight^T, both bf16.
VV_API int vv_test_mlp_bwd(const void* dy16, const void* u, const float* x1, const void* W2T, const void* W1T, const float* gamma, const float* dres,
                           int rows, int batch, int D, int f16, float eps, float* dx, void* dx16, void* stream) {
  if (!mlp_fused_supported(D, rows)) { set_error("vv_test_mlp_bwd: D must be 64 / 96 / 128 / 192 and rows a multiple of 128"); return -2; }
  MlpArgs a{};
  a.rows = rows; a.batch = batch; a.f16 = f16; a.eps = eps; a.x1 = x1; a.gamma = gamma; a.dres = dres; a.dx = dx; a.dx16 = (bf16*)dx16;
  MlpDesc d;
  const char* er = make_mlp_desc(&d, D, true, (const bf16*)W2T, (const bf16*)W1T, (const bf16*)u, (const bf16*)dy16, a);
  if (er) { set_error("%s", er); return -2; }
  launch_mlp(d, (cudaStream_t)stream);
  if (cudaGetLastError() != cudaSuccess) { set_error("vv_test_mlp_bwd: launch failed"); return -1; }
  return 0;
}

// Kernel-level hook: out_16 (batch, rows, n_out) = 16bit(normalise(x) W^T + bias), W (batch, n_out, D) 16-bit, bias (batch, n_out) fp32.
VV_API int vv_test_lin_fwd(const float* x, const void* W, const float* bias, int rows, int batch, int D, int n_out, int f16, float eps,
                           void* out16, void* stream) {
  if (!mlp_fused_supported(D, rows)) { set_error("vv_test_lin_fwd: D must be 64 / 96 / 128 / 192 and rows a multiple of 128"); return -2; }
  MlpArgs a{};
  a.rows = rows; a.batch = batch; a.f16 = f16; a.eps = eps; a.x1 = x; a.b1 = bias; a.n_out = n_out;
  MlpDesc d;
  const char* er = make_lin_desc(&d, D, (const bf16*)W, (bf16*)out16, n_out, (long long)rows * n_out, a);
  if (er) { set_error("%s", er); return -2; }
  launch_mlp(d, (cudaStream_t)stream);
  if (cudaGetLastError() != cudaSuccess) { set_error("vv_test_lin_fwd: launch failed"); return -1; }
  return 0;
}

}  // extern "C"
