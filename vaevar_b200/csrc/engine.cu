// The 4D-Var cost-and-gradient engine: weight packing, per-application launch plans (forward and hand-derived
// input-VJP of LGUnet_all, networks_old/transformer.py:747-752), the cost J(z) and grad_z J of
// da_4dvar.py:1183-1208 / 1242-1246, and the C ABI of include/vaevar.h.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <nvtx3/nvToolsExt.h>

#include "engine.h"
#include "seams.cuh"

namespace vv {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
#define VV_CUDA(x)                                                                                   \
  do {                                                                                               \
    cudaError_t _e = (x);                                                                            \
    if (_e != cudaSuccess) {                                                                         \
      set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(_e), __FILE__, __LINE__);            \
      return -1;                                                                                     \
    }                                                                                                \
  } while (0)
#define VV_CHECK(cond, ...)      \
  do {                           \
    if (!(cond)) {               \
      set_error(__VA_ARGS__);    \
      return -2;                 \
    }                            \
  } while (0)

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
// dst: the forward operand (fp16 when the forward pass runs in fp16, else bf16); dstT: its transpose for the
// input-gradient GEMMs, always bf16 (gradients need bf16's range).
__global__ void pack_w_kernel(bf16* dst, bf16* dstT, const float* src, int rows, int cols, int f16) {
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float w = src[i];
    if (f16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(w);
    else dst[i] = __float2bfloat16(w);
    if (dstT) dstT[(long long)c * rows + r] = __float2bfloat16(w);
  }
}
// LayerNorm folded into the following Linear (see GemmArgs::ln_stats): one warp per output row n of W (N x K, fp32).
//   dst[n,k]  = 16bit(W[n,k] gamma[k])            (forward operand; the gradient GEMMs keep the unfolded W^T)
//   colsum[n] = sum_k float(dst[n,k])             (of the ROUNDED values: the mean term must cancel what the MMA summed)
//   cbias[n]  = bias[n] + sum_k beta[k] W[n,k]
__global__ void fold_ln_kernel(bf16* dst, float* colsum, float* cbias, const float* W, const float* gamma, const float* beta,
                               const float* bias, int rows, int cols, int f16) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= rows) return;
  float s = 0.f, c = 0.f;
  for (int k = lane; k < cols; k += 32) {
    const float w = W[(long long)n * cols + k];
    const float wf = w * gamma[k];
    float r;
    if (f16) { const __half h = __float2half_rn(wf); reinterpret_cast<__half*>(dst)[(long long)n * cols + k] = h; r = __half2float(h); }
    else { const bf16 h = __float2bfloat16(wf); dst[(long long)n * cols + k] = h; r = __bfloat162float(h); }
    s += r;
    c = fmaf(beta[k], w, c);
  }
  s = warp_sum(s); c = warp_sum(c);
  if (lane == 0) { colsum[n] = s; cbias[n] = bias[n] + c; }
}
__global__ void finalize_J_kernel(const double* dots, const double* jobs, float coeff, double* out) {
  const double jr = 0.5 * dots[0];
  out[1] = jr;
  out[2] = jobs[0];
  out[0] = jr + (double)coeff * jobs[0];
}

void launch_pack_w(bf16* dst, bf16* dstT, const float* src, int rows, int cols, int f16) {
  pack_w_kernel<<<592, 256>>>(dst, dstT, src, rows, cols, f16);
}
void launch_fold_ln(bf16* dst, float* colsum, float* cbias, const float* W, const float* gamma, const float* beta, const float* bias,
                    int rows, int cols, int f16) {
  fold_ln_kernel<<<(rows + 7) / 8, 256>>>(dst, colsum, cbias, W, gamma, beta, bias, rows, cols, f16);
}

bool nvtx_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VV_NVTX");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
void nvtx_push(const char* name) { if (nvtx_enabled()) nvtxRangePushA(name); }
void nvtx_pop() { if (nvtx_enabled()) nvtxRangePop(); }
static const char* op_family(Op::Kind k) {
  switch (k) {
    case Op::GEMM: return "gemm (tcgen05)";
    case Op::LN_F: return "layernorm fwd";
    case Op::LN_B: return "layernorm bwd";
    case Op::ATT_F: return "window attention fwd";
    case Op::ATT_B: return "window attention bwd";
    case Op::P2T: return "patch pixels->tokens";
    case Op::T2P: return "patch tokens->pixels";
    case Op::ROPE: return "rope2";
    case Op::ATT1: return "SD_attn";
    case Op::PE32: return "patch embed (3,2)";
    case Op::CT32: return "conv-transpose head (3,2)";
    case Op::MLP_F: return "fused tower MLP fwd (tcgen05)";
    case Op::MLP_B: return "fused tower MLP bwd (tcgen05)";
    case Op::LIN_F: return "fused tower norm1 + qkv (tcgen05)";
  }
  return "op";
}

int Plan::run(cudaStream_t s) const {
  const bool nv = nvtx_enabled();
  if (nv && !label.empty()) nvtxRangePushA(label.c_str());
  for (const Op& o : ops) {
    if (nv) nvtxRangePushA(op_family(o.kind));
    switch (o.kind) {
      case Op::GEMM: launch_gemm(o.gemm, s); break;
      case Op::LN_F: launch_ln_fwd(o.lnf, s); break;
      case Op::LN_B: launch_ln_bwd(o.lnb, s); break;
      case Op::ATT_F: launch_attn_fwd(o.att, s); break;
      case Op::ATT_B: launch_attn_bwd(o.att, s); break;
      case Op::P2T: launch_p2t(o.patch, s); break;
      case Op::T2P: launch_t2p(o.patch, s); break;
      case Op::ROPE: launch_rope(o.rope, s); break;
      case Op::ATT1: launch_attn1(o.att1, s); break;
      case Op::PE32: launch_patch32(o.pe32, s); break;
      case Op::CT32: launch_convt32(o.ct32, s); break;
      case Op::MLP_F: case Op::MLP_B: case Op::LIN_F: launch_mlp(o.mlp, s); break;
    }
    if (nv) nvtxRangePop();
  }
  if (nv && !label.empty()) nvtxRangePop();
  return (int)ops.size();
}

bool p2t_supported(int D);

}  // namespace vv

using namespace vv;

// ---------------------------------------------------------------------------------------------
// allocation
// ---------------------------------------------------------------------------------------------
template <typename T>
static T* dalloc(vv_engine* e, size_t n) {
  void* p = nullptr;
  if (n == 0) n = 1;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  e->allocs.push_back(p);
  return static_cast<T*>(p);
}
// Release a buffer obtained from dalloc (re-allocation of per-case buffers that outgrew their capacity).
template <typename T>
static void dfree(vv_engine* e, T*& p) {
  if (!p) return;
  auto it = std::find(e->allocs.begin(), e->allocs.end(), static_cast<void*>(p));
  if (it != e->allocs.end()) e->allocs.erase(it);
  cudaFree(p);
  p = nullptr;
}
template <typename T>
static T* dupload(vv_engine* e, const std::vector<T>& h) {
  T* d = dalloc<T>(e, h.size());
  if (d && !h.empty()) cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

// ---------------------------------------------------------------------------------------------
// network geometry
// ---------------------------------------------------------------------------------------------
static int net_init(Net& n, const vv_net_config& c) {
  n.c = c;
  n.G = c.n_groups; n.D = c.enc_dim; n.E = c.embed_dim; n.H = c.img_h; n.W = c.img_w;
  VV_CHECK(n.G >= 1 && n.G <= VV_MAX_GROUPS, "n_groups out of range");
  VV_CHECK(c.window == 4, "only window_size 4 is built (nf_model/parameters0_old.yaml:29,77)");
  VV_CHECK(n.H % 16 == 0 && n.W % 64 == 0, "img_size must be a multiple of (16, 64)");
  n.h0 = n.H / 2; n.w0 = n.W / 2; n.h1 = n.H / 4; n.w1 = n.W / 4;
  n.L0 = n.h0 * n.w0; n.L1 = n.h1 * n.w1;
  n.cin = n.cout = 0;
  for (int g = 0; g < n.G; ++g) { n.cin += c.in_chans[g]; n.cout += c.out_chans[g]; }
  n.ckeep = c.keep_out > 0 ? c.keep_out : n.cout;
  VV_CHECK(n.ckeep <= n.cout, "keep_out exceeds the output channels");
  VV_CHECK(n.D % c.enc_heads[0] == 0 && n.D / c.enc_heads[0] == 32, "tower stage-0 head_dim must be 32");
  VV_CHECK(2 * n.D / c.enc_heads[1] == 32, "tower stage-1 head_dim must be 32");
  for (int l = 0; l < c.n_lg; ++l) VV_CHECK(n.E / c.lg_heads[l] == 192 && n.E % c.lg_heads[l] == 0, "trunk head_dim must be 192");
  VV_CHECK(ln_supported(MAP_PLAIN, n.D) && ln_supported(MAP_PLAIN, 2 * n.D) && ln_supported(MAP_PLAIN, n.E) &&
               ln_supported(MAP_MERGE, 4 * n.D) && ln_supported(MAP_EXPAND, n.D),
           "LayerNorm width not instantiated for enc_dim=%d embed_dim=%d", n.D, n.E);
  VV_CHECK(p2t_supported(n.D), "enc_dim %d not supported by the patch kernels (32, 64, 96, 128)", n.D);
  for (int g = 0; g < n.G; ++g) VV_CHECK(c.in_chans[g] * 4 <= 128 && c.out_chans[g] * 4 <= 128, "group has too many channels");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------
struct WeightReader {
  vv_engine* e; Net* n; bool ok = true;
  const float* dev(const std::string& name, long long numel) {
    auto it = n->staged.find(name);
    if (it == n->staged.end()) { set_error("missing weight '%s'", name.c_str()); ok = false; return nullptr; }
    long long m = 1;
    for (auto v : it->second.second) m *= v;
    if (m != numel) { set_error("weight '%s' has %lld elements, expected %lld", name.c_str(), m, numel); ok = false; return nullptr; }
    return it->second.first;
  }
  std::vector<float> host(const std::string& name, long long numel) {
    std::vector<float> h((size_t)numel, 0.f);
    const float* d = dev(name, numel);
    if (d) cudaMemcpy(h.data(), d, (size_t)numel * sizeof(float), cudaMemcpyDeviceToHost);
    return h;
  }
};

static int g_pack_f16 = 0;   // set by finalize_net from the engine configuration
static void pack_matrix(bf16* dst, bf16* dstT, const float* src, int rows, int cols) {
  if (!src) return;
  pack_w_kernel<<<592, 256>>>(dst, dstT, src, rows, cols, g_pack_f16);
}

// Gather the (2ws-1)^2 x heads table into dense [heads][16][16] (swinblock.py:88-103, 154-157).
static std::vector<float> dense_relbias(const std::vector<float>& table, int heads) {
  std::vector<float> out((size_t)heads * 256);
  for (int i = 0; i < 16; ++i)
    for (int j = 0; j < 16; ++j) {
      const int idx = ((i >> 2) - (j >> 2) + 3) * 7 + ((i & 3) - (j & 3) + 3);
      for (int h = 0; h < heads; ++h) out[((size_t)h * 16 + i) * 16 + j] = table[(size_t)idx * heads + h];
    }
  return out;
}

static int build_blocks(vv_engine* e, WeightReader& R, std::vector<BlockW>& out, const std::vector<std::string>& stage_prefix,
                        int depth, int d, int heads) {
  const int G = (int)stage_prefix.size();
  out.resize(depth);
  for (int b = 0; b < depth; ++b) {
    BlockW& w = out[b];
    w.d = d; w.heads = heads; w.G = G;
    const size_t dd = (size_t)d * d;
    w.Wqkv = dalloc<bf16>(e, G * 3 * dd); w.WqkvT = dalloc<bf16>(e, G * 3 * dd);
    w.Wproj = dalloc<bf16>(e, G * dd); w.WprojT = dalloc<bf16>(e, G * dd);
    // fc1 | fc2 (and fc2^T | fc1^T) share one allocation each: the fused tower MLP reads both, and the GEMM in front of it
    // prefetches the pair into L2 as one region
    w.W1 = dalloc<bf16>(e, G * 8 * dd); w.W2T = dalloc<bf16>(e, G * 8 * dd);
    w.W2 = w.W1 ? w.W1 + G * 4 * dd : nullptr; w.W1T = w.W2T ? w.W2T + G * 4 * dd : nullptr;
    if (!w.Wqkv || !w.WqkvT || !w.Wproj || !w.WprojT || !w.W1 || !w.W1T || !w.W2 || !w.W2T) return -1;
    std::vector<float> bqkv, bproj, b1, b2, g1, be1, g2, be2, rb;
    auto app = [](std::vector<float>& dst, const std::vector<float>& src) { dst.insert(dst.end(), src.begin(), src.end()); };
    for (int g = 0; g < G; ++g) {
      const std::string p = stage_prefix[g] + ".blocks." + std::to_string(b);
      pack_matrix(w.Wqkv + g * 3 * dd, w.WqkvT + g * 3 * dd, R.dev(p + ".attn.qkv.weight", 3 * dd), 3 * d, d);
      pack_matrix(w.Wproj + g * dd, w.WprojT + g * dd, R.dev(p + ".attn.proj.weight", dd), d, d);
      pack_matrix(w.W1 + g * 4 * dd, w.W1T + g * 4 * dd, R.dev(p + ".mlp.fc1.weight", 4 * dd), 4 * d, d);
      pack_matrix(w.W2 + g * 4 * dd, w.W2T + g * 4 * dd, R.dev(p + ".mlp.fc2.weight", 4 * dd), d, 4 * d);
      app(bqkv, R.host(p + ".attn.qkv.bias", 3 * d));
      app(bproj, R.host(p + ".attn.proj.bias", d));
      app(b1, R.host(p + ".mlp.fc1.bias", 4 * d));
      app(b2, R.host(p + ".mlp.fc2.bias", d));
      app(g1, R.host(p + ".norm1.weight", d)); app(be1, R.host(p + ".norm1.bias", d));
      app(g2, R.host(p + ".norm2.weight", d)); app(be2, R.host(p + ".norm2.bias", d));
      app(rb, dense_relbias(R.host(p + ".attn.relative_position_bias_table", 49 * heads), heads));
      if (!R.ok) return -2;
    }
    w.bqkv = dupload(e, bqkv); w.bproj = dupload(e, bproj); w.b1 = dupload(e, b1); w.b2 = dupload(e, b2);
    w.g1 = dupload(e, g1); w.be1 = dupload(e, be1); w.g2 = dupload(e, g2); w.be2 = dupload(e, be2);
    w.relbias = dupload(e, rb);
    // norm1 folded into qkv, norm2 into fc1 (forward operands only)
    w.sqkv = dalloc<float>(e, (size_t)G * 3 * d); w.cqkv = dalloc<float>(e, (size_t)G * 3 * d);
    w.s1 = dalloc<float>(e, (size_t)G * 4 * d); w.c1 = dalloc<float>(e, (size_t)G * 4 * d);
    if (!w.sqkv || !w.cqkv || !w.s1 || !w.c1) return -1;
    for (int g = 0; g < G && !e->cfg.no_ln_fold; ++g) {
      const std::string p = stage_prefix[g] + ".blocks." + std::to_string(b);
      const float* wq = R.dev(p + ".attn.qkv.weight", 3 * dd);
      const float* w1 = R.dev(p + ".mlp.fc1.weight", 4 * dd);
      if (!R.ok) return -2;
      fold_ln_kernel<<<(3 * d + 7) / 8, 256>>>(w.Wqkv + g * 3 * dd, w.sqkv + (size_t)g * 3 * d, w.cqkv + (size_t)g * 3 * d, wq, w.g1 + (size_t)g * d,
                                                w.be1 + (size_t)g * d, w.bqkv + (size_t)g * 3 * d, 3 * d, d, g_pack_f16);
      fold_ln_kernel<<<(4 * d + 7) / 8, 256>>>(w.W1 + g * 4 * dd, w.s1 + (size_t)g * 4 * d, w.c1 + (size_t)g * 4 * d, w1, w.g2 + (size_t)g * d,
                                                w.be2 + (size_t)g * d, w.b1 + (size_t)g * 4 * d, 4 * d, d, g_pack_f16);
    }
  }
  return 0;
}

static int finalize_net(vv_engine* e, Net& n) {
  WeightReader R{e, &n};
  g_pack_f16 = e->cfg.forward_fp16 ? 1 : 0;
  const int G = n.G, D = n.D, E = n.E;
  std::vector<std::string> pe0, pe1, pu0, pu1;
  for (int g = 0; g < G; ++g) {
    pe0.push_back("enc.enc_list." + std::to_string(g) + ".layers.0");
    pe1.push_back("enc.enc_list." + std::to_string(g) + ".layers.1");
    pu0.push_back("dec.dec_list." + std::to_string(g) + ".layers_up.0");
    pu1.push_back("dec.dec_list." + std::to_string(g) + ".layers_up.1");
  }
  int rc;
  if ((rc = build_blocks(e, R, n.e0, pe0, n.c.enc_depth[0], D, n.c.enc_heads[0]))) return rc;
  if ((rc = build_blocks(e, R, n.e1, pe1, n.c.enc_depth[1], 2 * D, n.c.enc_heads[1]))) return rc;
  if ((rc = build_blocks(e, R, n.u0, pu0, n.c.enc_depth[1], 2 * D, n.c.enc_heads[1]))) return rc;
  if ((rc = build_blocks(e, R, n.u1, pu1, n.c.enc_depth[0], D, n.c.enc_heads[0]))) return rc;
  n.lg.clear();
  for (int l = 0; l < n.c.n_lg; ++l) {
    std::vector<BlockW> tmp;
    if ((rc = build_blocks(e, R, tmp, {"net.layers." + std::to_string(l)}, n.c.lg_depth[l], E, n.c.lg_heads[l]))) return rc;
    n.lg.insert(n.lg.end(), tmp.begin(), tmp.end());
  }

  // ---- patch embed (Conv2d k2 s2, transformer.py:35) and APE (:351, :394) ----
  {
    std::vector<int> kcnt, cbase, chan;
    std::vector<float> Wp, bias;
    int c0 = 0;
    n.ape = dalloc<float>(e, (size_t)G * n.L0 * D);
    for (int g = 0; g < G; ++g) {
      const int cg = n.c.in_chans[g];
      const std::string p = "enc.enc_list." + std::to_string(g);
      std::vector<float> w = R.host(p + ".patch_embed.proj.weight", (long long)D * cg * 4);
      std::vector<float> b = R.host(p + ".patch_embed.proj.bias", D);
      kcnt.push_back(cg); cbase.push_back(c0);
      for (int ci = 0; ci < cg; ++ci) chan.push_back(c0 + ci);
      for (int ci = 0; ci < cg; ++ci)
        for (int pp = 0; pp < 4; ++pp)
          for (int c = 0; c < D; ++c) Wp.push_back(w[((size_t)c * cg + ci) * 4 + pp]);
      bias.insert(bias.end(), b.begin(), b.end());
      const float* ape = R.dev(p + ".absolute_pos_embed", (long long)n.L0 * D);
      if (ape) cudaMemcpy(n.ape + (size_t)g * n.L0 * D, ape, (size_t)n.L0 * D * sizeof(float), cudaMemcpyDeviceToDevice);
      c0 += cg;
    }
    n.embed.kcnt = dupload(e, kcnt); n.embed.cbase = dupload(e, cbase); n.embed.chan = dupload(e, chan);
    n.embed.Wp = dupload(e, Wp); n.embed.bias = dupload(e, bias); n.embed.nslots = c0;
    n.embed.max_cnt = *std::max_element(kcnt.begin(), kcnt.end());
  }
  // ---- final projection (ConvTranspose2d k2 s2, transformer.py:593-594) with the mean/std half shuffle (:616-623) ----
  {
    std::vector<int> kcnt, cbase, chan;
    std::vector<float> Wp, bias;
    int mean_total = 0;
    for (int g = 0; g < G; ++g) mean_total += n.c.out_chans[g] / 2;
    int mean_off = 0, std_off = 0, slots = 0;
    for (int g = 0; g < G; ++g) {
      const int cg = n.c.out_chans[g], half = cg / 2;
      std::vector<float> w = R.host("dec.final_proj_list." + std::to_string(g) + ".weight", (long long)D * cg * 4);
      std::vector<float> b = R.host("dec.final_proj_list." + std::to_string(g) + ".bias", cg);
      cbase.push_back(slots);
      int cnt = 0;
      for (int k = 0; k < cg; ++k) {
        const int oc = k < half ? mean_off + k : mean_total + std_off + (k - half);
        if (oc >= n.ckeep) continue;
        chan.push_back(oc);
        for (int pp = 0; pp < 4; ++pp)
          for (int c = 0; c < D; ++c) Wp.push_back(w[((size_t)c * cg + k) * 4 + pp]);
        bias.push_back(b[k]);
        ++cnt; ++slots;
      }
      kcnt.push_back(cnt);
      mean_off += half; std_off += cg - half;
    }
    VV_CHECK(slots == n.ckeep, "keep_out=%d does not align with the mean/std channel layout (%d slots)", n.ckeep, slots);
    n.fin.kcnt = dupload(e, kcnt); n.fin.cbase = dupload(e, cbase); n.fin.chan = dupload(e, chan);
    n.fin.Wp = dupload(e, Wp); n.fin.bias = dupload(e, bias); n.fin.nslots = slots;
    n.fin.max_cnt = *std::max_element(kcnt.begin(), kcnt.end());
  }
  // ---- stacked per-group seams ----
  auto stack_vec = [&](const char* fmt_suffix, const std::vector<std::string>& prefixes, long long numel) {
    std::vector<float> all;
    for (auto& p : prefixes) {
      std::vector<float> v = R.host(p + fmt_suffix, numel);
      all.insert(all.end(), v.begin(), v.end());
    }
    return dupload(e, all);
  };
  std::vector<std::string> penc, pdec;
  for (int g = 0; g < G; ++g) {
    penc.push_back("enc.enc_list." + std::to_string(g));
    pdec.push_back("dec.dec_list." + std::to_string(g));
  }
  n.mg_g = stack_vec(".layers.1.downsample.norm.weight", penc, 4 * D);
  n.mg_b = stack_vec(".layers.1.downsample.norm.bias", penc, 4 * D);
  n.en_g = stack_vec(".norm.weight", penc, 2 * D);
  n.en_b = stack_vec(".norm.bias", penc, 2 * D);
  n.ex_g = stack_vec(".layers_up.0.upsample.norm.weight", pdec, D);
  n.ex_b = stack_vec(".layers_up.0.upsample.norm.bias", pdec, D);
  n.nu_g = stack_vec(".norm_up.weight", pdec, D);
  n.nu_b = stack_vec(".norm_up.bias", pdec, D);
  n.bc0 = stack_vec(".concat_back_dim.0.bias", pdec, 2 * D);
  n.bc1 = stack_vec(".concat_back_dim.1.bias", pdec, D);
  const size_t DD = (size_t)D * D;
  n.Wred = dalloc<bf16>(e, G * 8 * DD); n.WredT = dalloc<bf16>(e, G * 8 * DD);
  n.Wc0 = dalloc<bf16>(e, G * 8 * DD); n.Wc0T = dalloc<bf16>(e, G * 8 * DD);
  n.Wex = dalloc<bf16>(e, G * 8 * DD); n.WexT = dalloc<bf16>(e, G * 8 * DD);
  n.Wc1 = dalloc<bf16>(e, G * 2 * DD); n.Wc1T = dalloc<bf16>(e, G * 2 * DD);
  for (int g = 0; g < G; ++g) {
    pack_matrix(n.Wred + g * 8 * DD, n.WredT + g * 8 * DD, R.dev(penc[g] + ".layers.1.downsample.reduction.weight", 8 * DD), 2 * D, 4 * D);
    pack_matrix(n.Wc0 + g * 8 * DD, n.Wc0T + g * 8 * DD, R.dev(pdec[g] + ".concat_back_dim.0.weight", 8 * DD), 2 * D, 4 * D);
    pack_matrix(n.Wex + g * 8 * DD, n.WexT + g * 8 * DD, R.dev(pdec[g] + ".layers_up.0.upsample.expand.weight", 8 * DD), 4 * D, 2 * D);
    pack_matrix(n.Wc1 + g * 2 * DD, n.Wc1T + g * 2 * DD, R.dev(pdec[g] + ".concat_back_dim.1.weight", 2 * DD), D, 2 * D);
  }
  const size_t EP = (size_t)E * G * 2 * D;
  n.Wep = dalloc<bf16>(e, EP); n.WepT = dalloc<bf16>(e, EP);
  n.Wdp = dalloc<bf16>(e, EP); n.WdpT = dalloc<bf16>(e, EP);
  pack_matrix(n.Wep, n.WepT, R.dev("enc.proj.weight", EP), E, G * 2 * D);
  pack_matrix(n.Wdp, n.WdpT, R.dev("dec.proj.weight", EP), G * 2 * D, E);
  n.bep = dupload(e, R.host("enc.proj.bias", E));
  n.bdp = dupload(e, R.host("dec.proj.bias", G * 2 * D));
  n.pos = dalloc<float>(e, (size_t)n.L1 * E);
  if (const float* p = R.dev("net.pos_embed", (long long)n.L1 * E))
    cudaMemcpy(n.pos, p, (size_t)n.L1 * E * sizeof(float), cudaMemcpyDeviceToDevice);
  if (!R.ok) return -2;
  VV_CUDA(cudaDeviceSynchronize());
  for (auto& kv : n.staged) cudaFree(kv.second.first);
  n.staged.clear();
  n.finalized = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// plan construction
// ---------------------------------------------------------------------------------------------
struct Temps {
  bf16 *h, *ao, *a, *du, *dao, *dqkv, *dx1b, *dhb;
  float *dh, *dx1;
  float* lnst; size_t lnst_cap;      // LayerNorm statistics partials: float2 [batch][parts][rows]; capacity in float2
  float* lnshift;                    // per-row shift of the centred 16-bit copy: [batch][rows]
  // seams (forward)
  bf16 *MB, *EPIN, *TB, *CAT0, *CAT1, *U1B;
  float* NU;
  // gradients per stage (fp32 + bf16), seams (backward)
  float *gU1, *gU0, *gT, *gE1, *gE0; bf16 *gU1b, *gU0b, *gTb, *gE1b, *gE0b;
  float *dNU, *dC1a, *dSK0, *dEX, *dSK1, *dEP, *dMB; bf16 *dEXb, *DPb;
};

struct Builder {
  vv_engine* e; Net* n; Temps t; const char* err = nullptr;
  int f16 = 0;        // forward activations / weights in fp16 (gradients always bf16)
  bool fold = true;   // norm1 / norm2 folded into the qkv / fc1 GEMMs (vv_config::no_ln_fold)
  bool fuse_mlp = true;   // tower blocks: norm2 + fc1 + GELU + fc2 + residual (and its input-VJP) as one kernel (VV_NO_FUSED_MLP=1: off)
  bool mlp_fused(const BlockW& w, int rows) const { return fuse_mlp && fold && mlp_fused_supported(w.d, rows); }
  void mlp(Plan& P, bool bwd, int D, const bf16* W1, const bf16* W2, const bf16* u, const bf16* dy16, const MlpArgs& a) {
    Op o{}; o.kind = bwd ? Op::MLP_B : Op::MLP_F;
    const char* er = make_mlp_desc(&o.mlp, D, bwd, W1, W2, u, dy16, a);
    if (er && !err) err = er;
    P.ops.push_back(o);
  }

  void gemm(Plan& P, const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb, long long b_bs, GemmArgs g) {
    Op o{}; o.kind = Op::GEMM;
    const char* er = make_gemm_desc(&o.gemm, A, lda, a_bs, B, ldb, b_bs, g);
    if (er && !err) err = er;
    P.ops.push_back(o);
  }
  // forward GEMM (operands in the forward format) / gradient GEMM (bf16 operands; the saved pre-activation keeps the forward format)
  GemmArgs ga(int M, int N, int K, int batch) const {
    GemmArgs g{}; g.M = M; g.N = N; g.K = K; g.batch = batch; g.epi = EPI_LINEAR; g.f16 = f16; g.aux_f16 = f16; return g;
  }
  GemmArgs gb(int M, int N, int K, int batch) const {
    GemmArgs g = ga(M, N, K, batch); g.f16 = 0; return g;
  }
  void ln_f(Plan& P, int rows, int C, int batch, int map, int gh, int gw, float eps, const float* x, long long ld_x, long long x_bs,
            const float* gamma, const float* beta, bf16* ob, long long ld_ob, long long ob_bs, float* of, long long ld_of, long long of_bs) {
    Op o{}; o.kind = Op::LN_F;
    o.lnf = LnArgs{rows, C, batch, map, gh, gw, eps, x, ld_x, x_bs, gamma, beta, (long long)C, ob, ld_ob, ob_bs, of, ld_of, of_bs, f16, nullptr, nullptr};
    P.ops.push_back(o);
  }
  // statistics-only LayerNorm pass (raw 16-bit copy of x into t.h + (sum, sumsq) per row) for a stage input no GEMM produced
  void ln_stats(Plan& P, int rows, int C, int batch, const float* x) {
    const long long rd = (long long)rows * C;
    Op o{}; o.kind = Op::LN_F;
    o.lnf = LnArgs{rows, C, batch, MAP_PLAIN, 0, 0, 0.f, x, (long long)C, rd, nullptr, nullptr, 0, t.h, (long long)C, rd, nullptr, 0, 0, f16, t.lnst, t.lnshift};
    P.ops.push_back(o);
  }
  void ln_b(Plan& P, int rows, int C, int batch, int map, int gh, int gw, float eps, const float* x, long long ld_x, long long x_bs,
            const float* gamma, const float* dy, long long ld_dy, long long dy_bs, const float* dres, long long ld_dres, long long dres_bs,
            float* dx, long long ld_dx, long long dx_bs, bf16* dxb, long long ld_dxb, long long dxb_bs, const bf16* dy16 = nullptr) {
    Op o{}; o.kind = Op::LN_B;
    o.lnb = LnBwdArgs{rows, C, batch, map, gh, gw, eps, x, ld_x, x_bs, gamma, (long long)C, dy, ld_dy, dy_bs, dres, ld_dres, dres_bs,
                      dx, ld_dx, dx_bs, dxb, ld_dxb, dxb_bs, dy16};
    P.ops.push_back(o);
  }
  // GEMM whose fp32 output is the input of a LayerNorm folded into the NEXT GEMM: it also writes the raw 16-bit copy of its
  // output rows (width C), centred on the row's stage-input mean (t.lnshift), into t.h and their (mean, M2) partials into t.lnst.
  // Returns the number of partials per row; prod_bn receives the producer's tile width (the consumer needs it to weigh them).
  int gemm_with_stats(Plan& P, const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb, long long b_bs, GemmArgs g, int C,
                      int& prod_bn) {
    g.out_bf16 = t.h; g.ld_bf16 = C; g.bf16_bs = (long long)g.M * C;
    g.stats_out = t.lnst;
    g.ln_shift = t.lnshift;
    gemm(P, A, lda, a_bs, B, ldb, b_bs, g);
    GemmDesc& d = P.ops.back().gemm;
    const int parts = 2 * ((g.N + d.bn - 1) / d.bn);
    d.a.stats_out_bs = (long long)parts * g.M * 2;
    if ((size_t)parts * g.M * g.batch > t.lnst_cap && !err) err = "LayerNorm statistics buffer too small";
    prod_bn = d.bn;
    return parts;
  }
  void fold_ln(GemmArgs& g, int parts, int prod_bn, const float* colsum, int C, float eps) const {
    g.ln_stats = t.lnst; g.ln_parts = parts; g.ln_stats_bs = (long long)parts * g.M * 2; g.ln_colsum = colsum;
    g.ln_prod_bn = prod_bn; g.ln_c = C; g.ln_shift = t.lnshift; g.ln_health = e->ln_health;
    g.ln_inv_c = 1.0f / (float)C; g.ln_eps = eps;
  }

  // SwinTransformerBlock.forward, swinblock.py:265-309
  // norm1 / norm2 are folded into the qkv / fc1 GEMMs (GemmArgs::ln_stats): no LayerNorm launches inside a stage.
  // parts: in = number of statistics partials already in t.lnst for `x` (0: none -- a statistics pass is issued, which also fixes
  //        the rows' shifts for the whole stage); out = partials for x_out if emit_next (the next block's norm1), else 0.
  // prod_bn: tile width of the GEMM that produced those partials (0: the statistics kernel).
  // fold == false: plain LayerNorm kernels (two-pass fp32 statistics) in front of plain GEMMs.
  void block_fwd(Plan& P, const BlockW& w, int gh, int gw, int shift, const float* x, float* x_out, BlkStash& st,
                 bf16* copy_b, long long ld_c, long long bs_c, int& parts, int& prod_bn, bool emit_next) {
    const int G = w.G, d = w.d, rows = gh * gw;
    const long long rd = (long long)rows * d;
    const bool fused = mlp_fused(w, rows);      // tower block: norm1 + qkv and the whole MLP half on mlp_fused_kernel
    GemmArgs g = ga(rows, 3 * d, d, G);
    g.out_bf16 = st.qkv; g.ld_bf16 = 3 * d; g.bf16_bs = 3 * rd; g.bias_bs = 3 * d;
    if (fused) {
      MlpArgs q{};
      q.rows = rows; q.batch = G; q.f16 = f16; q.eps = 1e-5f; q.x1 = x; q.b1 = w.cqkv; q.n_out = 3 * d;
      Op o{}; o.kind = Op::LIN_F;
      const char* er = make_lin_desc(&o.mlp, d, w.Wqkv, st.qkv, 3LL * d, 3 * rd, q);
      if (er && !err) err = er;
      P.ops.push_back(o);
    } else if (fold) {
      if (parts == 0) { ln_stats(P, rows, d, G, x); parts = 1; prod_bn = 0; }
      g.bias = w.cqkv;
      fold_ln(g, parts, prod_bn, w.sqkv, d, 1e-5f);
    } else {
      ln_f(P, rows, d, G, MAP_PLAIN, gh, gw, 1e-5f, x, d, rd, w.g1, w.be1, t.h, d, rd, nullptr, 0, 0);
      g.bias = w.bqkv;
    }
    if (!fused) gemm(P, t.h, d, rd, w.Wqkv, d, 3LL * d * d, g);
    Op o{}; o.kind = Op::ATT_F;
    o.att = AttnArgs{gh, gw, w.heads, d / w.heads, shift, G, st.qkv, 3LL * d, 3 * rd, w.relbias, (long long)w.heads * 256, t.ao, d, rd, nullptr, nullptr, f16};
    P.ops.push_back(o);
    g = ga(rows, d, d, G);
    g.bias = w.bproj; g.bias_bs = d; g.res = x; g.ld_res = d; g.res_bs = rd; g.out_f32 = st.x1; g.ld_f32 = d; g.f32_bs = rd;
    if (fused) {
      // the MLP half as one kernel: norm2 is computed from x1 inside it, so proj only has to leave the fp32 residual stream; the next
      // block's norm1 is computed inside its qkv kernel, so no statistics / centred copy of the output are needed either
      gemm(P, t.ao, d, rd, w.Wproj, d, (long long)d * d, g);
      MlpArgs m{};
      m.rows = rows; m.batch = G; m.f16 = f16; m.eps = 1e-5f; m.x1 = st.x1; m.b1 = w.c1; m.b2 = w.b2; m.out_f32 = x_out; m.u_out = st.u;
      if (copy_b && !emit_next) { m.out16 = copy_b; m.ld16 = ld_c; m.bs16 = bs_c; }
      parts = 0;
      mlp(P, false, d, w.W1, w.W2, st.u, nullptr, m);
      return;
    }
    int parts2 = 0, bn2 = 0;
    if (fold) parts2 = gemm_with_stats(P, t.ao, d, rd, w.Wproj, d, (long long)d * d, g, d, bn2);
    else gemm(P, t.ao, d, rd, w.Wproj, d, (long long)d * d, g);
    g = ga(rows, 4 * d, d, G);
    g.epi = EPI_GELU; g.bias_bs = 4 * d; g.aux_out = st.u; g.ld_aux = 4 * d; g.aux_bs = 4 * rd;
    g.out_bf16 = t.a; g.ld_bf16 = 4 * d; g.bf16_bs = 4 * rd;
    if (fold) {
      g.bias = w.c1;
      fold_ln(g, parts2, bn2, w.s1, d, 1e-5f);
    } else {
      ln_f(P, rows, d, G, MAP_PLAIN, gh, gw, 1e-5f, st.x1, d, rd, w.g2, w.be2, t.h, d, rd, nullptr, 0, 0);
      g.bias = w.b1;
    }
    gemm(P, t.h, d, rd, w.W1, d, 4LL * d * d, g);
    g = ga(rows, d, 4 * d, G);
    g.bias = w.b2; g.bias_bs = d; g.res = st.x1; g.ld_res = d; g.res_bs = rd; g.out_f32 = x_out; g.ld_f32 = d; g.f32_bs = rd;
    if (emit_next && fold) {
      parts = gemm_with_stats(P, t.a, 4 * d, 4 * rd, w.W2, 4 * d, 4LL * d * d, g, d, prod_bn);
    } else {
      if (copy_b && !emit_next) { g.out_bf16 = copy_b; g.ld_bf16 = ld_c; g.bf16_bs = bs_c; }
      gemm(P, t.a, 4 * d, 4 * rd, w.W2, 4 * d, 4LL * d * d, g);
      parts = 0;
    }
  }
  // Input-VJP of the block; the gradient (g32 fp32 + g16 bf16, [G][rows][d]) is updated in place.
  void block_bwd(Plan& P, const BlockW& w, int gh, int gw, int shift, const float* x, BlkStash& st, float* g32, bf16* g16) {
    const int G = w.G, d = w.d, rows = gh * gw;
    const long long rd = (long long)rows * d;
    GemmArgs g{};
    if (mlp_fused(w, rows)) {                                 // dx1 = LN2^T((dy W2 . gelu'(u)) W1) + dy in one kernel
      MlpArgs m{};
      m.rows = rows; m.batch = G; m.f16 = f16; m.eps = 1e-5f; m.x1 = st.x1; m.gamma = w.g2; m.dres = g32; m.dx = t.dx1; m.dx16 = t.dx1b;
      mlp(P, true, d, w.W2T, w.W1T, st.u, g16, m);
    } else {
      g = gb(rows, 4 * d, d, G);                              // d(gelu out) = dy W2 ; du = . * gelu'(u)
      g.epi = EPI_DGELU; g.aux_in = st.u; g.ld_aux = 4 * d; g.aux_bs = 4 * rd; g.out_bf16 = t.du; g.ld_bf16 = 4 * d; g.bf16_bs = 4 * rd;
      gemm(P, g16, d, rd, w.W2T, d, 4LL * d * d, g);
      g = gb(rows, d, 4 * d, G);                              // d(LN2 out) = du W1
      g.out_bf16 = t.dhb; g.ld_bf16 = d; g.bf16_bs = rd;      // bf16 like every other gradient that enters a GEMM / LayerNorm adjoint
      gemm(P, t.du, 4 * d, 4 * rd, w.W1T, 4 * d, 4LL * d * d, g);
      ln_b(P, rows, d, G, MAP_PLAIN, gh, gw, 1e-5f, st.x1, d, rd, w.g2, nullptr, d, rd, g32, d, rd, t.dx1, d, rd, t.dx1b, d, rd, t.dhb);
    }
    g = gb(rows, d, d, G);                                    // d(attn out) = dx1 Wproj
    g.out_bf16 = t.dao; g.ld_bf16 = d; g.bf16_bs = rd;
    gemm(P, t.dx1b, d, rd, w.WprojT, d, (long long)d * d, g);
    Op o{}; o.kind = Op::ATT_B;
    o.att = AttnArgs{gh, gw, w.heads, d / w.heads, shift, G, st.qkv, 3LL * d, 3 * rd, w.relbias, (long long)w.heads * 256, nullptr, d, rd, t.dao, t.dqkv, f16};
    P.ops.push_back(o);
    g = gb(rows, d, 3 * d, G);                                // d(LN1 out) = dqkv Wqkv
    g.out_bf16 = t.dhb; g.ld_bf16 = d; g.bf16_bs = rd;
    gemm(P, t.dqkv, 3 * d, 3 * rd, w.WqkvT, 3 * d, 3LL * d * d, g);
    ln_b(P, rows, d, G, MAP_PLAIN, gh, gw, 1e-5f, x, d, rd, w.g1, nullptr, d, rd, t.dx1, d, rd, g32, d, rd, g16, d, rd, t.dhb);
  }
  void stage_fwd(Plan& P, const std::vector<BlockW>& ws, int gh, int gw, StageStash& st, bf16* copy_b, long long ld_c, long long bs_c) {
    int parts = 0, prod_bn = 0;
    for (size_t b = 0; b < ws.size(); ++b) {
      const bool last = b + 1 == ws.size();
      block_fwd(P, ws[b], gh, gw, (b % 2) ? 2 : 0, st.x[b], st.x[b + 1], st.b[b], last ? copy_b : nullptr, ld_c, bs_c, parts, prod_bn, !last);
    }
  }
  void stage_bwd(Plan& P, const std::vector<BlockW>& ws, int gh, int gw, StageStash& st, float* g32, bf16* g16) {
    for (int b = (int)ws.size() - 1; b >= 0; --b) block_bwd(P, ws[b], gh, gw, (b % 2) ? 2 : 0, st.x[b], st.b[b], g32, g16);
  }

  // LGUnet_all.forward, transformer.py:747-752 (Enc_net :554-568, LG_net :698-712, Dec_net :599-625)
  void net_fwd(Plan& P, Stash& S, const float* in, float* out) {
    Net& N = *n;
    const int G = N.G, D = N.D, E = N.E, L0 = N.L0, L1 = N.L1;
    Op o{}; o.kind = Op::P2T;
    o.patch = PatchArgs{N.H, N.W, G, D, N.embed.kcnt, N.embed.cbase, N.embed.chan, N.embed.Wp, N.embed.bias, N.ape, in, S.e0.x[0], nullptr, nullptr, N.embed.max_cnt};
    P.ops.push_back(o);
    stage_fwd(P, N.e0, N.h0, N.w0, S.e0, t.CAT1 + D, 2 * D, (long long)L0 * 2 * D);            // skip 0 -> CAT1[:, D:2D]
    float* S0 = S.e0.x.back();
    ln_f(P, L1, 4 * D, G, MAP_MERGE, N.h0, N.w0, 1e-6f, S0, D, (long long)L0 * D, N.mg_g, N.mg_b, t.MB, 4 * D, (long long)L1 * 4 * D, nullptr, 0, 0);
    GemmArgs g = ga(L1, 2 * D, 4 * D, G);
    g.out_f32 = S.e1.x[0]; g.ld_f32 = 2 * D; g.f32_bs = (long long)L1 * 2 * D;
    gemm(P, t.MB, 4 * D, (long long)L1 * 4 * D, N.Wred, 4 * D, 8LL * D * D, g);
    stage_fwd(P, N.e1, N.h1, N.w1, S.e1, t.CAT0 + 2 * D, 4 * D, (long long)L1 * 4 * D);        // skip 1 -> CAT0[:, 2D:4D]
    float* S1 = S.e1.x.back();
    ln_f(P, L1, 2 * D, G, MAP_PLAIN, N.h1, N.w1, 1e-6f, S1, 2 * D, (long long)L1 * 2 * D, N.en_g, N.en_b, t.EPIN, (long long)G * 2 * D, 2 * D, nullptr, 0, 0);
    g = ga(L1, E, G * 2 * D, 1);
    g.bias = N.bep; g.res = N.pos; g.ld_res = E; g.out_f32 = S.lg.x[0]; g.ld_f32 = E;          // + pos_embed (:704)
    gemm(P, t.EPIN, (long long)G * 2 * D, 0, N.Wep, (long long)G * 2 * D, 0, g);
    stage_fwd_trunk(P, S);
    g = ga(L1, G * 2 * D, E, 1);                                                                // Dec_net.proj, split over towers (:600-601)
    g.bias = N.bdp; g.out_bf16 = t.CAT0; g.ld_bf16 = 4 * D; g.split_n = 2 * D; g.split_stride = (long long)L1 * 4 * D;
    gemm(P, t.TB, E, 0, N.Wdp, E, 0, g);
    g = ga(L1, 2 * D, 4 * D, G);                                                                // concat_back_dim[0] (:468-469)
    g.bias = N.bc0; g.bias_bs = 2 * D; g.out_f32 = S.u0.x[0]; g.ld_f32 = 2 * D; g.f32_bs = (long long)L1 * 2 * D;
    gemm(P, t.CAT0, 4 * D, (long long)L1 * 4 * D, N.Wc0, 4 * D, 8LL * D * D, g);
    stage_fwd(P, N.u0, N.h1, N.w1, S.u0, t.U1B, 2 * D, (long long)L1 * 2 * D);
    g = ga(L1, 4 * D, 2 * D, G);                                                                // PatchExpand.expand (:110)
    g.out_f32 = S.EX; g.ld_f32 = 4 * D; g.f32_bs = (long long)L1 * 4 * D;
    gemm(P, t.U1B, 2 * D, (long long)L1 * 2 * D, N.Wex, 2 * D, 8LL * D * D, g);
    ln_f(P, L0, D, G, MAP_EXPAND, N.h0, N.w0, 1e-6f, S.EX, 4 * D, (long long)L1 * 4 * D, N.ex_g, N.ex_b, t.CAT1, 2 * D, (long long)L0 * 2 * D, nullptr, 0, 0);
    g = ga(L0, D, 2 * D, G);                                                                    // concat_back_dim[1]
    g.bias = N.bc1; g.bias_bs = D; g.out_f32 = S.u1.x[0]; g.ld_f32 = D; g.f32_bs = (long long)L0 * D;
    gemm(P, t.CAT1, 2 * D, (long long)L0 * 2 * D, N.Wc1, 2 * D, 2LL * D * D, g);
    stage_fwd(P, N.u1, N.h0, N.w0, S.u1, nullptr, 0, 0);
    ln_f(P, L0, D, G, MAP_PLAIN, N.h0, N.w0, 1e-6f, S.u1.x.back(), D, (long long)L0 * D, N.nu_g, N.nu_b, nullptr, 0, 0, t.NU, D, (long long)L0 * D);
    o = Op{}; o.kind = Op::T2P;
    o.patch = PatchArgs{N.H, N.W, G, D, N.fin.kcnt, N.fin.cbase, N.fin.chan, N.fin.Wp, N.fin.bias, nullptr, nullptr, nullptr, t.NU, out, N.fin.max_cnt};
    P.ops.push_back(o);
  }
  void stage_fwd_trunk(Plan& P, Stash& S) {
    Net& N = *n;
    int parts = 0, prod_bn = 0;
    for (size_t b = 0; b < N.lg.size(); ++b) {
      const bool last = b + 1 == N.lg.size();
      block_fwd(P, N.lg[b], N.h1, N.w1, shift_of_trunk(b), S.lg.x[b], S.lg.x[b + 1], S.lg.b[b], last ? t.TB : nullptr, N.E, 0, parts, prod_bn, !last);
    }
  }
  int shift_of_trunk(size_t b) const {          // block index inside its Layer decides the shift (transformer.py:502)
    size_t k = b;
    for (int l = 0; l < n->c.n_lg; ++l) {
      if (k < (size_t)n->c.lg_depth[l]) return (k % 2) ? 2 : 0;
      k -= n->c.lg_depth[l];
    }
    return 0;
  }

  // Hand-derived input-VJP of the whole application.  dout: (ckeep,H,W), din: (cin,H,W).
  void net_bwd(Plan& P, Stash& S, const float* dout, float* din) {
    Net& N = *n;
    const int G = N.G, D = N.D, E = N.E, L0 = N.L0, L1 = N.L1;
    Op o{}; o.kind = Op::P2T;                                                                     // ConvTranspose2d^T
    o.patch = PatchArgs{N.H, N.W, G, D, N.fin.kcnt, N.fin.cbase, N.fin.chan, N.fin.Wp, nullptr, nullptr, dout, t.dNU, nullptr, nullptr, N.fin.max_cnt};
    P.ops.push_back(o);
    ln_b(P, L0, D, G, MAP_PLAIN, N.h0, N.w0, 1e-6f, S.u1.x.back(), D, (long long)L0 * D, N.nu_g, t.dNU, D, (long long)L0 * D, nullptr, 0, 0,
         t.gU1, D, (long long)L0 * D, t.gU1b, D, (long long)L0 * D);
    stage_bwd(P, N.u1, N.h0, N.w0, S.u1, t.gU1, t.gU1b);
    // concat_back_dim[1]^T: first D input columns -> PatchExpand norm output, last D -> skip 0
    GemmArgs g = gb(L0, D, D, G);
    g.out_f32 = t.dC1a; g.ld_f32 = D; g.f32_bs = (long long)L0 * D;
    gemm(P, t.gU1b, D, (long long)L0 * D, N.Wc1T, D, 2LL * D * D, g);
    g = gb(L0, D, D, G);
    g.out_f32 = t.dSK0; g.ld_f32 = D; g.f32_bs = (long long)L0 * D;
    gemm(P, t.gU1b, D, (long long)L0 * D, N.Wc1T + (long long)D * D, D, 2LL * D * D, g);
    ln_b(P, L0, D, G, MAP_EXPAND, N.h0, N.w0, 1e-6f, S.EX, 4 * D, (long long)L1 * 4 * D, N.ex_g, t.dC1a, D, (long long)L0 * D, nullptr, 0, 0,
         t.dEX, 4 * D, (long long)L1 * 4 * D, t.dEXb, 4 * D, (long long)L1 * 4 * D);
    g = gb(L1, 2 * D, 4 * D, G);                                                                  // expand^T
    g.out_f32 = t.gU0; g.ld_f32 = 2 * D; g.f32_bs = (long long)L1 * 2 * D; g.out_bf16 = t.gU0b; g.ld_bf16 = 2 * D; g.bf16_bs = (long long)L1 * 2 * D;
    gemm(P, t.dEXb, 4 * D, (long long)L1 * 4 * D, N.WexT, 4 * D, 8LL * D * D, g);
    stage_bwd(P, N.u0, N.h1, N.w1, S.u0, t.gU0, t.gU0b);
    // concat_back_dim[0]^T: first 2D columns -> Dec_net.proj output slice of tower g, last 2D -> skip 1
    g = gb(L1, 2 * D, 2 * D, G);
    g.out_bf16 = t.DPb; g.ld_bf16 = (long long)G * 2 * D; g.bf16_bs = 2 * D;
    gemm(P, t.gU0b, 2 * D, (long long)L1 * 2 * D, N.Wc0T, 2 * D, 8LL * D * D, g);
    g = gb(L1, 2 * D, 2 * D, G);
    g.out_f32 = t.dSK1; g.ld_f32 = 2 * D; g.f32_bs = (long long)L1 * 2 * D;
    gemm(P, t.gU0b, 2 * D, (long long)L1 * 2 * D, N.Wc0T + 4LL * D * D, 2 * D, 8LL * D * D, g);
    g = gb(L1, E, G * 2 * D, 1);                                                                  // Dec_net.proj^T
    g.out_f32 = t.gT; g.ld_f32 = E; g.out_bf16 = t.gTb; g.ld_bf16 = E;
    gemm(P, t.DPb, (long long)G * 2 * D, 0, N.WdpT, (long long)G * 2 * D, 0, g);
    for (int b = (int)N.lg.size() - 1; b >= 0; --b) block_bwd(P, N.lg[b], N.h1, N.w1, shift_of_trunk(b), S.lg.x[b], S.lg.b[b], t.gT, t.gTb);
    g = gb(L1, G * 2 * D, E, 1);                                                                  // Enc_net.proj^T
    g.out_f32 = t.dEP; g.ld_f32 = (long long)G * 2 * D;
    gemm(P, t.gTb, E, 0, N.WepT, E, 0, g);
    ln_b(P, L1, 2 * D, G, MAP_PLAIN, N.h1, N.w1, 1e-6f, S.e1.x.back(), 2 * D, (long long)L1 * 2 * D, N.en_g, t.dEP, (long long)G * 2 * D, 2 * D,
         t.dSK1, 2 * D, (long long)L1 * 2 * D, t.gE1, 2 * D, (long long)L1 * 2 * D, t.gE1b, 2 * D, (long long)L1 * 2 * D);
    stage_bwd(P, N.e1, N.h1, N.w1, S.e1, t.gE1, t.gE1b);
    g = gb(L1, 4 * D, 2 * D, G);                                                                  // reduction^T
    g.out_f32 = t.dMB; g.ld_f32 = 4 * D; g.f32_bs = (long long)L1 * 4 * D;
    gemm(P, t.gE1b, 2 * D, (long long)L1 * 2 * D, N.WredT, 2 * D, 8LL * D * D, g);
    ln_b(P, L1, 4 * D, G, MAP_MERGE, N.h0, N.w0, 1e-6f, S.e0.x.back(), D, (long long)L0 * D, N.mg_g, t.dMB, 4 * D, (long long)L1 * 4 * D,
         t.dSK0, D, (long long)L0 * D, t.gE0, D, (long long)L0 * D, t.gE0b, D, (long long)L0 * D);
    stage_bwd(P, N.e0, N.h0, N.w0, S.e0, t.gE0, t.gE0b);
    o = Op{}; o.kind = Op::T2P;                                                                   // Conv2d^T
    o.patch = PatchArgs{N.H, N.W, G, D, N.embed.kcnt, N.embed.cbase, N.embed.chan, N.embed.Wp, nullptr, nullptr, nullptr, nullptr, t.gE0, din, N.embed.max_cnt};
    P.ops.push_back(o);
  }
};

static int alloc_stage(vv_engine* e, StageStash& st, const std::vector<BlockW>& ws, long long rows) {
  if (ws.empty()) return 0;
  const int G = ws[0].G, d = ws[0].d;
  const size_t rd = (size_t)G * rows * d;
  st.x.resize(ws.size() + 1);
  st.b.resize(ws.size());
  for (auto& p : st.x) if (!(p = dalloc<float>(e, rd))) return -1;
  for (auto& b : st.b) {
    b.qkv = dalloc<bf16>(e, 3 * rd); b.x1 = dalloc<float>(e, rd); b.u = dalloc<bf16>(e, 4 * rd);
    if (!b.qkv || !b.x1 || !b.u) return -1;
  }
  return 0;
}
static int alloc_stash(vv_engine* e, Net& n, Stash& S) {
  if (alloc_stage(e, S.e0, n.e0, n.L0) || alloc_stage(e, S.e1, n.e1, n.L1) || alloc_stage(e, S.lg, n.lg, n.L1) ||
      alloc_stage(e, S.u0, n.u0, n.L1) || alloc_stage(e, S.u1, n.u1, n.L0))
    return -1;
  S.EX = dalloc<float>(e, (size_t)n.G * n.L1 * 4 * n.D);
  return S.EX ? 0 : -1;
}

static int alloc_temps(vv_engine* e, Temps& t) {
  size_t m_rd = 0, m_l0d = 0, m_l1d = 0, m_l1e = 0, m_l1gd = 0;
  for (int k = 0; k < 2; ++k) {
    Net& n = e->net[k];
    if (!n.finalized) continue;
    const size_t G = n.G;
    m_rd = std::max({m_rd, G * n.L0 * n.D, G * n.L1 * 2 * n.D, (size_t)n.L1 * n.E});
    m_l0d = std::max(m_l0d, G * n.L0 * n.D);
    m_l1d = std::max(m_l1d, G * n.L1 * n.D);
    m_l1e = std::max(m_l1e, (size_t)n.L1 * n.E);
    m_l1gd = std::max(m_l1gd, G * n.L1 * 2 * n.D);
  }
  t.h = dalloc<bf16>(e, m_rd); t.ao = dalloc<bf16>(e, m_rd); t.a = dalloc<bf16>(e, 4 * m_rd); t.du = dalloc<bf16>(e, 4 * m_rd);
  t.dao = dalloc<bf16>(e, m_rd); t.dqkv = dalloc<bf16>(e, 3 * m_rd); t.dx1b = dalloc<bf16>(e, m_rd);
  t.dh = dalloc<float>(e, m_rd); t.dx1 = dalloc<float>(e, m_rd); t.dhb = dalloc<bf16>(e, m_rd);
  t.MB = dalloc<bf16>(e, 4 * m_l1d); t.EPIN = dalloc<bf16>(e, m_l1gd); t.TB = dalloc<bf16>(e, m_l1e);
  t.CAT0 = dalloc<bf16>(e, 4 * m_l1d); t.CAT1 = dalloc<bf16>(e, 2 * m_l0d); t.U1B = dalloc<bf16>(e, 2 * m_l1d);
  t.NU = dalloc<float>(e, m_l0d);
  {
    size_t tower_rows = 0, trunk_rows = 0;
    for (int k = 0; k < 2; ++k)
      if (e->net[k].finalized) { tower_rows = std::max(tower_rows, (size_t)e->net[k].G * e->net[k].L0); trunk_rows = std::max(trunk_rows, (size_t)e->net[k].L1); }
    t.lnst_cap = std::max(8 * tower_rows, 40 * trunk_rows);
    t.lnst = dalloc<float>(e, 2 * t.lnst_cap);
    t.lnshift = dalloc<float>(e, std::max(tower_rows, trunk_rows));
    if (!t.lnst || !t.lnshift) return -1;
  }
  t.gU1 = dalloc<float>(e, m_l0d); t.gU1b = dalloc<bf16>(e, m_l0d);
  t.gU0 = dalloc<float>(e, 2 * m_l1d); t.gU0b = dalloc<bf16>(e, 2 * m_l1d);
  t.gT = dalloc<float>(e, m_l1e); t.gTb = dalloc<bf16>(e, m_l1e);
  t.gE1 = dalloc<float>(e, 2 * m_l1d); t.gE1b = dalloc<bf16>(e, 2 * m_l1d);
  t.gE0 = dalloc<float>(e, m_l0d); t.gE0b = dalloc<bf16>(e, m_l0d);
  t.dNU = dalloc<float>(e, m_l0d); t.dC1a = dalloc<float>(e, m_l0d); t.dSK0 = dalloc<float>(e, m_l0d);
  t.dEX = dalloc<float>(e, 4 * m_l1d); t.dEXb = dalloc<bf16>(e, 4 * m_l1d);
  t.dSK1 = dalloc<float>(e, 2 * m_l1d); t.dEP = dalloc<float>(e, m_l1gd); t.dMB = dalloc<float>(e, 4 * m_l1d);
  t.DPb = dalloc<bf16>(e, m_l1gd);
  return (t.h && t.DPb && t.dMB && t.a && t.du) ? 0 : -1;
}

static int build_plans(vv_engine* e) {
  if (e->plans_built) return 0;
  VV_CHECK(e->net[0].finalized, "decoder weights not finalized");
  const int T = e->cfg.T;
  VV_CHECK(T == 1 || e->net[1].finalized, "flow weights not finalized");
  const bool flow = e->net[1].finalized;
  const int napp = flow ? std::max(T, 2) : 1;
  Temps t{};
  if (alloc_temps(e, t)) return -1;
  e->ln_health = dalloc<unsigned int>(e, 2);
  if (!e->ln_health) return -1;
  cudaMemset(e->ln_health, 0, 2 * sizeof(unsigned int));
  const size_t CHW = (size_t)e->C * e->HW;
  e->Z = dalloc<float>(e, (size_t)e->Zc * e->HW); e->GZ = dalloc<float>(e, (size_t)e->Zc * e->HW);
  e->DOUT = dalloc<float>(e, CHW); e->GD = dalloc<float>(e, CHW); e->XB = dalloc<float>(e, CHW);
  e->XN = dalloc<float>(e, CHW * napp); e->Gb[0] = dalloc<float>(e, CHW); e->Gb[1] = dalloc<float>(e, CHW);
  const int rb = reduce_blocks();
  e->partials = dalloc<double>(e, rb); e->dots = dalloc<double>(e, 8); e->dot_scratch = dalloc<double>(e, 4 * rb); e->Jbuf = dalloc<double>(e, 4);
  if (!e->Z || !e->XN || !e->Gb[1] || !e->Jbuf) return -1;
  e->stash.resize(napp); e->fwd.resize(napp); e->bwd.resize(napp);
  const bool share = e->cfg.recompute != 0;
  for (int a = 0; a < napp; ++a) {
    Net& n = e->net[a == 0 ? 0 : 1];
    if (share && a >= 2) e->stash[a] = e->stash[1];
    else if (alloc_stash(e, n, e->stash[a])) return -1;
    Builder B{e, &n, t};
    B.f16 = e->cfg.forward_fp16 ? 1 : 0;
    B.fold = e->cfg.no_ln_fold == 0;
    B.fuse_mlp = getenv("VV_NO_FUSED_MLP") == nullptr;
    if (a == 0) {
      B.net_fwd(e->fwd[a], e->stash[a], e->Z, e->DOUT);
      B.net_bwd(e->bwd[a], e->stash[a], e->GD, e->GZ);
    } else {
      B.net_fwd(e->fwd[a], e->stash[a], e->XN + (size_t)(a - 1) * CHW, e->XN + (size_t)a * CHW);
      B.net_bwd(e->bwd[a], e->stash[a], e->Gb[a % 2], e->Gb[(a - 1) % 2]);
    }
    VV_CHECK(!B.err, "plan construction failed: %s", B.err);
    const std::string who = a == 0 ? std::string("decoder") : "flow[" + std::to_string(a) + "]";
    e->fwd[a].label = who + " forward";
    e->bwd[a].label = who + " backward (input-VJP)";
  }
  // Chain the GEMMs in execution order (fwd[0..napp-1], then bwd[napp-1..0], wrapping around to the next evaluation): each
  // one prefetches the weights of its successor into L2 while its own epilogue drains.
  if (!getenv("VV_NO_WEIGHT_PREFETCH")) {
    std::vector<GemmDesc*> chain;
    for (int a = 0; a < napp; ++a)
      for (Op& o : e->fwd[a].ops) if (o.kind == Op::GEMM) chain.push_back(&o.gemm);
    for (int a = napp - 1; a >= 0; --a)
      for (Op& o : e->bwd[a].ops) if (o.kind == Op::GEMM) chain.push_back(&o.gemm);
    for (size_t i = 0; i < chain.size(); ++i) {
      const GemmDesc* nx = chain[(i + 1) % chain.size()];
      chain[i]->a.pf_ptr = nx->b_ptr;
      chain[i]->a.pf_bytes = nx->b_ptr ? nx->b_bytes : 0;
    }
    // a fused tower MLP streams fc1 | fc2 (backward: fc2^T | fc1^T, 1.8 / 7 MB for all six towers) through its weight ring from the first
    // cycle on: the GEMM launched right before it pulls that region into L2 as well
    auto mlp_prefetch = [](Plan& P) {
      for (size_t i = 1; i < P.ops.size(); ++i) {
        Op& o = P.ops[i];
        if (o.kind != Op::MLP_F && o.kind != Op::MLP_B) continue;
        for (size_t j = i; j-- > 0 && i - j <= 3;) {                  // the nearest GEMM in front of it (backward: behind a LayerNorm launch)
          if (P.ops[j].kind != Op::GEMM) continue;
          P.ops[j].gemm.a.pf2_ptr = o.mlp.w_ptr;
          P.ops[j].gemm.a.pf2_bytes = o.mlp.w_bytes;
          break;
        }
      }
    };
    for (int a = 0; a < napp; ++a) { mlp_prefetch(e->fwd[a]); mlp_prefetch(e->bwd[a]); }
  }
  e->plans_built = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// cost and gradient
// ---------------------------------------------------------------------------------------------
static int enqueue_forward(vv_engine* e, cudaStream_t s, bool with_obs) {
  int launches = 0;
  const int T = e->cfg.T;
  const size_t CHW = (size_t)e->C * e->HW;
  launches += e->fwd[0].run(s);
  if (!e->native) {
    // x_0 = xb + D(z) stdTr sigma  (da_4dvar.py:1187)  ->  normalised: (xb - mu)/sigma + D(z) stdTr
    launch_chan_affine(e->XN, e->DOUT, e->stdTr, e->XB, e->inv_sigma, e->neg_mu_sig, e->C, e->HW, s); ++launches;
    for (int t = 1; t < T; ++t) launches += e->fwd[t].run(s);                                // x_t = M(x_{t-1})  (:1191-1193)
  } else {
    // Analysis grid finer than the network grid (the reference's 721x1440 over 128x256).  Nearest resampling is an index map, so the
    // fields on the analysis grid are never materialised: with F_0 = D(z) stdTr and F_t = M(N_{t-1}) on the network grid,
    //   x_t on the analysis grid  = up(F_t) sigma + (xb for t = 0, mu otherwise)            (vae.py:90, da_4dvar.py:1187, 678-681)
    //   N_t (input of the next M) = down((x_t - mu) / sigma) = S(F_t) (+ down((xb - mu) / sigma) for t = 0),  S = down o up  (:667-671)
    // and the observation term reads up(F_t) through indices composed with the up-sampling map at vv_set_case_native.
    const int H = e->net[0].H, W = e->net[0].W;
    launch_chan_affine(e->XF, e->DOUT, e->stdTr, nullptr, nullptr, nullptr, e->C, e->HW, s);
    launch_seam_gather(e->XN, e->XF, e->XBN, e->s_row, e->s_col, e->C, H, W, s);
    launches += 2;
    for (int t = 1; t < T; ++t) {
      launches += e->fwd[t].run(s);
      cudaMemcpyAsync(e->XF + (size_t)t * CHW, e->XN + (size_t)t * CHW, CHW * sizeof(float), cudaMemcpyDeviceToDevice, s);
      if (t + 1 < T) { launch_seam_gather(e->XN + (size_t)t * CHW, e->XF + (size_t)t * CHW, nullptr, e->s_row, e->s_col, e->C, H, W, s); ++launches; }
    }
  }
  if (with_obs) {
    nvtx_push("observation term + J");
    if (e->taps)
      launch_obs_taps_misfit(e->XF, e->tap_ia, e->tap_coef, e->taps, e->yobs, e->rinv, e->n_obs, e->obs_coeff, e->resid, e->partials,
                             reduce_blocks(), s);
    else
      launch_obs_misfit(e->native ? e->XF : e->XN, e->idx, e->yobs, e->rinv, e->sigma, e->mean, e->n_obs, e->HW, e->C, e->obs_coeff, e->resid,
                        e->partials, reduce_blocks(), s);
    launch_reduce_partials(e->partials, reduce_blocks(), e->Jbuf + 3, s);
    DotPairs dp{}; dp.a[0] = e->Z; dp.b[0] = e->Z; dp.n_pairs = 1;
    launch_multi_dot(dp, (long long)e->Zc * e->HW, e->dots, e->dot_scratch, s);
    finalize_J_kernel<<<1, 1, 0, s>>>(e->dots, e->Jbuf + 3, e->obs_coeff, e->Jbuf);
    launches += 5;
    nvtx_pop();
  }
  (void)CHW;
  return launches;
}

static int enqueue_backward(vv_engine* e, cudaStream_t s) {
  int launches = 0;
  const int T = e->cfg.T;
  const size_t CHW = (size_t)e->C * e->HW;
  float* Gt = e->Gb[(T - 1) % 2];
  cudaMemsetAsync(Gt, 0, CHW * sizeof(float), s);
  // native geometry: several observations share a network-grid cell (sorted runs); otherwise the indices are unique
  auto obs_adjoint = [&](float* G, int t) {
    if (e->taps)
      launch_obs_taps_adjoint(G, e->pair_cell, e->pair_src, e->pair_coef, e->resid, e->obs_off[t] * e->taps, e->obs_off[t + 1] * e->taps,
                              (long long)t * CHW, s);
    else if (e->native) launch_obs_adjoint_runs(G, e->idx, e->resid, e->obs_off[t], e->obs_off[t + 1], (long long)t * CHW, s);
    else launch_obs_adjoint(G, e->idx, e->resid, e->obs_off[t], e->obs_off[t + 1], (long long)t * CHW, s);
  };
  obs_adjoint(Gt, T - 1); ++launches;
  for (int t = T - 1; t >= 1; --t) {
    if (e->cfg.recompute && t != T - 1) launches += e->fwd[t].run(s);      // stash shared by the flow applications: rebuild step t
    launches += e->bwd[t].run(s);
    float* G = e->Gb[(t - 1) % 2];
    if (e->native) {                                                       // dJ/dN_{t-1} -> dJ/dF_{t-1} through S^T
      cudaMemcpyAsync(e->TMPF, G, CHW * sizeof(float), cudaMemcpyDeviceToDevice, s);
      launch_seam_gather_adjoint(G, e->TMPF, e->s_row_lo, e->s_col_lo, e->C, e->net[0].H, e->net[0].W, s); ++launches;
    }
    obs_adjoint(G, t - 1); ++launches;
  }
  launch_chan_affine(e->GD, e->Gb[0], e->stdTr, nullptr, nullptr, nullptr, e->C, e->HW, s); ++launches;   // dJ/dD = G_0 stdTr
  launches += e->bwd[0].run(s);
  launch_axpby(e->GZ, e->Z, nullptr, 1.0, nullptr, 1.0, (long long)e->Zc * e->HW, s); ++launches;         // + d(|z|^2/2)/dz
  return launches;
}

namespace vv {
int fence_in(vv_engine* e, cudaStream_t user) {
  VV_CUDA(cudaEventRecord(e->ev_in, user));
  VV_CUDA(cudaStreamWaitEvent(e->stream, e->ev_in, 0));
  return 0;
}
int fence_out(vv_engine* e, cudaStream_t user) {
  VV_CUDA(cudaEventRecord(e->ev_out, e->stream));
  VV_CUDA(cudaStreamWaitEvent(user, e->ev_out, 0));
  return 0;
}
int engine_cost_grad(vv_engine* e, const float* z, double* Jout, float* grad, cudaStream_t s) {
  VV_CHECK(e->have_consts, "vv_set_constants has not been called");
  VV_CHECK(e->have_case, "vv_set_case has not been called");
  int rc = build_plans(e);
  if (rc) return rc;
  const size_t zbytes = (size_t)e->Zc * e->HW * sizeof(float);
  if (z != e->Z) VV_CUDA(cudaMemcpyAsync(e->Z, z, zbytes, cudaMemcpyDeviceToDevice, s));
  if (e->cfg.use_graph && e->eager_runs >= 1) {
    if (!e->graph_cg) {
      cudaGraph_t graph;
      VV_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      int l = enqueue_forward(e, s, true);
      l += enqueue_backward(e, s);
      e->last_launches = l;
      VV_CUDA(cudaStreamEndCapture(s, &graph));
      VV_CUDA(cudaGraphInstantiate(&e->graph_cg, graph, 0));
      cudaGraphDestroy(graph);
    }
    VV_CUDA(cudaGraphLaunch(e->graph_cg, s));
  } else {
    int l = enqueue_forward(e, s, true);
    l += enqueue_backward(e, s);
    e->last_launches = l;
    e->eager_runs++;
  }
  if (Jout) VV_CUDA(cudaMemcpyAsync(Jout, e->Jbuf, 3 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (grad && grad != e->GZ) VV_CUDA(cudaMemcpyAsync(grad, e->GZ, zbytes, cudaMemcpyDeviceToDevice, s));
  VV_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace vv

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

VV_API const char* vv_last_error(void) { return g_err; }

VV_API int vv_set_device(int ordinal) {
  VV_CUDA(cudaSetDevice(ordinal));
  return 0;
}

VV_API int vv_engine_create(const vv_config* cfg, vv_engine** out) {
  VV_CHECK(cfg && out, "null argument");
  int dev = 0;
  VV_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VV_CUDA(cudaGetDeviceProperties(&prop, dev));
  VV_CHECK(prop.major == 10, "vaevar_b200 needs an sm_100 GPU (found sm_%d%d); there is no fallback path", prop.major, prop.minor);
  VV_CHECK(cfg->T >= 1, "T must be >= 1");
  vv_engine* e = new vv_engine();
  e->cfg = *cfg;
  int rc = net_init(e->net[0], cfg->dec);
  if (!rc && cfg->has_flow) rc = net_init(e->net[1], cfg->flow);
  if (rc) { delete e; return rc; }
  if (cfg->T > 1 && !cfg->has_flow) { set_error("T > 1 needs a flow network"); delete e; return -2; }
  e->C = e->net[0].ckeep;
  e->Zc = e->net[0].cin;
  e->HW = (long long)cfg->dec.img_h * cfg->dec.img_w;
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming) != cudaSuccess) {
    set_error("could not create the engine stream / events");
    delete e; return -1;
  }
  if (cfg->has_flow) {
    Net& f = e->net[1];
    if (f.cin != e->C || f.ckeep != e->C || f.H != e->net[0].H || f.W != e->net[0].W) {
      set_error("flow network must map %d -> %d channels on the decoder grid", e->C, e->C);
      delete e; return -2;
    }
  }
  *out = e;
  return 0;
}

VV_API void vv_engine_destroy(vv_engine* e) {
  if (!e) return;
  cudaDeviceSynchronize();
  if (e->graph_cg) cudaGraphExecDestroy(e->graph_cg);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  for (int k = 0; k < 2; ++k)
    for (auto& kv : e->net[k].staged) cudaFree(kv.second.first);
  for (void* p : e->allocs) cudaFree(p);
  delete e;
}

VV_API int vv_set_weight(vv_engine* e, int net, const char* name, const float* data_dev, const int64_t* shape, int ndim) {
  VV_CHECK(e && name && data_dev && net >= 0 && net < 2, "bad argument");
  Net& n = e->net[net];
  VV_CHECK(!n.finalized, "weights already finalized");
  const std::string nm(name);
  if (nm.find("relative_position_index") != std::string::npos || nm.find("attn_mask") != std::string::npos) return 0;
  size_t numel = 1;
  std::vector<int64_t> shp(shape, shape + ndim);
  for (auto v : shp) numel *= (size_t)v;
  float* d = nullptr;
  VV_CUDA(cudaMalloc(&d, numel * sizeof(float)));
  VV_CUDA(cudaMemcpy(d, data_dev, numel * sizeof(float), cudaMemcpyDeviceToDevice));
  auto it = n.staged.find(nm);
  if (it != n.staged.end()) cudaFree(it->second.first);
  n.staged[nm] = {d, shp};
  return 0;
}

VV_API int vv_finalize_weights(vv_engine* e) {
  VV_CHECK(e, "null engine");
  int rc = finalize_net(e, e->net[0]);
  if (!rc && e->cfg.has_flow) rc = finalize_net(e, e->net[1]);
  return rc;
}

VV_API int vv_set_constants(vv_engine* e, const float* mean, const float* std, const float* stdTr) {
  VV_CHECK(e && mean && std && stdTr, "null argument");
  const int C = e->C;
  std::vector<float> m(mean, mean + C), s(std, std + C), t(stdTr, stdTr + C), is(C), nm(C);
  for (int c = 0; c < C; ++c) { is[c] = 1.0f / s[c]; nm[c] = -m[c] / s[c]; }
  if (e->have_consts) {            // a second call updates the constants in place: the launch graph keeps the same pointers
    const size_t nb = (size_t)C * sizeof(float);
    VV_CUDA(cudaDeviceSynchronize());
    VV_CUDA(cudaMemcpy(e->mean, m.data(), nb, cudaMemcpyHostToDevice)); VV_CUDA(cudaMemcpy(e->sigma, s.data(), nb, cudaMemcpyHostToDevice));
    VV_CUDA(cudaMemcpy(e->stdTr, t.data(), nb, cudaMemcpyHostToDevice)); VV_CUDA(cudaMemcpy(e->inv_sigma, is.data(), nb, cudaMemcpyHostToDevice));
    VV_CUDA(cudaMemcpy(e->neg_mu_sig, nm.data(), nb, cudaMemcpyHostToDevice));
  } else {
    e->mean = dupload(e, m); e->sigma = dupload(e, s); e->stdTr = dupload(e, t); e->inv_sigma = dupload(e, is); e->neg_mu_sig = dupload(e, nm);
    VV_CHECK(e->mean && e->sigma && e->stdTr && e->inv_sigma && e->neg_mu_sig, "out of memory for the constants");
  }
  e->have_consts = true;
  e->generation++;
  return 0;
}

VV_API int vv_compact_mask(const float* H_dev, const float* yo_dev, const float* R_dev, int64_t n, int32_t* idx_out_dev, float* y_out_dev,
                    float* rinv_out_dev, int64_t* n_out_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const long long nchunks = (n + 1023) / 1024;
  int* counts = nullptr;
  VV_CUDA(cudaMalloc(&counts, (nchunks + 1) * sizeof(int)));
  launch_compact_count(H_dev, n, counts, s);
  launch_compact_scan(counts, nchunks, s);
  launch_compact_write(H_dev, yo_dev, R_dev, n, counts, idx_out_dev, y_out_dev, rinv_out_dev, s);
  int total = 0;
  cudaError_t er = cudaMemcpyAsync(&total, counts + nchunks, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (er == cudaSuccess) er = cudaStreamSynchronize(s);
  cudaFree(counts);
  VV_CHECK(er == cudaSuccess, "vv_compact_mask: %s", cudaGetErrorString(er));
  if (n_out_host) *n_out_host = total;
  return 0;
}

// (Re-)allocate the per-observation buffers for `total` observations.  The old buffers are freed first and the capacity is only
// raised once every allocation has succeeded; a launch graph that baked the old pointers in is dropped.
static int reserve_obs(vv_engine* e, long long total) {
  if (total <= e->obs_cap) return 0;
  const long long cap = total + total / 8 + 1024;
  if (e->graph_cg) { cudaGraphExecDestroy(e->graph_cg); e->graph_cg = nullptr; }
  cudaDeviceSynchronize();
  dfree(e, e->idx); dfree(e, e->yobs); dfree(e, e->rinv); dfree(e, e->resid);
  e->obs_cap = 0;
  e->idx = dalloc<int>(e, cap); e->yobs = dalloc<float>(e, cap); e->rinv = dalloc<float>(e, cap); e->resid = dalloc<float>(e, cap);
  VV_CHECK(e->idx && e->yobs && e->rinv && e->resid, "out of memory for %lld observations", total);
  e->obs_cap = cap;
  return 0;
}

VV_API int vv_set_case(vv_engine* e, const float* xb_dev, const float* yo_dev, const float* H_dev, const float* R_dev, float obs_coeff, void* stream) {
  VV_CHECK(e && xb_dev && yo_dev && H_dev && R_dev, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = build_plans(e);
  if (rc) return rc;
  const int T = e->cfg.T;
  const long long CHW = (long long)e->C * e->HW, n = CHW * T;
  VV_CHECK(CHW % 1024 == 0, "C*H*W must be a multiple of 1024");
  VV_CHECK(n < (1LL << 31), "observation space too large for int32 indices");
  const long long nchunks = n / 1024;
  if (!e->chunk_counts) e->chunk_counts = dalloc<int>(e, nchunks + 1);
  launch_compact_count(H_dev, n, e->chunk_counts, s);
  launch_compact_scan(e->chunk_counts, nchunks, s);
  std::vector<int> offs(nchunks + 1);
  VV_CUDA(cudaMemcpyAsync(offs.data(), e->chunk_counts, (nchunks + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
  VV_CUDA(cudaStreamSynchronize(s));
  const long long total = offs[nchunks];
  if ((rc = reserve_obs(e, total))) return rc;
  launch_compact_write(H_dev, yo_dev, R_dev, n, e->chunk_counts, e->idx, e->yobs, e->rinv, s);
  std::vector<long long> off(T + 1);
  for (int t = 0; t <= T; ++t) off[t] = offs[(size_t)(t * (CHW / 1024))];
  if (e->graph_cg && (off != e->obs_off || total != e->n_obs || obs_coeff != e->obs_coeff)) {
    cudaGraphExecDestroy(e->graph_cg); e->graph_cg = nullptr;     // launch parameters baked into the graph changed
  }
  if (e->native && e->graph_cg) { cudaGraphExecDestroy(e->graph_cg); e->graph_cg = nullptr; }
  e->native = false;
  e->taps = 0;
  e->obs_off = off;
  e->n_obs = total;
  e->obs_coeff = obs_coeff;
  VV_CUDA(cudaMemcpyAsync(e->XB, xb_dev, CHW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  VV_CUDA(cudaStreamSynchronize(s));
  e->have_case = true;
  e->generation++;
  return 0;
}

// A observed channels with K taps each (K = 0: the C state channels observed directly).
static int set_case_native_impl(vv_engine* e, const float* xb_dev, const float* yo_dev, const float* H_dev, const float* R_dev, int Hh, int Wh,
                                int A, int K, const int* tap_chan_host, const float* tap_w_host, float obs_coeff, void* stream) {
  VV_CHECK(e && xb_dev && yo_dev && H_dev && R_dev && Hh >= 1 && Wh >= 1, "bad argument");
  VV_CHECK(e->have_consts, "vv_set_constants has not been called");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = build_plans(e);
  if (rc) return rc;
  const int T = e->cfg.T, C = e->C, H = e->net[0].H, W = e->net[0].W;
  const long long CHW = (long long)C * e->HW, lvl = (long long)A * Hh * Wh, lvl_x = (long long)C * Hh * Wh;
  VV_CHECK(CHW * T < (1LL << 31) && lvl < (1LL << 31) && lvl_x < (1LL << 31), "observation space too large for int32 indices");
  // ordered compaction per time level on the analysis grid (== torch.nonzero order within the level)
  const long long nchunks = (lvl + 1023) / 1024;
  int* counts = nullptr;
  VV_CUDA(cudaMalloc(&counts, (size_t)T * (nchunks + 1) * sizeof(int)));
  std::vector<long long> off(T + 1, 0);
  std::vector<int> tot(T);
  for (int t = 0; t < T; ++t) {
    int* ct = counts + (size_t)t * (nchunks + 1);
    launch_compact_count(H_dev + (size_t)t * lvl, lvl, ct, s);
    launch_compact_scan(ct, nchunks, s);
    cudaMemcpyAsync(&tot[t], ct + nchunks, sizeof(int), cudaMemcpyDeviceToHost, s);
  }
  cudaError_t er = cudaStreamSynchronize(s);
  if (er != cudaSuccess) { cudaFree(counts); set_error("vv_set_case_native: %s", cudaGetErrorString(er)); return -1; }
  for (int t = 0; t < T; ++t) off[t + 1] = off[t] + tot[t];
  const long long total = off[T];
  int* idx_hr = nullptr; float *y_hr = nullptr, *ri_hr = nullptr;
  const size_t cap = (size_t)(total ? total : 1);
  if (cudaMalloc(&idx_hr, cap * sizeof(int)) != cudaSuccess || cudaMalloc(&y_hr, cap * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&ri_hr, cap * sizeof(float)) != cudaSuccess) {
    cudaFree(counts); cudaFree(idx_hr); cudaFree(y_hr); cudaFree(ri_hr);
    set_error("vv_set_case_native: out of memory for %lld observations", total);
    return -1;
  }
  for (int t = 0; t < T; ++t)
    launch_compact_write(H_dev + (size_t)t * lvl, yo_dev + (size_t)t * lvl, R_dev + (size_t)t * lvl, lvl, counts + (size_t)t * (nchunks + 1),
                         idx_hr + off[t], y_hr + off[t], ri_hr + off[t], s);
  rc = reserve_obs(e, total);
  if (rc) { cudaFree(counts); cudaFree(idx_hr); cudaFree(y_hr); cudaFree(ri_hr); return rc; }
  if (!e->XF) {
    e->XF = dalloc<float>(e, (size_t)CHW * T); e->XBN = dalloc<float>(e, CHW); e->TMPF = dalloc<float>(e, CHW);
    e->s_row = dalloc<int>(e, H); e->s_col = dalloc<int>(e, W); e->s_row_lo = dalloc<int>(e, H + 1); e->s_col_lo = dalloc<int>(e, W + 1);
  }
  if (lvl_x > e->xbh_cap) {
    cudaDeviceSynchronize();
    dfree(e, e->XBH); e->xbh_cap = 0;
    e->XBH = dalloc<float>(e, lvl_x);
    if (e->XBH) e->xbh_cap = lvl_x;
  }
  rc = (e->idx && e->yobs && e->rinv && e->resid && e->XF && e->XBN && e->TMPF && e->s_col_lo && e->XBH) ? 0 : -1;
  int* tap_chan = nullptr; float* tap_w = nullptr;
  if (!rc && K > 0) {
    if (total * K >= (1LL << 31)) { set_error("too many observation taps for int32 indices"); rc = -2; }
    if (!rc && total * K > e->pair_cap) {
      const long long cap = total * K + total * K / 8 + 1024;
      cudaDeviceSynchronize();
      dfree(e, e->tap_ia); dfree(e, e->pair_cell); dfree(e, e->pair_src); dfree(e, e->tap_coef); dfree(e, e->pair_coef);
      e->pair_cap = 0;
      e->tap_ia = dalloc<int>(e, cap); e->pair_cell = dalloc<int>(e, cap); e->pair_src = dalloc<int>(e, cap);
      e->tap_coef = dalloc<float>(e, cap); e->pair_coef = dalloc<float>(e, cap);
      if (!e->pair_coef || !e->tap_ia || !e->pair_cell || !e->pair_src || !e->tap_coef) rc = -1;
      else e->pair_cap = cap;
    }
    if (!rc && (cudaMalloc(&tap_chan, (size_t)A * K * sizeof(int)) != cudaSuccess || cudaMalloc(&tap_w, (size_t)A * K * sizeof(float)) != cudaSuccess)) rc = -1;
    if (!rc) {
      cudaMemcpyAsync(tap_chan, tap_chan_host, (size_t)A * K * sizeof(int), cudaMemcpyHostToDevice, s);
      cudaMemcpyAsync(tap_w, tap_w_host, (size_t)A * K * sizeof(float), cudaMemcpyHostToDevice, s);
    }
  }
  if (!rc) {
    std::vector<int> row, col, row_lo, col_lo;
    host_seam_tables(H, W, Hh, Wh, row, col, row_lo, col_lo);
    cudaMemcpyAsync(e->s_row, row.data(), H * sizeof(int), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(e->s_col, col.data(), W * sizeof(int), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(e->s_row_lo, row_lo.data(), (H + 1) * sizeof(int), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(e->s_col_lo, col_lo.data(), (W + 1) * sizeof(int), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(e->XBH, xb_dev, lvl_x * sizeof(float), cudaMemcpyDeviceToDevice, s);
    launch_resample(xb_dev, e->XBN, C, Hh, Wh, H, W, 1, e->mean, e->sigma, s);               // down((xb - mu) / sigma), da_4dvar.py:667-671
    cudaMemsetAsync(e->XB, 0, CHW * sizeof(float), s);
    if (K > 0)
      rc = native_compose_taps(idx_hr, y_hr, ri_hr, off.data(), T, xb_dev, e->mean, e->sigma, tap_chan, tap_w, K, C, H, W, Hh, Wh, e->tap_ia,
                               e->tap_coef, e->yobs, e->rinv, e->pair_cell, e->pair_src, e->pair_coef, s);
    else
      rc = native_compose_sort(idx_hr, y_hr, ri_hr, off.data(), T, xb_dev, e->mean, C, H, W, Hh, Wh, e->idx, e->yobs, e->rinv, s);
    if (!rc && cudaStreamSynchronize(s) != cudaSuccess) { set_error("vv_set_case_native: %s", cudaGetErrorString(cudaGetLastError())); rc = -1; }
  } else if (rc == -1) {
    set_error("vv_set_case_native: out of memory");
  }
  cudaFree(counts); cudaFree(idx_hr); cudaFree(y_hr); cudaFree(ri_hr); cudaFree(tap_chan); cudaFree(tap_w);
  if (rc) return rc;
  if (e->graph_cg) { cudaGraphExecDestroy(e->graph_cg); e->graph_cg = nullptr; }            // launch parameters baked into the graph changed
  e->native = true; e->Hh = Hh; e->Wh = Wh;
  e->taps = K;
  e->obs_off = off;
  e->n_obs = total;
  e->obs_coeff = obs_coeff;
  e->have_case = true;
  e->generation++;
  return 0;
}

VV_API int vv_set_case_native(vv_engine* e, const float* xb_dev, const float* yo_dev, const float* H_dev, const float* R_dev, int Hh, int Wh,
                              float obs_coeff, void* stream) {
  return set_case_native_impl(e, xb_dev, yo_dev, H_dev, R_dev, Hh, Wh, e ? e->C : 0, 0, nullptr, nullptr, obs_coeff, stream);
}

VV_API int vv_set_case_obsop(vv_engine* e, const float* xb_dev, const float* yo_dev, const float* H_dev, const float* R_dev, int Hh, int Wh,
                             int n_obs_channels, int taps, const int32_t* tap_chan_host, const float* tap_w_host, float obs_coeff,
                             void* stream) {
  VV_CHECK(e && n_obs_channels >= 1 && taps >= 1 && taps <= 64 && tap_chan_host && tap_w_host, "bad observation operator");
  for (int i = 0; i < n_obs_channels * taps; ++i)
    VV_CHECK(tap_chan_host[i] >= 0 && tap_chan_host[i] < e->C, "tap %d reads state channel %d (of %d)", i, tap_chan_host[i], e->C);
  return set_case_native_impl(e, xb_dev, yo_dev, H_dev, R_dev, Hh, Wh, n_obs_channels, taps, tap_chan_host, tap_w_host, obs_coeff, stream);
}

VV_API int vv_decode_native(vv_engine* e, const float* z_dev, float* x_phys_out_dev, void* stream) {
  VV_CHECK(e && z_dev && x_phys_out_dev, "null argument");
  VV_CHECK(e->have_consts && e->have_case && e->native, "vv_set_case_native has not been called");
  cudaStream_t s = (cudaStream_t)stream;
  VV_CUDA(cudaMemcpyAsync(e->Z, z_dev, (size_t)e->Zc * e->HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  e->fwd[0].run(s);
  launch_decode_hr(e->DOUT, e->stdTr, e->sigma, e->XBH, x_phys_out_dev, e->C, e->net[0].H, e->net[0].W, e->Hh, e->Wh, s);
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_metrics_grid(vv_engine* e, const float* x_phys_dev, const float* gt_phys_dev, int H, int W, double* out_dev, void* stream) {
  VV_CHECK(e && x_phys_dev && gt_phys_dev && out_dev && H >= 2 && W >= 1, "bad argument");
  VV_CHECK(e->have_consts, "vv_set_constants has not been called");
  if (!e->met_part) {
    e->met_part = dalloc<double>(e, (size_t)metrics_scratch_doubles(e->C));
    VV_CHECK(e->met_part, "out of memory for the metric scratch");
  }
  if (H > e->met_w_cap) {
    e->met_w = dalloc<float>(e, (size_t)H);
    VV_CHECK(e->met_w, "out of memory for the metric scratch");
    e->met_w_cap = H;
  }
  launch_metrics(x_phys_dev, gt_phys_dev, e->mean, e->sigma, e->C, H, W, e->met_w, e->met_part, out_dev, (cudaStream_t)stream);
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_metrics(vv_engine* e, const float* x_phys_dev, const float* gt_phys_dev, double* out_dev, void* stream) {
  VV_CHECK(e, "null argument");
  return vv_metrics_grid(e, x_phys_dev, gt_phys_dev, e->net[0].H, e->net[0].W, out_dev, stream);
}

VV_API int vv_num_obs(vv_engine* e, int64_t* n_obs) {
  VV_CHECK(e && n_obs, "null argument");
  *n_obs = e->n_obs;
  return 0;
}

}  // extern "C"

// Debug (VV_NAN_PROBE=1, called by the optimiser the first time a closure returns a non-finite J): where along the window did the
// first non-finite value appear?  Non-finite count and largest finite magnitude of the latent, the decoder output, every state of
// the trajectory and every residual-stream buffer of every application's stash, in execution order, to stderr.
__global__ void nan_probe_kernel(const float* x, long long n, unsigned int* out) {
  unsigned int bad = 0, mx = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (isfinite(v)) mx = max(mx, __float_as_uint(fabsf(v))); else ++bad;
  }
  if (bad) atomicAdd(out, bad);
  atomicMax(out + 1, mx);
}
namespace vv {
void engine_nan_probe(vv_engine* e, const float* z) {
  unsigned int* d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return;
  auto probe = [&](const char* what, int app, int k, const float* x, size_t n) {
    if (!x || !n) return;
    cudaMemset(d, 0, 8);
    nan_probe_kernel<<<592, 256, 0, e->stream>>>(x, (long long)n, d);
    unsigned int h[2] = {0, 0};
    cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, e->stream);
    cudaStreamSynchronize(e->stream);
    float mx; memcpy(&mx, &h[1], 4);
    fprintf(stderr, "[vaevar nan-probe] %-10s app %d #%d: non-finite %u of %zu, max |finite| %.4g\n", what, app, k, h[0], n, mx);
  };
  const size_t CHW = (size_t)e->C * e->HW;
  probe("z", -1, 0, z, (size_t)e->Zc * e->HW);
  for (size_t a = 0; a < e->stash.size(); ++a) {
    Net& n = e->net[a == 0 ? 0 : 1];
    Stash& S = e->stash[a];
    if (a >= 1) probe("x_in", (int)a, 0, e->XN + (a - 1) * CHW, CHW);
    struct { const char* nm; StageStash* st; size_t rows; } stages[5] = {{"enc0", &S.e0, (size_t)n.L0}, {"enc1", &S.e1, (size_t)n.L1}, {"trunk", &S.lg, (size_t)n.L1},
                                                                            {"dec0", &S.u0, (size_t)n.L1}, {"dec1", &S.u1, (size_t)n.L0}};
    const std::vector<BlockW>* ws[5] = {&n.e0, &n.e1, &n.lg, &n.u0, &n.u1};
    for (int q = 0; q < 5; ++q) {
      if (ws[q]->empty()) continue;
      const size_t rd = (size_t)(*ws[q])[0].G * stages[q].rows * (*ws[q])[0].d;
      for (size_t k = 0; k < stages[q].st->x.size(); ++k) {
        probe(stages[q].nm, (int)a, (int)k, stages[q].st->x[k], rd);
        if (k < stages[q].st->b.size()) probe("  .x1", (int)a, (int)k, stages[q].st->b[k].x1, rd);
      }
    }
    if (a == 0) probe("D(z)", 0, 0, e->DOUT, CHW);
    else probe("x_out", (int)a, 0, e->XN + a * CHW, CHW);
  }
  cudaFree(d);
}
}  // namespace vv

extern "C" {

VV_API int vv_ln_fold_health(vv_engine* e, uint32_t counts_host[2]) {
  VV_CHECK(e && counts_host, "null argument");
  counts_host[0] = counts_host[1] = 0;
  if (!e->ln_health) return 0;
  VV_CUDA(cudaStreamSynchronize(e->stream));
  VV_CUDA(cudaDeviceSynchronize());
  VV_CUDA(cudaMemcpy(counts_host, e->ln_health, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  VV_CUDA(cudaMemset(e->ln_health, 0, 2 * sizeof(uint32_t)));
  return 0;
}

VV_API int vv_cost_grad(vv_engine* e, const float* z_dev, double* J_out_dev, float* grad_dev, void* stream) {
  VV_CHECK(e && z_dev, "null argument");
  cudaStream_t user = (cudaStream_t)stream;
  int rc = fence_in(e, user);
  if (!rc) rc = engine_cost_grad(e, z_dev, J_out_dev, grad_dev, e->stream);
  if (!rc) rc = fence_out(e, user);
  return rc;
}

VV_API int vv_cost(vv_engine* e, const float* z_dev, double* J_out_dev, void* stream) {
  VV_CHECK(e && z_dev, "null argument");
  VV_CHECK(e->have_consts && e->have_case, "constants / case not set");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = build_plans(e);
  if (rc) return rc;
  VV_CUDA(cudaMemcpyAsync(e->Z, z_dev, (size_t)e->Zc * e->HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  enqueue_forward(e, s, true);
  if (J_out_dev) VV_CUDA(cudaMemcpyAsync(J_out_dev, e->Jbuf, 3 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_decode(vv_engine* e, const float* z_dev, float* x_phys_out_dev, void* stream) {
  VV_CHECK(e && z_dev && x_phys_out_dev, "null argument");
  VV_CHECK(e->have_consts && e->have_case, "constants / case (xb) not set");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = build_plans(e);
  if (rc) return rc;
  VV_CUDA(cudaMemcpyAsync(e->Z, z_dev, (size_t)e->Zc * e->HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  e->fwd[0].run(s);
  // (D(z) stdTr) sigma + xb   (da_4dvar.py:1259, 1306)
  launch_chan_affine(e->DOUT, e->DOUT, e->stdTr, nullptr, nullptr, nullptr, e->C, e->HW, s);
  launch_chan_affine(x_phys_out_dev, e->DOUT, e->sigma, e->XB, nullptr, nullptr, e->C, e->HW, s);
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_integrate(vv_engine* e, const float* x_in, float* x_out, int steps, void* stream) {
  VV_CHECK(e && x_in && x_out && steps >= 1, "bad argument");
  VV_CHECK(e->net[1].finalized && e->have_consts, "flow network / constants not set");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = build_plans(e);
  if (rc) return rc;
  const size_t CHW = (size_t)e->C * e->HW;
  // normalise (da_4dvar.py:667), apply M `steps` times (:673-676), de-normalise (:681)
  launch_chan_affine(e->XN, x_in, e->inv_sigma, nullptr, nullptr, e->neg_mu_sig, e->C, e->HW, s);
  for (int k = 0; k < steps; ++k) {
    e->fwd[1].run(s);
    if (k + 1 < steps) VV_CUDA(cudaMemcpyAsync(e->XN, e->XN + CHW, CHW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  launch_chan_affine(x_out, e->XN + CHW, e->sigma, nullptr, nullptr, e->mean, e->C, e->HW, s);
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_net_forward(vv_engine* e, int net, const float* in_dev, float* out_dev, void* stream) {
  VV_CHECK(e && in_dev && out_dev && (net == 0 || net == 1), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = build_plans(e);
  if (rc) return rc;
  const size_t CHW = (size_t)e->C * e->HW;
  Net& n = e->net[net];
  VV_CHECK(n.finalized, "network %d has no weights", net);
  float* in = net == 0 ? e->Z : e->XN;
  float* out = net == 0 ? e->DOUT : e->XN + CHW;
  VV_CUDA(cudaMemcpyAsync(in, in_dev, (size_t)n.cin * e->HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  e->fwd[net].run(s);
  VV_CUDA(cudaMemcpyAsync(out_dev, out, (size_t)n.ckeep * e->HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_net_vjp(vv_engine* e, int net, const float* in_dev, const float* dout_dev, float* din_dev, void* stream) {
  VV_CHECK(e && in_dev && dout_dev && din_dev && (net == 0 || net == 1), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = build_plans(e);
  if (rc) return rc;
  Net& n = e->net[net];
  VV_CHECK(n.finalized, "network %d has no weights", net);
  float* in = net == 0 ? e->Z : e->XN;
  float* dout = net == 0 ? e->GD : e->Gb[1];
  float* din = net == 0 ? e->GZ : e->Gb[0];
  VV_CUDA(cudaMemcpyAsync(in, in_dev, (size_t)n.cin * e->HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  VV_CUDA(cudaMemcpyAsync(dout, dout_dev, (size_t)n.ckeep * e->HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  e->fwd[net].run(s);
  e->bwd[net].run(s);
  VV_CUDA(cudaMemcpyAsync(din_dev, din, (size_t)n.cin * e->HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_test_obs(vv_engine* e, const float* xn_dev, double* J_obs_dev, float* grad_xn_dev, void* stream) {
  VV_CHECK(e && xn_dev && e->have_case && e->have_consts, "bad argument / case not set");
  cudaStream_t s = (cudaStream_t)stream;
  const int T = e->cfg.T;
  const size_t CHW = (size_t)e->C * e->HW;
  launch_obs_misfit(xn_dev, e->idx, e->yobs, e->rinv, e->sigma, e->mean, e->n_obs, e->HW, e->C, e->obs_coeff, e->resid, e->partials,
                    reduce_blocks(), s);
  launch_reduce_partials(e->partials, reduce_blocks(), J_obs_dev, s);
  if (grad_xn_dev) {
    VV_CUDA(cudaMemsetAsync(grad_xn_dev, 0, CHW * T * sizeof(float), s));
    for (int t = 0; t < T; ++t)
      launch_obs_adjoint(grad_xn_dev + (size_t)t * CHW, e->idx, e->resid, e->obs_off[t], e->obs_off[t + 1], (long long)t * CHW, s);
  }
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_last_launch_count(vv_engine* e) { return e ? e->last_launches : 0; }

VV_API int vv_profile_ops(vv_engine* e, int app, int bwd, int reps, float* ms_out, int* kind_out, double* flop_out, int* mnk_out, int cap,
                          int flush_l2) {
  VV_CHECK(e && ms_out && kind_out && reps >= 1, "bad argument");
  int rc = build_plans(e);
  if (rc) return rc;
  VV_CHECK(app >= 0 && app < (int)e->fwd.size(), "no such application");
  const Plan& P = bwd ? e->bwd[app] : e->fwd[app];
  cudaEvent_t e0, e1;
  VV_CUDA(cudaEventCreate(&e0)); VV_CUDA(cudaEventCreate(&e1));
  cudaStream_t s = e->stream;
  // flush_l2: every timed launch starts on a cold L2 (a 256 MiB scratch is overwritten first, outside the timed events) -- the
  // HBM-bound kernels are then rated on DRAM traffic, not on a working set that stayed in the 126 MB L2 between repetitions.
  const size_t FLUSH = (size_t)256 << 20;
  void* scratch = nullptr;
  if (flush_l2) VV_CUDA(cudaMalloc(&scratch, FLUSH));
  e->fwd[app].run(s);                       // make sure the stash holds finite values
  int k = 0;
  for (const Op& o : P.ops) {
    Plan one; one.ops.push_back(o);
    one.run(s);
    float ms = 0.f;
    if (flush_l2) {
      for (int r = 0; r < reps; ++r) {
        cudaMemsetAsync(scratch, r & 0xff, FLUSH, s);
        cudaEventRecord(e0, s);
        one.run(s);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        ms += t;
      }
    } else {
      cudaEventRecord(e0, s);
      for (int r = 0; r < reps; ++r) one.run(s);
      cudaEventRecord(e1, s);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    if (k < cap) {
      ms_out[k] = ms / reps;
      kind_out[k] = (int)o.kind;
      const bool is_lin = o.kind == Op::LIN_F, is_mlp = o.kind == Op::MLP_F || o.kind == Op::MLP_B || is_lin;
      if (flop_out) flop_out[k] = o.kind == Op::GEMM ? 2.0 * o.gemm.a.M * o.gemm.a.N * o.gemm.a.K * o.gemm.a.batch
                                  : is_lin ? 2.0 * o.mlp.a.rows * o.mlp.a.n_out * o.mlp.D * o.mlp.a.batch
                                  : is_mlp ? 16.0 * o.mlp.a.rows * o.mlp.D * o.mlp.D * o.mlp.a.batch : 0.0;      // two GEMMs of rows x 4D x D
      if (mnk_out && is_mlp) {
        mnk_out[4 * k] = o.mlp.a.rows; mnk_out[4 * k + 1] = is_lin ? o.mlp.a.n_out : 4 * o.mlp.D; mnk_out[4 * k + 2] = o.mlp.D; mnk_out[4 * k + 3] = o.mlp.a.batch;
      } else if (mnk_out) {
        mnk_out[4 * k] = o.kind == Op::GEMM ? o.gemm.a.M : (o.kind == Op::LN_F ? o.lnf.rows : o.kind == Op::LN_B ? o.lnb.rows : 0);
        mnk_out[4 * k + 1] = o.kind == Op::GEMM ? o.gemm.a.N : (o.kind == Op::LN_F ? o.lnf.C : o.kind == Op::LN_B ? o.lnb.C : (o.kind == Op::ATT_F || o.kind == Op::ATT_B) ? o.att.hd : 0);
        mnk_out[4 * k + 2] = o.kind == Op::GEMM ? o.gemm.a.K : 0;
        mnk_out[4 * k + 3] = o.kind == Op::GEMM ? o.gemm.a.batch : (o.kind == Op::LN_F ? o.lnf.batch : o.kind == Op::LN_B ? o.lnb.batch : (o.kind == Op::ATT_F || o.kind == Op::ATT_B) ? o.att.batch : 0);
      }
    }
    ++k;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (scratch) cudaFree(scratch);
  VV_CUDA(cudaGetLastError());
  return k;
}

// ---- kernel-level hooks -----------------------------------------------------------------------
VV_API int vv_test_gemm(const void* A, const void* B, const float* bias, const float* res, float* out_f32, void* out_bf16, void* aux, int M,
                 int N, int K, int batch, int epi, void* stream) {
  GemmArgs g{};
  g.f16 = (epi >> 4) & 1;          // bit 4: operands, 16-bit outputs and the saved pre-activation are fp16
  g.aux_f16 = g.f16;
  epi &= 15;
  g.M = M; g.N = N; g.K = K; g.batch = batch; g.epi = epi;
  g.bias = bias; g.bias_bs = N;
  g.res = res; g.ld_res = N; g.res_bs = (long long)M * N;
  if (epi == EPI_GELU) g.aux_out = (bf16*)aux;
  if (epi == EPI_DGELU) g.aux_in = (const bf16*)aux;
  g.ld_aux = N; g.aux_bs = (long long)M * N;
  g.out_f32 = out_f32; g.ld_f32 = N; g.f32_bs = (long long)M * N;
  g.out_bf16 = (bf16*)out_bf16; g.ld_bf16 = N; g.bf16_bs = (long long)M * N;
  GemmDesc d;
  const char* er = make_gemm_desc(&d, (const bf16*)A, K, (long long)M * K, (const bf16*)B, K, (long long)N * K, g);
  VV_CHECK(!er, "%s", er);
  launch_gemm(d, (cudaStream_t)stream);
  VV_CUDA(cudaGetLastError());
  return 0;
}

// Folded-LayerNorm GEMM hook: out = epi( LN(x) W^T + b ) computed as rstd ((x - shift) W'^T - (mean - shift) s) + c from the centred
// 16-bit copy x - shift (A), the folded operand W' (B), c (bias), s (colsum) and the per-row (mean, M2) partials [batch][parts][M]
// float2 a producer with tile width prod_bn left (prod_bn = 0: one partial over the whole row).  stats_out / out_f32: when given,
// the GEMM acts as the statistics PRODUCER for its fp32 output instead (partials [batch][2 * n_tiles][M]; its 16-bit copy is
// centred on `shift`); parts_out[0] = partials per row, parts_out[1] = its tile width.  shift: [batch][M] or null; health: 2 counters.
VV_API int vv_test_gemm_ln(const void* A, const void* B, const float* cbias, const float* colsum, const float* stats, int parts, int C,
                    float eps, void* out_16, void* aux_16, float* out_f32, float* stats_out, int* parts_out, int M, int N, int K,
                    int batch, int epi, const float* shift, int prod_bn, uint32_t* health, const float* res, void* stream) {
  GemmArgs g{};
  g.f16 = (epi >> 4) & 1; g.aux_f16 = g.f16;
  epi &= 15;
  g.M = M; g.N = N; g.K = K; g.batch = batch; g.epi = epi;
  g.bias = cbias; g.bias_bs = N;
  if (stats) {
    g.ln_stats = stats; g.ln_parts = parts; g.ln_stats_bs = (long long)parts * M * 2; g.ln_colsum = colsum;
    g.ln_prod_bn = prod_bn; g.ln_c = C; g.ln_health = health;
    g.ln_inv_c = 1.0f / (float)C; g.ln_eps = eps;
  }
  g.ln_shift = shift;
  g.res = res; g.ld_res = N; g.res_bs = (long long)M * N;
  if (epi == EPI_GELU) g.aux_out = (bf16*)aux_16;
  g.ld_aux = N; g.aux_bs = (long long)M * N;
  g.out_bf16 = (bf16*)out_16; g.ld_bf16 = N; g.bf16_bs = (long long)M * N;
  g.out_f32 = out_f32; g.ld_f32 = N; g.f32_bs = (long long)M * N;
  g.stats_out = stats_out;
  GemmDesc d;
  const char* er = make_gemm_desc(&d, (const bf16*)A, K, (long long)M * K, (const bf16*)B, K, (long long)N * K, g);
  VV_CHECK(!er, "%s", er);
  const int po = 2 * ((N + d.bn - 1) / d.bn);
  d.a.stats_out_bs = (long long)po * M * 2;
  if (parts_out) { parts_out[0] = po; parts_out[1] = d.bn; }
  launch_gemm(d, (cudaStream_t)stream);
  VV_CUDA(cudaGetLastError());
  return 0;
}

// Statistics pass of a stage input (the LayerNorm kernel's statistics-only mode): x (rows, C) fp32 -> centred 16-bit copy, (mean, M2)
// per row (float2) and the row shifts.
VV_API int vv_test_ln_stats(const float* x, void* out_16, float* stats, float* shift, int rows, int C, int f16, void* stream) {
  VV_CHECK(ln_supported(MAP_PLAIN, C), "LayerNorm width %d not instantiated", C);
  LnArgs a{rows, C, 1, MAP_PLAIN, 0, 0, 0.f, x, C, 0, nullptr, nullptr, 0, (bf16*)out_16, C, 0, nullptr, 0, 0, f16, stats, shift};
  launch_ln_fwd(a, (cudaStream_t)stream);
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_debug_gemm_trace(void* dev_ptr) {
  set_gemm_trace((unsigned long long*)dev_ptr);
  return 0;
}
VV_API int vv_debug_gemm_mode(int mode) {
  set_gemm_debug_mode(mode);
  return 0;
}

VV_API int vv_test_layernorm(const float* x, const float* gamma, const float* beta, float* y, const float* dy, float* dx, int rows, int C,
                      float eps, void* stream) {
  VV_CHECK(ln_supported(MAP_PLAIN, C), "LayerNorm width %d not instantiated", C);
  if (y) {
    LnArgs a{rows, C, 1, MAP_PLAIN, 0, 0, eps, x, C, 0, gamma, beta, 0, nullptr, 0, 0, y, C, 0, 0, nullptr, nullptr};
    launch_ln_fwd(a, (cudaStream_t)stream);
  }
  if (dy && dx) {
    LnBwdArgs a{rows, C, 1, MAP_PLAIN, 0, 0, eps, x, C, 0, gamma, 0, dy, C, 0, nullptr, 0, 0, dx, C, 0, nullptr, 0, 0};
    launch_ln_bwd(a, (cudaStream_t)stream);
  }
  VV_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_test_winattn(const void* qkv, const float* relbias, void* out, const void* dout, void* dqkv, int gh, int gw, int heads, int hd,
                    int shift, int f16, void* stream) {
  VV_CHECK(hd == 32 || hd == 192, "head_dim %d not instantiated", hd);
  const long long d = (long long)heads * hd;
  AttnArgs a{gh, gw, heads, hd, shift, 1, (const bf16*)qkv, 3 * d, 0, relbias, 0, (bf16*)out, d, 0, (const bf16*)dout, (bf16*)dqkv, f16};
  if (out) launch_attn_fwd(a, (cudaStream_t)stream);
  if (dout && dqkv) launch_attn_bwd(a, (cudaStream_t)stream);
  VV_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
