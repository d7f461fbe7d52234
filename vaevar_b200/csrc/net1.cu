// The forecast network LGUnet_all_1 (networks/LGUnet_all.py:743-777) as a forward-only launch plan behind its own handle
// (vv_net1_*, include/vaevar.h): the DA cycle applies it once per cycle without differentiating it (da_4dvar.py:1329, 666-681).
//
//   Enc_net / Transformer_Encoder   networks/LGUnet_all.py:553-590, 345-412   patch (3, 2) / stride 2 embedding + APE, three tower
//                                                                             levels (d, 2d, 4d) joined by PatchMerging (:64-98)
//   LG_net                          :653-739                                  pos_embed; first stage = one window over the whole
//                                                                             grid, later stages shifted wh x ww windows
//   Dec_net / Transformer_Decoder   :592-650, 414-477                         concat_back_dim + towers + PatchExpand (:101-118),
//                                                                             ConvTranspose2d (3, 2) / stride 2 head, mean | std
//   Windowattn_block                networks/utils/Blocks.py:103-159          pre-norm, LayerNorm eps 1e-6
//   SD_attn + rope2                 networks/utils/Attention.py:467-664, positional_encodings.py:230-268
//
// The six variable-group towers run as batched launches (batch = group).  Linears are the tcgen05 GEMM of gemm_tcgen05.cuh (fp16
// operands, fp32 accumulation), LayerNorms the two-pass fp32 kernels of kernels.cu (PatchMerging / PatchExpand folded into their
// addressing), attention / RoPE / patch operators the kernels of net1_kernels.cu.  Nothing is stashed: two fp32 residual buffers
// per level, 16-bit scratch sized for the finest level.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "engine.h"

using namespace vv;

#define N1_CUDA(x)                                                                                   \
  do {                                                                                               \
    cudaError_t _e = (x);                                                                            \
    if (_e != cudaSuccess) {                                                                         \
      set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(_e), __FILE__, __LINE__);            \
      return -1;                                                                                     \
    }                                                                                                \
  } while (0)
#define N1_CHECK(cond, ...)      \
  do {                           \
    if (!(cond)) {               \
      set_error(__VA_ARGS__);    \
      return -2;                 \
    }                            \
  } while (0)

namespace {

struct N1Block {
  int d = 0, heads = 0, G = 0;
  bf16 *Wqkv = nullptr, *Wproj = nullptr, *W1 = nullptr, *W2 = nullptr;
  float *bqkv = nullptr, *bproj = nullptr, *b1 = nullptr, *b2 = nullptr, *g1 = nullptr, *be1 = nullptr, *g2 = nullptr, *be2 = nullptr;
};

}  // namespace

struct vv_net1 {
  vv_net1_config c{};
  int G = 0, D = 0, E = 0, H = 0, W = 0, nl = 0;
  int gh[VV_NET1_MAX_LEVELS] = {}, gw[VV_NET1_MAX_LEVELS] = {}, dl[VV_NET1_MAX_LEVELS] = {};
  long long L[VV_NET1_MAX_LEVELS] = {};
  int cin = 0, cout = 0, ckeep = 0;
  std::map<std::string, std::pair<float*, std::vector<int64_t>>> staged;
  std::vector<void*> allocs;
  long long bytes = 0;
  bool finalized = false;
  // packed weights
  std::vector<std::vector<N1Block>> enc, dec;    // enc[level][block]; dec[inx][block], inx = 0 is the coarsest level
  std::vector<N1Block> lg;                       // trunk blocks, all stages in order
  std::vector<int> lg_stage;                     // stage of every trunk block
  std::vector<int> lg_index;                     // index of the block inside its stage
  int *pe_kcnt = nullptr, *pe_cbase = nullptr; float *pe_W = nullptr, *pe_bias = nullptr, *ape = nullptr; int pe_max = 0;
  int *ct_kcnt = nullptr, *ct_cbase = nullptr, *ct_chan = nullptr; float *ct_W = nullptr, *ct_bias = nullptr; int ct_max = 0;
  std::vector<float*> mg_g, mg_b; std::vector<bf16*> Wred;           // PatchMerging into level l (index l, l >= 1)
  float *en_g = nullptr, *en_b = nullptr;                             // Transformer_Encoder.norm
  bf16 *Wep = nullptr, *Wdp = nullptr; float *bep = nullptr, *bdp = nullptr, *pos = nullptr;
  std::vector<bf16*> Wc; std::vector<float*> bc;                      // concat_back_dim[inx]
  std::vector<bf16*> Wex; std::vector<float*> ex_g, ex_b;             // layers_up[inx].upsample
  float *nu_g = nullptr, *nu_b = nullptr;
  std::map<long long, float2*> rope;                                  // key (wh, ww, hd) -> device table
  // plan
  Plan plan;
  bool plan_built = false;
  float *IN = nullptr, *OUT = nullptr;
  float *mean = nullptr, *sigma = nullptr, *inv_sigma = nullptr, *neg_mu_sig = nullptr;
  bool have_consts = false;
  int last_launches = 0;
};

namespace {

template <typename T>
T* n1_alloc(vv_net1* n, size_t count) {
  void* p = nullptr;
  if (count == 0) count = 1;
  if (cudaMalloc(&p, count * sizeof(T)) != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  n->allocs.push_back(p);
  n->bytes += (long long)(count * sizeof(T));
  return static_cast<T*>(p);
}
template <typename T>
T* n1_upload(vv_net1* n, const std::vector<T>& h) {
  T* d = n1_alloc<T>(n, h.size());
  if (d && !h.empty()) cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

struct Reader {
  vv_net1* n; bool ok = true;
  const float* dev(const std::string& name, long long numel) {
    auto it = n->staged.find(name);
    if (it == n->staged.end()) { if (ok) set_error("missing weight '%s'", name.c_str()); ok = false; return nullptr; }
    long long m = 1;
    for (auto v : it->second.second) m *= v;
    if (m != numel) { if (ok) set_error("weight '%s' has %lld elements, expected %lld", name.c_str(), m, numel); ok = false; return nullptr; }
    return it->second.first;
  }
  std::vector<float> host(const std::string& name, long long numel) {
    std::vector<float> h((size_t)numel, 0.f);
    const float* d = dev(name, numel);
    if (d) cudaMemcpy(h.data(), d, (size_t)numel * sizeof(float), cudaMemcpyDeviceToHost);
    return h;
  }
};

void app(std::vector<float>& dst, const std::vector<float>& src) { dst.insert(dst.end(), src.begin(), src.end()); }

// Windowattn_block parameters of `depth` blocks for G identically shaped stacks (Blocks.py:103-140)
int build_blocks(vv_net1* n, Reader& R, std::vector<N1Block>& out, const std::vector<std::string>& prefix, int depth, int d, int heads) {
  const int G = (int)prefix.size();
  out.resize(depth);
  const size_t dd = (size_t)d * d;
  for (int b = 0; b < depth; ++b) {
    N1Block& w = out[b];
    w.d = d; w.heads = heads; w.G = G;
    w.Wqkv = n1_alloc<bf16>(n, G * 3 * dd); w.Wproj = n1_alloc<bf16>(n, G * dd);
    w.W1 = n1_alloc<bf16>(n, G * 4 * dd); w.W2 = n1_alloc<bf16>(n, G * 4 * dd);
    if (!w.Wqkv || !w.Wproj || !w.W1 || !w.W2) return -1;
    std::vector<float> bqkv, bproj, b1, b2, g1, be1, g2, be2;
    for (int g = 0; g < G; ++g) {
      const std::string p = prefix[g] + ".blocks." + std::to_string(b);
      const float* s;
      if ((s = R.dev(p + ".attn.qkv.weight", 3 * dd))) launch_pack_w(w.Wqkv + g * 3 * dd, nullptr, s, 3 * d, d, 1);
      if ((s = R.dev(p + ".attn.proj.weight", dd))) launch_pack_w(w.Wproj + g * dd, nullptr, s, d, d, 1);
      if ((s = R.dev(p + ".mlp.fc1.weight", 4 * dd))) launch_pack_w(w.W1 + g * 4 * dd, nullptr, s, 4 * d, d, 1);
      if ((s = R.dev(p + ".mlp.fc2.weight", 4 * dd))) launch_pack_w(w.W2 + g * 4 * dd, nullptr, s, d, 4 * d, 1);
      app(bqkv, R.host(p + ".attn.qkv.bias", 3 * d)); app(bproj, R.host(p + ".attn.proj.bias", d));
      app(b1, R.host(p + ".mlp.fc1.bias", 4 * d)); app(b2, R.host(p + ".mlp.fc2.bias", d));
      app(g1, R.host(p + ".norm.weight", d)); app(be1, R.host(p + ".norm.bias", d));
      app(g2, R.host(p + ".norm2.weight", d)); app(be2, R.host(p + ".norm2.bias", d));
      if (!R.ok) return -2;
    }
    w.bqkv = n1_upload(n, bqkv); w.bproj = n1_upload(n, bproj); w.b1 = n1_upload(n, b1); w.b2 = n1_upload(n, b2);
    w.g1 = n1_upload(n, g1); w.be1 = n1_upload(n, be1); w.g2 = n1_upload(n, g2); w.be2 = n1_upload(n, be2);
    if (!w.bqkv || !w.bproj || !w.b1 || !w.b2 || !w.g1 || !w.be1 || !w.g2 || !w.be2) return -1;
  }
  return 0;
}

// rope2 table of a (wh, ww) window for head width hd (positional_encodings.py:231-252): [wh * ww][hd / 2] (cos, sin); pairs
// j < hd / 4 turn with the row, the others with the column.
float2* rope_table(vv_net1* n, int wh, int ww, int hd) {
  const long long key = ((long long)wh << 40) | ((long long)ww << 16) | hd;
  auto it = n->rope.find(key);
  if (it != n->rope.end()) return it->second;
  const int half = hd / 2, d1 = half / 2, d2 = half - d1;
  std::vector<float2> t((size_t)wh * ww * half);
  for (int r = 0; r < wh; ++r)
    for (int c = 0; c < ww; ++c)
      for (int j = 0; j < half; ++j) {
        // torch evaluates 10000 ** -(arange / d) and the product with the coordinate in fp32
        float ang;
        if (j < d1) ang = (float)r * powf(10000.f, -((float)j / (float)d1));
        else ang = (float)c * powf(10000.f, -((float)(j - d1) / (float)d2));
        t[((size_t)r * ww + c) * half + j] = make_float2((float)cos((double)ang), (float)sin((double)ang));
      }
  float2* d = n1_upload(n, t);
  n->rope[key] = d;
  return d;
}

struct Scratch {
  bf16 *h, *qkv, *ao, *a;       // LayerNorm output, qkv, attention output, MLP hidden (sized for the widest stage)
  bf16* vt; long long vt_ld;    // V^T of the whole-grid trunk stage for the tcgen05 attention ([embed_dim][round_up(tokens, 64)])
};

struct N1Builder {
  vv_net1* n; Scratch t; const char* err = nullptr;

  void gemm(const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb, long long b_bs, const GemmArgs& g) {
    Op o{}; o.kind = Op::GEMM;
    const char* er = make_gemm_desc(&o.gemm, A, lda, a_bs, B, ldb, b_bs, g);
    if (er && !err) err = er;
    n->plan.ops.push_back(o);
  }
  static GemmArgs ga(long long M, int N, int K, int batch) {
    GemmArgs g{}; g.M = (int)M; g.N = N; g.K = K; g.batch = batch; g.epi = EPI_LINEAR; g.f16 = 1; g.aux_f16 = 1; return g;
  }
  void ln(long long rows, int C, int batch, int map, int gh, int gw, const float* x, long long ld_x, long long x_bs, const float* gamma,
          const float* beta, bf16* ob, long long ld_ob, long long ob_bs, float* of, long long ld_of, long long of_bs) {
    if (!ln_supported(map, C) && !err) err = "LayerNorm width not instantiated";
    Op o{}; o.kind = Op::LN_F;
    o.lnf = LnArgs{(int)rows, C, batch, map, gh, gw, 1e-6f, x, ld_x, x_bs, gamma, beta, (long long)C, ob, ld_ob, ob_bs, of, ld_of, of_bs, 1, nullptr, nullptr};
    n->plan.ops.push_back(o);
  }
  // Windowattn_block.forward, pre-norm (Blocks.py:142-157): x -> x1 = x + attn(norm(x)) -> out = x1 + mlp(norm2(x1)).
  // (wh, ww): window; shifted: roll by half a window and mask the last window row's two latitude bands (Attention.py:551-560, 520-548).
  void block(const N1Block& w, int gh, int gw, int wh, int ww, bool shifted, float* x, float* x1, float* out, bf16* copy16, long long ld_c,
             long long bs_c) {
    const int G = w.G, d = w.d, hd = d / w.heads;
    const long long rows = (long long)gh * gw, rd = rows * d;
    const int sh = shifted ? wh / 2 : 0, sw = shifted ? ww / 2 : 0;
    if (!attn1_supported(hd) && !err) err = "attention head width not instantiated (32, 64, 192)";
    if ((gh % wh || gw % ww) && !err) err = "token grid is not a multiple of the attention window";
    ln(rows, d, G, MAP_PLAIN, gh, gw, x, d, rd, w.g1, w.be1, t.h, d, rd, nullptr, 0, 0);
    GemmArgs g = ga(rows, 3 * d, d, G);
    g.bias = w.bqkv; g.bias_bs = 3 * d; g.out_bf16 = t.qkv; g.ld_bf16 = 3 * d; g.bf16_bs = 3 * rd;
    gemm(t.h, d, rd, w.Wqkv, d, 3LL * d * d, g);
    const float2* rtab = rope_table(n, wh, ww, hd);
    const bool fused_rope = attn1_fuses_rope(wh, ww);             // short windows rotate q, k inside the attention kernel
    Op o{};
    if (!fused_rope) {
      o.kind = Op::ROPE;
      o.rope = RopeArgs{gh, gw, wh, ww, sh, sw, w.heads, hd, G, t.qkv, 3LL * d, 3 * rd, rtab};
      n->plan.ops.push_back(o);
    }
    o = Op{}; o.kind = Op::ATT1;
    const int mask = (sw > 0 && ww != gw) ? 1 : 0;                 // Attention.py:553: no mask when the window spans the whole width
    o.att1 = Attn1Args{gh, gw, wh, ww, sh, sw, w.heads, hd, G, mask, t.qkv, 3LL * d, 3 * rd, t.ao, (long long)d, rd, 1.0f / sqrtf((float)hd),
                       (G == 1 && d == n->E) ? t.vt : nullptr, t.vt_ld, fused_rope ? rtab : nullptr};
    n->plan.ops.push_back(o);
    g = ga(rows, d, d, G);
    g.bias = w.bproj; g.bias_bs = d; g.res = x; g.ld_res = d; g.res_bs = rd; g.out_f32 = x1; g.ld_f32 = d; g.f32_bs = rd;
    gemm(t.ao, d, rd, w.Wproj, d, (long long)d * d, g);
    ln(rows, d, G, MAP_PLAIN, gh, gw, x1, d, rd, w.g2, w.be2, t.h, d, rd, nullptr, 0, 0);
    g = ga(rows, 4 * d, d, G);
    g.epi = EPI_GELU; g.bias = w.b1; g.bias_bs = 4 * d; g.out_bf16 = t.a; g.ld_bf16 = 4 * d; g.bf16_bs = 4 * rd;
    gemm(t.h, d, rd, w.W1, d, 4LL * d * d, g);
    g = ga(rows, d, 4 * d, G);
    g.bias = w.b2; g.bias_bs = d; g.res = x1; g.ld_res = d; g.res_bs = rd; g.out_f32 = out; g.ld_f32 = d; g.f32_bs = rd;
    if (copy16) { g.out_bf16 = copy16; g.ld_bf16 = ld_c; g.bf16_bs = bs_c; }
    gemm(t.a, 4 * d, 4 * rd, w.W2, 4 * d, 4LL * d * d, g);
  }
  // a stage of blocks on buffers xa (input and output) / xb (scratch); the last block may also emit a 16-bit copy
  void stage(const std::vector<N1Block>& ws, int gh, int gw, int wh, int ww, bool alternate, float* xa, float* xb, bf16* copy16, long long ld_c,
             long long bs_c) {
    for (size_t b = 0; b < ws.size(); ++b) {
      const bool last = b + 1 == ws.size();
      block(ws[b], gh, gw, wh, ww, alternate && (b % 2 == 1), xa, xb, xa, last ? copy16 : nullptr, ld_c, bs_c);
    }
  }
};

int net1_init(vv_net1& n, const vv_net1_config& c) {
  n.c = c;
  n.G = c.n_groups; n.D = c.enc_dim; n.E = c.embed_dim; n.H = c.img_h; n.W = c.img_w; n.nl = c.n_levels;
  N1_CHECK(n.G >= 1 && n.G <= VV_MAX_GROUPS, "n_groups out of range");
  N1_CHECK(n.nl >= 2 && n.nl <= VV_NET1_MAX_LEVELS, "n_levels out of range");
  N1_CHECK(c.n_lg >= 1 && c.n_lg <= VV_MAX_LG, "n_lg out of range");
  N1_CHECK(c.win_h >= 1 && c.win_w >= 1, "bad window");
  N1_CHECK(n.H >= 3 && n.W >= 2, "image too small for the (3, 2) patch");
  n.gh[0] = (n.H - 3) / 2 + 1; n.gw[0] = (n.W - 2) / 2 + 1;          // PatchEmbed, kernel (3, 2), stride 2 (LGUnet_all.py:14-50)
  N1_CHECK(2 * (n.gh[0] - 1) + 3 == n.H && 2 * (n.gw[0] - 1) + 2 == n.W, "img_size %dx%d is not covered by the (3, 2) / stride-2 patches", n.H, n.W);
  n.dl[0] = n.D; n.L[0] = (long long)n.gh[0] * n.gw[0];
  for (int l = 1; l < n.nl; ++l) {
    N1_CHECK(n.gh[l - 1] % 2 == 0 && n.gw[l - 1] % 2 == 0, "token grid of level %d is odd", l - 1);
    n.gh[l] = n.gh[l - 1] / 2; n.gw[l] = n.gw[l - 1] / 2; n.dl[l] = 2 * n.dl[l - 1]; n.L[l] = (long long)n.gh[l] * n.gw[l];
  }
  for (int l = 0; l < n.nl; ++l) {
    N1_CHECK(n.gh[l] % c.win_h == 0 && n.gw[l] % c.win_w == 0, "token grid %dx%d of level %d is not a multiple of the %dx%d window", n.gh[l],
             n.gw[l], l, c.win_h, c.win_w);
    N1_CHECK(c.enc_heads[l] >= 1 && n.dl[l] % c.enc_heads[l] == 0 && attn1_supported(n.dl[l] / c.enc_heads[l]),
             "tower level %d: head width %d not built (32, 64, 192)", l, c.enc_heads[l] ? n.dl[l] / c.enc_heads[l] : 0);
    N1_CHECK(ln_supported(MAP_PLAIN, n.dl[l]), "LayerNorm width %d not instantiated", n.dl[l]);
    if (l > 0) N1_CHECK(ln_supported(MAP_MERGE, 4 * n.dl[l - 1]) && ln_supported(MAP_EXPAND, n.dl[l - 1]), "PatchMerging / PatchExpand LayerNorm widths of level %d not instantiated", l);
    N1_CHECK(c.enc_depth[l] >= 1, "tower depth must be >= 1");
  }
  for (int s = 0; s < c.n_lg; ++s)
    N1_CHECK(c.lg_heads[s] >= 1 && n.E % c.lg_heads[s] == 0 && attn1_supported(n.E / c.lg_heads[s]), "trunk head width not built (32, 64, 192)");
  N1_CHECK(ln_supported(MAP_PLAIN, n.E), "LayerNorm width %d not instantiated", n.E);
  N1_CHECK(patch32_supported(n.D), "enc_dim %d not supported by the patch kernels (32, 64, 96, 128)", n.D);
  n.cin = n.cout = 0;
  for (int g = 0; g < n.G; ++g) {
    N1_CHECK(c.in_chans[g] >= 1 && c.out_chans[g] >= 2 && c.out_chans[g] % 2 == 0 && c.in_chans[g] <= 32 && c.out_chans[g] <= 64, "bad channel lists");
    n.cin += c.in_chans[g]; n.cout += c.out_chans[g];
  }
  n.ckeep = c.keep_out > 0 ? c.keep_out : n.cout;
  N1_CHECK(n.ckeep <= n.cout, "keep_out exceeds the output channels");
  return 0;
}

int net1_finalize(vv_net1* n) {
  Reader R{n};
  const int G = n->G, nl = n->nl, E = n->E, D = n->D;
  auto pre = [&](const char* fmt_a, const std::string& tail) {
    std::vector<std::string> v;
    for (int g = 0; g < G; ++g) v.push_back(std::string(fmt_a) + std::to_string(g) + tail);
    return v;
  };
  int rc;
  n->enc.resize(nl); n->dec.resize(nl);
  for (int l = 0; l < nl; ++l)
    if ((rc = build_blocks(n, R, n->enc[l], pre("enc.enc_list.", ".layers." + std::to_string(l)), n->c.enc_depth[l], n->dl[l], n->c.enc_heads[l]))) return rc;
  for (int inx = 0; inx < nl; ++inx) {
    const int l = nl - 1 - inx;
    if ((rc = build_blocks(n, R, n->dec[inx], pre("dec.dec_list.", ".layers_up." + std::to_string(inx)), n->c.enc_depth[l], n->dl[l], n->c.enc_heads[l]))) return rc;
  }
  n->lg.clear(); n->lg_stage.clear(); n->lg_index.clear();
  for (int s = 0; s < n->c.n_lg; ++s) {
    std::vector<N1Block> tmp;
    if ((rc = build_blocks(n, R, tmp, {"net.layers." + std::to_string(s)}, n->c.lg_depth[s], E, n->c.lg_heads[s]))) return rc;
    for (size_t b = 0; b < tmp.size(); ++b) { n->lg.push_back(tmp[b]); n->lg_stage.push_back(s); n->lg_index.push_back((int)b); }
  }
  auto stack = [&](const std::vector<std::string>& names, long long numel) {
    std::vector<float> all;
    for (auto& nm : names) app(all, R.host(nm, numel));
    return n1_upload(n, all);
  };
  // ---- patch embedding (Conv2d kernel (3, 2), stride 2) + absolute position embedding ----
  {
    std::vector<int> kcnt, cbase;
    std::vector<float> Wp, bias;
    int c0 = 0;
    n->ape = n1_alloc<float>(n, (size_t)G * n->L[0] * D);
    if (!n->ape) return -1;
    for (int g = 0; g < G; ++g) {
      const int cg = n->c.in_chans[g];
      const std::string p = "enc.enc_list." + std::to_string(g);
      std::vector<float> w = R.host(p + ".patch_embed.proj.weight", (long long)D * cg * 6);
      app(bias, R.host(p + ".patch_embed.proj.bias", D));
      kcnt.push_back(cg); cbase.push_back(c0);
      for (int ci = 0; ci < cg; ++ci)
        for (int k = 0; k < 6; ++k)
          for (int c = 0; c < D; ++c) Wp.push_back(w[((size_t)c * cg + ci) * 6 + k]);
      const float* ape = R.dev(p + ".absolute_pos_embed", n->L[0] * D);
      if (ape) cudaMemcpy(n->ape + (size_t)g * n->L[0] * D, ape, (size_t)n->L[0] * D * sizeof(float), cudaMemcpyDeviceToDevice);
      c0 += cg;
    }
    n->pe_kcnt = n1_upload(n, kcnt); n->pe_cbase = n1_upload(n, cbase); n->pe_W = n1_upload(n, Wp); n->pe_bias = n1_upload(n, bias);
    n->pe_max = *std::max_element(kcnt.begin(), kcnt.end());
  }
  // ---- ConvTranspose2d head (kernel (3, 2), stride 2) with the mean | std channel shuffle (LGUnet_all.py:640-650) ----
  {
    std::vector<int> kcnt, cbase, chan;
    std::vector<float> Wt, bias;
    int mean_total = 0;
    for (int g = 0; g < G; ++g) mean_total += n->c.out_chans[g] / 2;
    int mean_off = 0, std_off = 0, slots = 0;
    for (int g = 0; g < G; ++g) {
      const int cg = n->c.out_chans[g], half = cg / 2;
      std::vector<float> w = R.host("dec.final_proj_list." + std::to_string(g) + ".weight", (long long)D * cg * 6);
      std::vector<float> b = R.host("dec.final_proj_list." + std::to_string(g) + ".bias", cg);
      cbase.push_back(slots);
      int cnt = 0;
      for (int k = 0; k < cg; ++k) {
        const int oc = k < half ? mean_off + k : mean_total + std_off + (k - half);
        if (oc >= n->ckeep) continue;
        chan.push_back(oc);
        for (int kk = 0; kk < 6; ++kk)
          for (int c = 0; c < D; ++c) Wt.push_back(w[((size_t)c * cg + k) * 6 + kk]);
        bias.push_back(b[k]);
        ++cnt; ++slots;
      }
      kcnt.push_back(cnt);
      mean_off += half; std_off += cg - half;
    }
    N1_CHECK(slots == n->ckeep, "keep_out=%d does not align with the mean | std channel layout (%d slots)", n->ckeep, slots);
    n->ct_kcnt = n1_upload(n, kcnt); n->ct_cbase = n1_upload(n, cbase); n->ct_chan = n1_upload(n, chan);
    n->ct_W = n1_upload(n, Wt); n->ct_bias = n1_upload(n, bias);
    n->ct_max = *std::max_element(kcnt.begin(), kcnt.end());
  }
  // ---- seams between the stages ----
  n->mg_g.assign(nl, nullptr); n->mg_b.assign(nl, nullptr); n->Wred.assign(nl, nullptr);
  for (int l = 1; l < nl; ++l) {
    const int dp = n->dl[l - 1];
    const std::string tail = ".layers." + std::to_string(l) + ".downsample";
    n->mg_g[l] = stack(pre("enc.enc_list.", tail + ".norm.weight"), 4 * dp);
    n->mg_b[l] = stack(pre("enc.enc_list.", tail + ".norm.bias"), 4 * dp);
    n->Wred[l] = n1_alloc<bf16>(n, (size_t)G * 8 * dp * dp);
    if (!n->Wred[l]) return -1;
    for (int g = 0; g < G; ++g)
      if (const float* s = R.dev("enc.enc_list." + std::to_string(g) + tail + ".reduction.weight", 8LL * dp * dp))
        launch_pack_w(n->Wred[l] + (size_t)g * 8 * dp * dp, nullptr, s, 2 * dp, 4 * dp, 1);
  }
  const int dt = n->dl[nl - 1];
  n->en_g = stack(pre("enc.enc_list.", ".norm.weight"), dt);
  n->en_b = stack(pre("enc.enc_list.", ".norm.bias"), dt);
  const size_t EP = (size_t)E * G * dt;
  n->Wep = n1_alloc<bf16>(n, EP); n->Wdp = n1_alloc<bf16>(n, EP);
  if (!n->Wep || !n->Wdp) return -1;
  if (const float* s = R.dev("enc.proj.weight", EP)) launch_pack_w(n->Wep, nullptr, s, E, G * dt, 1);
  if (const float* s = R.dev("dec.proj.weight", EP)) launch_pack_w(n->Wdp, nullptr, s, G * dt, E, 1);
  n->bep = n1_upload(n, R.host("enc.proj.bias", E));
  n->bdp = n1_upload(n, R.host("dec.proj.bias", (long long)G * dt));
  n->pos = n1_alloc<float>(n, (size_t)n->L[nl - 1] * E);
  if (!n->pos) return -1;
  if (const float* p = R.dev("net.pos_embed", n->L[nl - 1] * E)) cudaMemcpy(n->pos, p, (size_t)n->L[nl - 1] * E * sizeof(float), cudaMemcpyDeviceToDevice);
  n->Wc.assign(nl, nullptr); n->bc.assign(nl, nullptr); n->Wex.assign(nl, nullptr); n->ex_g.assign(nl, nullptr); n->ex_b.assign(nl, nullptr);
  for (int inx = 0; inx < nl; ++inx) {
    const int d = n->dl[nl - 1 - inx];
    const std::string cb = ".concat_back_dim." + std::to_string(inx);
    n->bc[inx] = stack(pre("dec.dec_list.", cb + ".bias"), d);
    n->Wc[inx] = n1_alloc<bf16>(n, (size_t)G * 2 * d * d);
    if (!n->Wc[inx]) return -1;
    for (int g = 0; g < G; ++g)
      if (const float* s = R.dev("dec.dec_list." + std::to_string(g) + cb + ".weight", 2LL * d * d))
        launch_pack_w(n->Wc[inx] + (size_t)g * 2 * d * d, nullptr, s, d, 2 * d, 1);
    if (inx + 1 < nl) {
      const std::string up = ".layers_up." + std::to_string(inx) + ".upsample";
      n->ex_g[inx] = stack(pre("dec.dec_list.", up + ".norm.weight"), d / 2);
      n->ex_b[inx] = stack(pre("dec.dec_list.", up + ".norm.bias"), d / 2);
      n->Wex[inx] = n1_alloc<bf16>(n, (size_t)G * 2 * d * d);
      if (!n->Wex[inx]) return -1;
      for (int g = 0; g < G; ++g)
        if (const float* s = R.dev("dec.dec_list." + std::to_string(g) + up + ".expand.weight", 2LL * d * d))
          launch_pack_w(n->Wex[inx] + (size_t)g * 2 * d * d, nullptr, s, 2 * d, d, 1);
    }
  }
  n->nu_g = stack(pre("dec.dec_list.", ".norm_up.weight"), D);
  n->nu_b = stack(pre("dec.dec_list.", ".norm_up.bias"), D);
  if (!R.ok) return -2;
  N1_CUDA(cudaDeviceSynchronize());
  for (auto& kv : n->staged) cudaFree(kv.second.first);
  n->staged.clear();
  n->finalized = true;
  return 0;
}

int net1_build_plan(vv_net1* n) {
  if (n->plan_built) return 0;
  N1_CHECK(n->finalized, "vv_net1_finalize has not been called");
  const int G = n->G, nl = n->nl, E = n->E, D = n->D, top = nl - 1;
  const long long HW = (long long)n->H * n->W;
  // ---- buffers ----
  size_t m_rd = (size_t)n->L[top] * E;
  for (int l = 0; l < nl; ++l) m_rd = std::max<size_t>(m_rd, (size_t)G * n->L[l] * n->dl[l]);
  Scratch t{};
  t.h = n1_alloc<bf16>(n, m_rd); t.qkv = n1_alloc<bf16>(n, 3 * m_rd); t.ao = n1_alloc<bf16>(n, m_rd); t.a = n1_alloc<bf16>(n, 4 * m_rd);
  t.vt_ld = (n->L[top] + 63) / 64 * 64;
  t.vt = n1_alloc<bf16>(n, (size_t)E * t.vt_ld);
  if (!t.vt) return -1;
  cudaMemset(t.vt, 0, (size_t)E * t.vt_ld * sizeof(bf16));
  std::vector<float*> XA(nl), XB(nl); std::vector<bf16*> CAT(nl);
  for (int l = 0; l < nl; ++l) {
    const size_t sz = (size_t)G * n->L[l] * n->dl[l];
    XA[l] = n1_alloc<float>(n, sz); XB[l] = n1_alloc<float>(n, sz); CAT[l] = n1_alloc<bf16>(n, 2 * sz);
    if (!XA[l] || !XB[l] || !CAT[l]) return -1;
  }
  float* TA = n1_alloc<float>(n, (size_t)n->L[top] * E); float* TBf = n1_alloc<float>(n, (size_t)n->L[top] * E);
  bf16* T16 = n1_alloc<bf16>(n, (size_t)n->L[top] * E);
  bf16* EPIN = n1_alloc<bf16>(n, (size_t)G * n->L[top] * n->dl[top]);
  size_t m_mb = 0, m_ex = 0;                         // PatchMerging operand / PatchExpand output of the largest level pair
  for (int l = 1; l < nl; ++l) { m_mb = std::max<size_t>(m_mb, (size_t)G * n->L[l] * 4 * n->dl[l - 1]); m_ex = std::max<size_t>(m_ex, (size_t)G * n->L[l] * 2 * n->dl[l]); }
  bf16* MB = n1_alloc<bf16>(n, m_mb); float* EX = n1_alloc<float>(n, m_ex);
  bf16* U16 = n1_alloc<bf16>(n, m_rd);
  float* NU = n1_alloc<float>(n, (size_t)G * n->L[0] * D);
  n->IN = n1_alloc<float>(n, (size_t)std::max(n->cin, n->ckeep) * HW); n->OUT = n1_alloc<float>(n, (size_t)n->ckeep * HW);
  if (!t.h || !t.qkv || !t.ao || !t.a || !TA || !TBf || !T16 || !EPIN || !MB || !EX || !U16 || !NU || !n->IN || !n->OUT) return -1;

  N1Builder B{n, t};
  const int wh = n->c.win_h, ww = n->c.win_w;
  // ---- Enc_net (LGUnet_all.py:575-590) ----
  Op o{}; o.kind = Op::PE32;
  o.pe32 = Patch32Args{n->H, n->W, n->gh[0], n->gw[0], G, D, n->pe_kcnt, n->pe_cbase, n->pe_W, n->pe_bias, n->ape, n->IN, XA[0], n->pe_max};
  n->plan.ops.push_back(o);
  for (int l = 0; l < nl; ++l) {
    const int d = n->dl[l];
    const long long Ll = n->L[l];
    if (l > 0) {                                     // PatchMerging (:80-98): gather + LayerNorm(4 d') + Linear(4 d' -> 2 d')
      const int dp = n->dl[l - 1];
      B.ln(Ll, 4 * dp, G, MAP_MERGE, n->gh[l - 1], n->gw[l - 1], XA[l - 1], dp, n->L[l - 1] * dp, n->mg_g[l], n->mg_b[l], MB, 4 * dp, Ll * 4 * dp, nullptr, 0, 0);
      GemmArgs g = N1Builder::ga(Ll, d, 4 * dp, G);
      g.out_f32 = XA[l]; g.ld_f32 = d; g.f32_bs = Ll * d;
      B.gemm(MB, 4 * dp, Ll * 4 * dp, n->Wred[l], 4 * dp, 8LL * dp * dp, g);
    }
    // the stage output is the skip connection: its 16-bit copy goes to the second half of the decoder's concat buffer
    B.stage(n->enc[l], n->gh[l], n->gw[l], wh, ww, true, XA[l], XB[l], CAT[l] + d, 2 * d, Ll * 2 * d);
  }
  {
    const int d = n->dl[top];
    const long long Lt = n->L[top];
    B.ln(Lt, d, G, MAP_PLAIN, n->gh[top], n->gw[top], XA[top], d, Lt * d, n->en_g, n->en_b, EPIN, (long long)G * d, d, nullptr, 0, 0);
    GemmArgs g = N1Builder::ga(Lt, E, G * d, 1);     // Enc_net.proj over the concatenated towers, + pos_embed (:722-724)
    g.bias = n->bep; g.res = n->pos; g.ld_res = E; g.out_f32 = TA; g.ld_f32 = E;
    B.gemm(EPIN, (long long)G * d, 0, n->Wep, (long long)G * d, 0, g);
  }
  // ---- LG_net (:725-739): stage 0 attends over the whole grid (unshifted), the others over shifted windows ----
  for (size_t b = 0; b < n->lg.size(); ++b) {
    const bool last = b + 1 == n->lg.size();
    const bool whole = n->lg_stage[b] == 0;
    B.block(n->lg[b], n->gh[top], n->gw[top], whole ? n->gh[top] : wh, whole ? n->gw[top] : ww, !whole && (n->lg_index[b] % 2 == 1), TA, TBf, TA,
            last ? T16 : nullptr, E, 0);
  }
  // ---- Dec_net (:625-650) ----
  {
    const int d = n->dl[top];
    const long long Lt = n->L[top];
    GemmArgs g = N1Builder::ga(Lt, G * d, E, 1);      // Dec_net.proj, split over the towers into the first half of CAT[top]
    g.bias = n->bdp; g.out_bf16 = CAT[top]; g.ld_bf16 = 2 * d; g.split_n = d; g.split_stride = Lt * 2 * d;
    if (d % GEMM_EC && !B.err) B.err = "top tower width must be a multiple of 32";
    B.gemm(T16, E, 0, n->Wdp, E, 0, g);
  }
  for (int inx = 0; inx < nl; ++inx) {
    const int l = top - inx, d = n->dl[l];
    const long long Ll = n->L[l];
    GemmArgs g = N1Builder::ga(Ll, d, 2 * d, G);      // concat_back_dim[inx] on [x | skip] (:468-470)
    g.bias = n->bc[inx]; g.bias_bs = d; g.out_f32 = XA[l]; g.ld_f32 = d; g.f32_bs = Ll * d;
    B.gemm(CAT[l], 2 * d, Ll * 2 * d, n->Wc[inx], 2 * d, 2LL * d * d, g);
    const bool more = inx + 1 < nl;
    B.stage(n->dec[inx], n->gh[l], n->gw[l], wh, ww, true, XA[l], XB[l], more ? U16 : nullptr, d, Ll * d);
    if (more) {                                       // PatchExpand (:107-118): Linear(d -> 2 d), pixel shuffle, LayerNorm(d / 2)
      g = N1Builder::ga(Ll, 2 * d, d, G);
      g.out_f32 = EX; g.ld_f32 = 2 * d; g.f32_bs = Ll * 2 * d;
      B.gemm(U16, d, Ll * d, n->Wex[inx], d, 2LL * d * d, g);
      const int dn = d / 2;
      B.ln(n->L[l - 1], dn, G, MAP_EXPAND, n->gh[l - 1], n->gw[l - 1], EX, 2 * d, Ll * 2 * d, n->ex_g[inx], n->ex_b[inx], CAT[l - 1], 2 * dn,
           n->L[l - 1] * 2 * dn, nullptr, 0, 0);
    }
  }
  B.ln(n->L[0], D, G, MAP_PLAIN, n->gh[0], n->gw[0], XA[0], D, n->L[0] * D, n->nu_g, n->nu_b, nullptr, 0, 0, NU, D, n->L[0] * D);
  o = Op{}; o.kind = Op::CT32;
  o.ct32 = ConvT32Args{n->H, n->W, n->gh[0], n->gw[0], G, D, n->ct_kcnt, n->ct_cbase, n->ct_chan, n->ct_W, n->ct_bias, NU, n->OUT, n->ct_max};
  n->plan.ops.push_back(o);
  N1_CHECK(!B.err, "plan construction failed: %s", B.err);
  n->plan.label = "LGUnet_all_1 forward";
  n->plan_built = true;
  return 0;
}

}  // namespace

extern "C" {

VV_API int vv_net1_create(const vv_net1_config* cfg, vv_net1** out) {
  N1_CHECK(cfg && out, "null argument");
  int dev = 0;
  N1_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  N1_CUDA(cudaGetDeviceProperties(&prop, dev));
  N1_CHECK(prop.major == 10, "vaevar_b200 needs an sm_100 GPU (found sm_%d%d); there is no fallback path", prop.major, prop.minor);
  vv_net1* n = new vv_net1();
  const int rc = net1_init(*n, *cfg);
  if (rc) { delete n; return rc; }
  *out = n;
  return 0;
}

VV_API void vv_net1_destroy(vv_net1* n) {
  if (!n) return;
  cudaDeviceSynchronize();
  for (auto& kv : n->staged) cudaFree(kv.second.first);
  for (void* p : n->allocs) cudaFree(p);
  delete n;
}

VV_API int vv_net1_set_weight(vv_net1* n, const char* name, const float* data_dev, const int64_t* shape, int ndim) {
  N1_CHECK(n && name && data_dev && shape && ndim >= 0, "bad argument");
  N1_CHECK(!n->finalized, "weights already finalized");
  size_t numel = 1;
  std::vector<int64_t> shp(shape, shape + ndim);
  for (auto v : shp) numel *= (size_t)v;
  float* d = nullptr;
  N1_CUDA(cudaMalloc(&d, std::max<size_t>(numel, 1) * sizeof(float)));
  N1_CUDA(cudaMemcpy(d, data_dev, numel * sizeof(float), cudaMemcpyDeviceToDevice));
  auto it = n->staged.find(name);
  if (it != n->staged.end()) cudaFree(it->second.first);
  n->staged[name] = {d, shp};
  return 0;
}

VV_API int vv_net1_finalize(vv_net1* n) {
  N1_CHECK(n, "null handle");
  N1_CHECK(!n->finalized, "weights already finalized");
  return net1_finalize(n);
}

VV_API int vv_net1_forward(vv_net1* n, const float* in_dev, float* out_dev, void* stream) {
  N1_CHECK(n && in_dev && out_dev, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = net1_build_plan(n);
  if (rc) return rc;
  const size_t HW = (size_t)n->H * n->W;
  N1_CUDA(cudaMemcpyAsync(n->IN, in_dev, (size_t)n->cin * HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  n->last_launches = n->plan.run(s);
  N1_CUDA(cudaMemcpyAsync(out_dev, n->OUT, (size_t)n->ckeep * HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  N1_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_net1_set_constants(vv_net1* n, const float* mean, const float* std) {
  N1_CHECK(n && mean && std, "null argument");
  N1_CHECK(n->cin == n->ckeep, "integrate needs a network that maps the state onto itself (keep_out == sum(in_chans))");
  const int C = n->cin;
  std::vector<float> m(mean, mean + C), sg(std, std + C), is(C), nm(C);
  for (int c = 0; c < C; ++c) { is[c] = 1.0f / sg[c]; nm[c] = -m[c] / sg[c]; }
  if (!n->have_consts) {
    n->mean = n1_alloc<float>(n, C); n->sigma = n1_alloc<float>(n, C); n->inv_sigma = n1_alloc<float>(n, C); n->neg_mu_sig = n1_alloc<float>(n, C);
    N1_CHECK(n->mean && n->sigma && n->inv_sigma && n->neg_mu_sig, "out of memory for the constants");
  }
  const size_t nb = (size_t)C * sizeof(float);
  N1_CUDA(cudaDeviceSynchronize());
  N1_CUDA(cudaMemcpy(n->mean, m.data(), nb, cudaMemcpyHostToDevice)); N1_CUDA(cudaMemcpy(n->sigma, sg.data(), nb, cudaMemcpyHostToDevice));
  N1_CUDA(cudaMemcpy(n->inv_sigma, is.data(), nb, cudaMemcpyHostToDevice)); N1_CUDA(cudaMemcpy(n->neg_mu_sig, nm.data(), nb, cudaMemcpyHostToDevice));
  n->have_consts = true;
  return 0;
}

VV_API int vv_net1_integrate(vv_net1* n, const float* x_in_dev, float* x_out_dev, int steps, void* stream) {
  N1_CHECK(n && x_in_dev && x_out_dev && steps >= 1, "bad argument");
  N1_CHECK(n->have_consts, "vv_net1_set_constants has not been called");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = net1_build_plan(n);
  if (rc) return rc;
  const int C = n->cin;
  const long long HW = (long long)n->H * n->W;
  // normalise (da_4dvar.py:667), apply the model `steps` times keeping the mean channels (:673-676), de-normalise (:681)
  launch_chan_affine(n->IN, x_in_dev, n->inv_sigma, nullptr, nullptr, n->neg_mu_sig, C, HW, s);
  int launches = 2;
  for (int k = 0; k < steps; ++k) {
    launches += n->plan.run(s);
    if (k + 1 < steps) N1_CUDA(cudaMemcpyAsync(n->IN, n->OUT, (size_t)C * HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  launch_chan_affine(x_out_dev, n->OUT, n->sigma, nullptr, nullptr, n->mean, C, HW, s);
  n->last_launches = launches;
  N1_CUDA(cudaGetLastError());
  return 0;
}

VV_API int vv_net1_last_launch_count(vv_net1* n) { return n ? n->last_launches : 0; }

// Steady-state time of every launch of the forward plan (each op `reps` times between CUDA events, on whatever the buffers hold
// after one full application).  kind: Op::Kind (0 GEMM, 1 LayerNorm, 7 rope2, 8 SD_attn, 9 patch embed, 10 conv-transpose head);
// flop_out: 2 M N K batch for GEMMs, 4 N^2 hd per (window, head) for attention; mnk_out: 4 ints per op.  Returns the op count.
VV_API int vv_net1_profile_ops(vv_net1* n, int reps, float* ms_out, int* kind_out, double* flop_out, int* mnk_out, int cap) {
  N1_CHECK(n && ms_out && kind_out && reps >= 1, "bad argument");
  int rc = net1_build_plan(n);
  if (rc) return rc;
  cudaStream_t s = nullptr;
  N1_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  N1_CUDA(cudaEventCreate(&e0)); N1_CUDA(cudaEventCreate(&e1));
  n->plan.run(s);
  int k = 0;
  for (const Op& o : n->plan.ops) {
    Plan one; one.ops.push_back(o);
    one.run(s);
    cudaEventRecord(e0, s);
    for (int r = 0; r < reps; ++r) one.run(s);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (k < cap) {
      ms_out[k] = ms / reps;
      kind_out[k] = (int)o.kind;
      double fl = 0.0;
      int m4[4] = {0, 0, 0, 0};
      if (o.kind == Op::GEMM) { fl = 2.0 * o.gemm.a.M * o.gemm.a.N * o.gemm.a.K * o.gemm.a.batch; m4[0] = o.gemm.a.M; m4[1] = o.gemm.a.N; m4[2] = o.gemm.a.K; m4[3] = o.gemm.a.batch; }
      else if (o.kind == Op::ATT1) {
        const double N = (double)o.att1.wh * o.att1.ww, nwin = (double)(o.att1.gh / o.att1.wh) * (o.att1.gw / o.att1.ww);
        fl = 4.0 * N * N * o.att1.hd * o.att1.heads * nwin * o.att1.batch;
        m4[0] = o.att1.gh * o.att1.gw; m4[1] = o.att1.wh * o.att1.ww; m4[2] = o.att1.hd; m4[3] = o.att1.batch * o.att1.heads;
      } else if (o.kind == Op::LN_F) { m4[0] = o.lnf.rows; m4[1] = o.lnf.C; m4[3] = o.lnf.batch; }
      else if (o.kind == Op::ROPE) { m4[0] = o.rope.gh * o.rope.gw; m4[1] = o.rope.wh * o.rope.ww; m4[2] = o.rope.hd; m4[3] = o.rope.batch * o.rope.heads; }
      if (flop_out) flop_out[k] = fl;
      if (mnk_out) for (int j = 0; j < 4; ++j) mnk_out[4 * k + j] = m4[j];
    }
    ++k;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s);
  N1_CUDA(cudaGetLastError());
  return k;
}

VV_API long long vv_net1_device_bytes(vv_net1* n) {
  return n ? n->bytes : 0;
}

// ---- kernel-level hooks (tests) ---------------------------------------------------------------------------------------------
// SD_attn core on a packed fp16 qkv buffer [gh * gw][3 * heads * hd] (Attention.py:560-640): rope2 on q, k in place, then the
// windowed softmax(scale q k^T + mask) v into out [gh * gw][heads * hd].  table_host: [wh * ww][hd / 2] (cos, sin) pairs.
VV_API int vv_test_attn1(void* qkv_dev, void* out_dev, const float* table_dev, int gh, int gw, int wh, int ww, int sh, int sw, int heads, int hd,
                         int mask, int use_tc, void* stream) {
  N1_CHECK(qkv_dev && out_dev && attn1_supported(hd), "bad argument / head width %d not instantiated", hd);
  N1_CHECK(gh % wh == 0 && gw % ww == 0, "grid is not a multiple of the window");
  cudaStream_t s = (cudaStream_t)stream;
  const long long d = (long long)heads * hd;
  const bool fused_rope = table_dev && attn1_fuses_rope(wh, ww);    // then q, k stay unrotated in qkv_dev
  if (table_dev && !fused_rope) {
    RopeArgs r{gh, gw, wh, ww, sh, sw, heads, hd, 1, (bf16*)qkv_dev, 3 * d, 0, (const float2*)table_dev};
    launch_rope(r, s);
  }
  Attn1Args a{gh, gw, wh, ww, sh, sw, heads, hd, 1, mask, (const bf16*)qkv_dev, 3 * d, 0, (bf16*)out_dev, d, 0, 1.0f / sqrtf((float)hd), nullptr, 0,
              fused_rope ? (const float2*)table_dev : nullptr};
  bf16* vt = nullptr;
  if (use_tc) {                                   // scratch for V^T: the whole-grid tcgen05 path (synchronises: test hook only)
    a.vt_ld = ((long long)gh * gw + 63) / 64 * 64;
    N1_CUDA(cudaMalloc(&vt, (size_t)d * a.vt_ld * sizeof(bf16)));
    N1_CUDA(cudaMemsetAsync(vt, 0, (size_t)d * a.vt_ld * sizeof(bf16), s));
    a.vt = vt;
    if (!attn1_tc_eligible(a)) { cudaFree(vt); set_error("shape not eligible for the tcgen05 attention (unshifted whole-grid window, head width 192, >= 512 tokens)"); return -2; }
  }
  launch_attn1(a, s);
  if (vt) { cudaStreamSynchronize(s); cudaFree(vt); }
  N1_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
