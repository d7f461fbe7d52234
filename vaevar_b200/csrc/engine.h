// Engine internals shared between engine.cu (plans, cost/grad) and lbfgs.cu (optimiser).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/vaevar.h"
#include "ops.h"

namespace vv {

void set_error(const char* fmt, ...);
// weight packing shared by engine.cu and net1.cu (definitions in engine.cu)
void launch_pack_w(bf16* dst, bf16* dstT, const float* src, int rows, int cols, int f16);
void launch_fold_ln(bf16* dst, float* colsum, float* cbias, const float* W, const float* gamma, const float* beta, const float* bias,
                    int rows, int cols, int f16);
bool ln_supported(int map, int C);

struct Op {
  enum Kind { GEMM, LN_F, LN_B, ATT_F, ATT_B, P2T, T2P, ROPE, ATT1, PE32, CT32, MLP_F, MLP_B, LIN_F } kind;
  GemmDesc gemm;
  LnArgs lnf;
  LnBwdArgs lnb;
  AttnArgs att;
  PatchArgs patch;
  RopeArgs rope;         // the last four: forecast network LGUnet_all_1 (net1.cu)
  Attn1Args att1;
  Patch32Args pe32;
  ConvT32Args ct32;
  MlpDesc mlp;           // fused tower MLP (mlp_fused.cuh), forward / backward
};
struct Plan {
  std::vector<Op> ops;
  std::string label;               // NVTX range around the plan (VV_NVTX=1): "decoder forward", "flow[3] backward", ...
  int run(cudaStream_t s) const;   // returns number of launches
};
// NVTX ranges (VV_NVTX=1): one per application plan and one per launch, named after the kernel family -- for nsys / ncu --nvtx timelines.
bool nvtx_enabled();
void nvtx_push(const char* name);
void nvtx_pop();

// One stack of G identically shaped Swin blocks (G towers batched, or G = 1 for the trunk).
struct BlockW {
  int d = 0, heads = 0, G = 0;
  bf16 *Wqkv = nullptr, *WqkvT = nullptr, *Wproj = nullptr, *WprojT = nullptr, *W1 = nullptr, *W1T = nullptr, *W2 = nullptr, *W2T = nullptr;
  float *bqkv = nullptr, *bproj = nullptr, *b1 = nullptr, *b2 = nullptr, *g1 = nullptr, *be1 = nullptr, *g2 = nullptr, *be2 = nullptr;
  float* relbias = nullptr;
  float *sqkv = nullptr, *cqkv = nullptr, *s1 = nullptr, *c1 = nullptr;   // LayerNorm folded into qkv / fc1: column sums, folded biases
};
struct BlkStash {
  bf16* qkv; float* x1; bf16* u;
};
struct StageStash {
  std::vector<float*> x;        // x[0] = stage input ... x[depth] = stage output (fp32 residual stream)
  std::vector<BlkStash> b;
};
struct Stash {
  StageStash e0, e1, lg, u0, u1;
  float* EX = nullptr;          // PatchExpand GEMM output, input of its LayerNorm
};

struct PatchPack {
  int* kcnt = nullptr; int* cbase = nullptr; int* chan = nullptr;   // device
  float* Wp = nullptr; float* bias = nullptr;
  int nslots = 0;
  int max_cnt = 0;             // largest kcnt[g]
};

struct Net {
  vv_net_config c{};
  int G = 0, D = 0, E = 0, H = 0, W = 0, h0 = 0, w0 = 0, h1 = 0, w1 = 0, L0 = 0, L1 = 0;
  int cin = 0, cout = 0, ckeep = 0;
  std::map<std::string, std::pair<float*, std::vector<int64_t>>> staged;   // fp32 device copies by reference name
  bool finalized = false;
  // packed weights
  PatchPack embed, fin;
  float* ape = nullptr;           // [G][L0][D]
  std::vector<BlockW> e0, e1, lg, u0, u1;
  float *mg_g = nullptr, *mg_b = nullptr; bf16 *Wred = nullptr, *WredT = nullptr;
  float *en_g = nullptr, *en_b = nullptr;
  bf16 *Wep = nullptr, *WepT = nullptr; float* bep = nullptr; float* pos = nullptr;
  bf16 *Wdp = nullptr, *WdpT = nullptr; float* bdp = nullptr;
  bf16 *Wc0 = nullptr, *Wc0T = nullptr; float* bc0 = nullptr;
  bf16 *Wex = nullptr, *WexT = nullptr; float *ex_g = nullptr, *ex_b = nullptr;
  bf16 *Wc1 = nullptr, *Wc1T = nullptr; float* bc1 = nullptr;
  float *nu_g = nullptr, *nu_b = nullptr;
};

}  // namespace vv

struct vv_engine {
  vv_config cfg{};
  vv::Net net[2];
  int C = 0;                 // state channels (flow in/out kept = decoder out)
  int Zc = 0;                // latent channels
  long long HW = 0;
  std::vector<void*> allocs;
  // constants
  float *mean = nullptr, *sigma = nullptr, *stdTr = nullptr, *inv_sigma = nullptr, *neg_mu_sig = nullptr;
  bool have_consts = false;
  // buffers
  float *Z = nullptr, *DOUT = nullptr, *XN = nullptr, *Gb[2] = {nullptr, nullptr}, *GD = nullptr, *GZ = nullptr, *XB = nullptr;
  // observation data
  int* idx = nullptr; float *yobs = nullptr, *rinv = nullptr, *resid = nullptr;
  int* chunk_counts = nullptr;
  long long n_obs = 0, obs_cap = 0;
  std::vector<long long> obs_off;   // [T+1]
  float obs_coeff = 1.f;
  bool have_case = false;
  unsigned long long generation = 0;   // bumped by every vv_set_case* / vv_set_constants: cached evaluations of an older case are stale
  unsigned int* ln_health = nullptr;       // [2] device counters of the folded LayerNorms (GemmArgs::ln_health)
  double *partials = nullptr, *dots = nullptr, *dot_scratch = nullptr, *Jbuf = nullptr;
  float* met_w = nullptr; double* met_part = nullptr;     // diagnostics scratch (latitude weights, block partials)
  int met_w_cap = 0;
  // native geometry (vv_set_case_native): analysis grid Hh x Wh finer than the network grid; see enqueue_forward
  bool native = false;
  int Hh = 0, Wh = 0;
  float *XF = nullptr, *XBN = nullptr, *TMPF = nullptr, *XBH = nullptr;   // F_t stack, down((xb - mu) / sigma), scratch field, xb on the analysis grid
  long long xbh_cap = 0;
  int *s_row = nullptr, *s_col = nullptr, *s_row_lo = nullptr, *s_col_lo = nullptr;   // S = down o up tables and adjoint ranges
  // channel-mixing observation operator (vv_set_case_obsop): K taps per observation, and the same taps sorted by cell for the adjoint
  int taps = 0;
  long long pair_cap = 0;
  int *tap_ia = nullptr, *pair_cell = nullptr, *pair_src = nullptr;
  float *tap_coef = nullptr, *pair_coef = nullptr;
  // plans: index 0 = decoder application, 1..T-1 = flow applications
  std::vector<vv::Stash> stash;
  std::vector<vv::Plan> fwd, bwd;
  bool plans_built = false;
  // private capturable stream (the caller's stream may be the legacy default stream, which cannot be captured)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // graphs
  cudaGraphExec_t graph_cg = nullptr;
  int eager_runs = 0;
  int last_launches = 0;
};

namespace vv {
// Runs on stream s, which must be capturable (not the legacy default stream): use e->stream.
int engine_cost_grad(vv_engine* e, const float* z, double* Jout, float* grad, cudaStream_t s);
// Order the engine's private stream after `user` (fence_in) / `user` after the private stream (fence_out).
int fence_in(vv_engine* e, cudaStream_t user);
int fence_out(vv_engine* e, cudaStream_t user);
// Debug: per-buffer non-finite counts of the window's trajectory and stashes to stderr (VV_NAN_PROBE=1).
void engine_nan_probe(vv_engine* e, const float* z);
}
