// Host-side launch descriptors for every kernel of the cost-and-gradient path.
// An application plan is a flat vector<Op>; running it is a loop of launches on one stream
// (captured into a CUDA graph by the engine).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "gemm_tcgen05.cuh"
#include "mlp_fused.cuh"

namespace vv {

typedef __nv_bfloat16 bf16;   // also used as the STORAGE type of fp16 activations (the format is a per-launch flag)

// Every kernel of the engine goes through this launcher: programmatic stream serialization lets the prologue of a kernel
// (barrier init, TMEM allocation, descriptor prefetch, index math) overlap the tail of the kernel before it; each kernel
// calls pdl_wait() before its first global access.  VV_NO_PDL=1 turns the attribute off.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- GEMM -----------------------------------------------------------------------------------
struct GemmDesc {
  CUtensorMap tmA, tmB;
  GemmStoreMaps sm;
  GemmArgs a;
  int bn;        // tile = 256 rows (one CTA pair) x bn columns
  const void* b_ptr;            // B operand as one contiguous block (null if it is a strided slice): what a previous GEMM prefetches
  unsigned long long b_bytes;
};
// A: [batch][M][K] 16-bit with row stride lda / batch stride a_bs (elements); B: [batch][N][K] likewise.
const char* make_gemm_desc(GemmDesc* d, const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb,
                           long long b_bs, const GemmArgs& args);
void launch_gemm(const GemmDesc& d, cudaStream_t s);
const char* encode_tma_2d_16(CUtensorMap* tm, const void* base, long long cols, long long rows, long long ld, int box_cols, int box_rows);
const char* encode_tma_3d_16(CUtensorMap* tm, const void* base, long long cols, long long rows, long long batch, long long ld, long long bs,
                             int box_cols, int box_rows);
int num_sms();
void set_gemm_debug_mode(int mode);                 // debug: 1 = MMA only, 2 = TMA feed only (results are garbage)
void set_gemm_trace(unsigned long long* dev_ptr);   // debug: GEMM descriptors built afterwards stamp clocks into dev_ptr (null = off)

// ---- fused tower MLP (mlp_fused.cuh): norm2 + fc1 + GELU + fc2 + residual forward, its input-VJP backward --------------------
struct MlpDesc {
  CUtensorMap tmW1, tmW2, tmA, tmU;
  MlpArgs a;
  int D, bwd, mode;             // mode: MlpMode
  const void* w_ptr;            // W1 | W2 as one contiguous block (null if they are not adjacent): what the GEMM before it prefetches
  unsigned long long w_bytes;
};
bool mlp_fused_supported(int D, int rows);
const char* make_mlp_desc(MlpDesc* d, int D, bool bwd, const __nv_bfloat16* W1, const __nv_bfloat16* W2, const __nv_bfloat16* u,
                          const __nv_bfloat16* dy16, const MlpArgs& args);
const char* make_lin_desc(MlpDesc* d, int D, const __nv_bfloat16* W, __nv_bfloat16* out16, long long ld_out, long long out_bs, const MlpArgs& args);
void launch_mlp(const MlpDesc& d, cudaStream_t s);

// ---- LayerNorm (one warp per row, fp32 statistics) -------------------------------------------
enum RowMap : int {
  MAP_PLAIN = 0,   // row r <-> token r
  MAP_MERGE = 1,   // PatchMerging gather (transformer.py:86-90): row (i,j) of width 4D = tokens (2i+a, 2j+b), chunk order (0,0),(1,0),(0,1),(1,1)
  MAP_EXPAND = 2,  // PatchExpand pixel shuffle (transformer.py:114): row (I,J) of width D = chunk (I%2, J%2) of token (I/2, J/2), width 4D
};
struct LnArgs {
  int rows, C, batch, map, gh, gw;   // gh,gw: the FINE token grid (MERGE: source grid, EXPAND: destination grid)
  float eps;
  const float* x; long long ld_x, x_bs;            // fp32 input (addressed through `map`)
  const float *gamma, *beta; long long gb_bs;
  bf16* out_bf16; long long ld_ob, ob_bs;          // outputs are plain rows
  float* out_f32; long long ld_of, of_bs;
  int out_f16;                                     // 1: the 16-bit output is IEEE fp16 (forward pass), 0: bf16
  float* stats_out;                                // non-null: "statistics only" mode for a LayerNorm folded into the next GEMM --
                                                   // stats_out[batch][rows] float2 = (mean, sum of squared deviations)
  float* shift_out;                                // statistics-only mode: [batch][rows] row shift (= mean); out_bf16 receives x - shift
                                                   // (null: shift 0, the raw row)
};
struct LnBwdArgs {
  int rows, C, batch, map, gh, gw;
  float eps;
  const float* x; long long ld_x, x_bs;            // the forward input (same addressing as forward)
  const float* gamma; long long gb_bs;
  const float* dy; long long ld_dy, dy_bs;         // gradient w.r.t. the LN output, plain rows, fp32 (or, if dy16 is set, bf16 at dy16)
  const float* dres; long long ld_dres, dres_bs;   // optional extra gradient added to dx (addressed like x)
  float* dx; long long ld_dx, dx_bs;               // gradient w.r.t. the LN input (addressed like x)
  bf16* dx_bf16; long long ld_dxb, dxb_bs;         // optional bf16 copy of dx (addressed like x)
  const bf16* dy16;                                // non-null: dy is read from here as bf16 (same ld_dy / dy_bs, in elements)
};
void launch_ln_fwd(const LnArgs& a, cudaStream_t s);
void launch_ln_bwd(const LnBwdArgs& a, cudaStream_t s);

// ---- 4x4-window attention (swinblock.py:133-172, 265-302) ------------------------------------
struct AttnArgs {
  int gh, gw, heads, hd, shift, batch;
  const bf16* qkv; long long ld_qkv, qkv_bs;       // [batch][gh*gw][3*heads*hd], token order = original grid
  const float* relbias; long long relbias_bs;      // [batch][heads][16][16] (table gathered at pack time)
  bf16* out; long long ld_o, o_bs;                 // forward: attention output [batch][tokens][heads*hd]
  const bf16* dout;                                // backward: gradient of `out` (same strides as out), bf16
  bf16* dqkv;                                      // backward: gradient of qkv (same strides as qkv), bf16
  int f16;                                         // 1: qkv (and the forward output) are IEEE fp16, 0: bf16
};
void launch_attn_fwd(const AttnArgs& a, cudaStream_t s);
void launch_attn_bwd(const AttnArgs& a, cudaStream_t s);

// ---- 2x2 / stride-2 patch operators as tiny GEMMs on CUDA cores --------------------------------
// P2T (pixels -> tokens): PatchEmbed forward (transformer.py:35,46 + APE :394) and ConvTranspose2d backward.
// T2P (tokens -> pixels): ConvTranspose2d forward (transformer.py:593-594,606) and PatchEmbed backward.
struct PatchArgs {
  int H, W, G, D;              // image size, groups, token width
  const int* kcnt;             // [G]   channels per group (device)
  const int* cbase;            // [G]   first slot of the group in `chan`
  const int* chan;             // [sum] NCHW channel index of every packed channel (device)
  const float* Wp;             // [sum*4][D] packed weights, row = (slot*4 + p1*2 + p2)
  const float* bias;           // P2T: [G][D] or null; T2P: [sum] or null
  const float* ape;            // P2T only: [G][L0][D] or null
  const float* img_in;         // P2T: NCHW image
  float* tok_out;              // P2T: [G][L0][D]
  const float* tok_in;         // T2P: [G][L0][D]
  float* img_out;              // T2P: NCHW image
  int max_cnt;                 // largest kcnt[g] (host copy; sizes the T2P weight tile); 0 = assume the maximum (32)
};
void launch_p2t(const PatchArgs& a, cudaStream_t s);
void launch_t2p(const PatchArgs& a, cudaStream_t s);

// ---- per-channel affine seams (da_4dvar.py:667,681,1187) ---------------------------------------
// out[c,p] = a[c,p]*sa[c] + (b ? b[c,p]*sb[c] : 0) + t[c]       (any of sa/sb/t may be null -> 1/1/0)
void launch_chan_affine(float* out, const float* a, const float* sa, const float* b, const float* sb, const float* t,
                        int C, long long HW, cudaStream_t s);

// ---- observation operator (da_4dvar.py:1195,1207) ----------------------------------------------
// Phase 1 of the ordered compaction: count nonzeros per 1024-element chunk.
void launch_compact_count(const float* H, long long n, int* chunk_counts, cudaStream_t s);
// Phase 2: exclusive scan of the chunk counts (single block), total in counts[nchunks].
void launch_compact_scan(int* chunk_counts, long long nchunks, cudaStream_t s);
// Phase 3: ordered write of the flat indices (ascending == torch.nonzero order) + gather of y and 1/R.
void launch_compact_write(const float* H, const float* yo, const float* R, long long n, const int* chunk_offsets,
                          int* idx, float* y, float* rinv, cudaStream_t s);
// J_obs partials + residuals: for obs k: x = xn[idx]*sigma_c + mu_c ; r = x - y ; J += 0.5*rinv*r^2 ; resid = coeff*sigma_c*rinv*r
void launch_obs_misfit(const float* xn_all, const int* idx, const float* y, const float* rinv, const float* sigma,
                       const float* mu, long long n_obs, long long HW, int C, float coeff, float* resid,
                       double* block_partials, int nblocks, cudaStream_t s);
// out[0] = sum(partials[0..n))  (fp64, one block)
void launch_reduce_partials(const double* partials, int n, double* out, cudaStream_t s);
// G[idx - base] += resid for obs in [k0,k1)   (indices unique -> no atomics)
void launch_obs_adjoint(float* G, const int* idx, const float* resid, long long k0, long long k1, long long base,
                        cudaStream_t s);

// ---- vector algebra for L-BFGS and J_reg (torch/optim/lbfgs.py:428-443) ---------------------------
// out[j] = sum_i a_j[i]*b_j[i] for up to 4 (a,b) pairs sharing one pass; fp64 block partials, finalised in-kernel
struct DotPairs { const float* a[4]; const float* b[4]; int n_pairs; };
void launch_multi_dot(const DotPairs& p, long long n, double* out /*[4] device*/, double* scratch, cudaStream_t s);
// y = alpha*x + beta*y with alpha/beta read from device doubles (null -> given host constant)
void launch_axpby(float* y, const float* x, const double* alpha_dev, double alpha_host, const double* beta_dev,
                  double beta_host, long long n, cudaStream_t s);
// y += scale * (*a1 - (a2 ? *a2 : 0)) * x   (second loop of the two-loop recursion: (al_i - be_i) s_i)
void launch_axpy_diff(float* y, const float* x, const double* a1_dev, const double* a2_dev, double scale, long long n, cudaStream_t s);
// max-abs and L1 norm of a vector: out[0]=max|x|, out[1]=sum|x|
void launch_absmax_l1(const float* x, long long n, double* out, double* scratch, cudaStream_t s);

// Latitude-weighted RMSE and bias of every channel (utils/metrics.py WRMSE / Bias as da_4dvar.py:1260-1264 applies them):
// x, gt physical (C,H,W); out[0..C) = WRMSE, out[C..2C) = Bias.  w_scratch: H floats, partials: metrics_scratch_doubles(C).
int metrics_scratch_doubles(int C);
void launch_metrics(const float* x, const float* gt, const float* mean, const float* sigma, int C, int H, int W, float* w_scratch,
                    double* partials, double* out, cudaStream_t s);

int reduce_blocks();  // number of blocks the reduction kernels use (size of scratch in doubles * 4)

}  // namespace vv

// =================================================================================================
// Forecast network LGUnet_all_1 (networks/LGUnet_all.py:743-777), forward only -- kernels it needs beyond the ones above.
// =================================================================================================
namespace vv {

// rope2 (networks/utils/positional_encodings.py:230-268) applied in place to the q and k thirds of a 16-bit qkv buffer.
// A token's position is its (row, col) inside its (shifted) window; table[pos][j] = (cos, sin) of the angle of pair j < hd / 2
// (pairs j < hd / 4 rotate with the row, the others with the column); the pair is (x[j], x[hd / 2 + j]).
struct RopeArgs {
  int gh, gw, wh, ww, sh, sw, heads, hd, batch;
  bf16* qkv; long long ld_qkv, qkv_bs;
  const float2* table;         // [wh * ww][hd / 2]
};
void launch_rope(const RopeArgs& a, cudaStream_t s);

// SD_attn (networks/utils/Attention.py:551-650, dilation 1) on rotated q, k: softmax(scale q k^T + mask) v over windows of
// wh x ww tokens of the rolled frame (roll by -(sh, sw), Attention.py:560), any window size up to the whole grid (the first
// LG stage, networks/LGUnet_all.py:689), flash-style (online softmax over key tiles).  mask: 0 / -inf between the two latitude
// bands of the last window row of a shifted frame (create_mask, Attention.py:520-548; longitude never masks).
struct Attn1Args {
  int gh, gw, wh, ww, sh, sw, heads, hd, batch, mask;
  const bf16* qkv; long long ld_qkv, qkv_bs;     // fp16 [batch][gh * gw][3 * heads * hd]
  bf16* out; long long ld_o, o_bs;               // fp16 [batch][gh * gw][heads * hd]
  float scale;
  bf16* vt; long long vt_ld;                     // optional scratch [heads * hd][vt_ld >= round_up(gh * gw, 64)] fp16: with it, an unshifted
                                                 // whole-grid window of head width 192 (the first LG stage) runs on the tcgen05 kernel
  const float2* rope;                            // optional rope2 table [wh * ww][hd / 2] (cos, sin): windows of <= 80 tokens rotate q and k
                                                 // on their way through shared memory (qkv is then left unrotated); null = q, k are rotated already
};
bool attn1_fuses_rope(int wh, int ww);
void launch_attn1(const Attn1Args& a, cudaStream_t s);
bool attn1_tc_eligible(const Attn1Args& a);
bool attn1_supported(int hd);

// PatchEmbed with a 3 x 2 kernel and stride 2 (networks/LGUnet_all.py:14-50) + absolute position embedding (:396-400):
// tok[g][i * w0 + j][c] = bias[g][c] + ape[g][tok][c] + sum_{ci, kr, kc} W[g][(ci * 3 + kr) * 2 + kc][c] img[chan(g, ci)][2 i + kr][2 j + kc]
struct Patch32Args {
  int H, W, h0, w0, G, D;
  const int* kcnt; const int* cbase;             // channels per group, first NCHW channel of the group (device)
  const float* Wp;                               // [sum_g kcnt[g] * 6][D], rows of group g start at cbase[g] * 6
  const float* bias;                             // [G][D]
  const float* ape;                              // [G][h0 * w0][D]
  const float* img; float* tok;
  int max_cnt;
};
void launch_patch32(const Patch32Args& a, cudaStream_t s);
bool patch32_supported(int D);

// The ConvTranspose2d head, kernel 3 x 2 / stride 2 (networks/LGUnet_all.py:606-650): adjacent patch rows overlap-add on even
// image rows.  img[chan[slot]][y][x] = bias[slot] + sum_c sum_{(i, kr): 2 i + kr = y} tok[g][i * w0 + x / 2][c] Wt[slot][kr][x % 2][c]
struct ConvT32Args {
  int H, W, h0, w0, G, D;
  const int* kcnt; const int* cbase; const int* chan;   // per group: output slots, first slot; NCHW channel of every slot (device)
  const float* Wt;                               // [slots][3][2][D]
  const float* bias;                             // [slots]
  const float* tok; float* img;
  int max_cnt;
};
void launch_convt32(const ConvT32Args& a, cudaStream_t s);

}  // namespace vv
