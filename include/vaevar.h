/* vaevar_b200 -- C ABI of the B200-native 4D-Var cost-and-gradient engine.
 *
 * The reference (xiaoyi018/VAE-Var) has no FFI of its own: its boundary for this path is the Python call
 * surface of da_4dvar.py / nf_model/vae.py / networks_old/transformer.py.  Every entry point below names the
 * reference line range it stands in for; vaevar_b200/*.py binds them with ctypes behind that same Python surface
 * (INTEGRATION.md shows the stub a maintainer of the reference would add).
 *
 * Conventions: plain pointers and sizes only.  `*_dev` pointers are CUDA device pointers owned by the caller;
 * `stream` is a cudaStream_t passed as void*; every call only ENQUEUES work on that stream unless stated.
 * Return value 0 = ok, negative = error (text from vv_last_error()).  Nothing throws across the boundary.
 * One engine per process / GPU; an engine is not thread-safe.  There is no CPU fallback anywhere.
 */
#ifndef VAEVAR_H_
#define VAEVAR_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VV_API __attribute__((visibility("default")))
#else
#define VV_API
#endif

#define VV_MAX_GROUPS 8
#define VV_MAX_LG 8

/* Hyper-parameters of one LGUnet_all (networks_old/transformer.py:717-718; nf_model/parameters0_old.yaml). */
typedef struct {
  int img_h, img_w;
  int n_groups;
  int in_chans[VV_MAX_GROUPS];   /* inchans_list  */
  int out_chans[VV_MAX_GROUPS];  /* outchans_list */
  int enc_dim, embed_dim, window;
  int enc_depth[2], enc_heads[2];
  int n_lg;
  int lg_depth[VV_MAX_LG], lg_heads[VV_MAX_LG];
  int keep_out; /* leading output channels that are produced / differentiated (69 for `model(x)[:, :69]`,
                   da_4dvar.py:674); 0 = all of them */
} vv_net_config;

typedef struct {
  vv_net_config dec;  /* VAE decoder D: VAE_lr.dec, nf_model/vae.py:69-70,83-85 */
  vv_net_config flow; /* forecast operator M: self.flow_model, da_4dvar.py:571-588 (ignored when T == 1 and has_flow == 0) */
  int has_flow;
  int T;         /* da_win: states in the assimilation window (>= 1) */
  int recompute; /* 0 = stash every application's activations; 1 = keep each application's input only and
                    recompute its forward inside the backward sweep */
  int use_graph; /* 1 = capture cost+grad into a CUDA graph on first use */
  int forward_fp16; /* 1 = forward activations and weights in IEEE fp16 (11-bit significand: 8x less rounding noise in
                       J(z) than bf16, which the strong-Wolfe line search needs), gradients in bf16; 0 = bf16 throughout */
  int no_ln_fold; /* 0 = norm1 / norm2 of every Swin block (swinblock.py:226,232) are folded into the qkv / fc1 GEMMs (the GEMM runs
                     on a row-centred 16-bit copy of the residual stream, see vv_ln_fold_health); 1 = they run as LayerNorm kernels
                     with two-pass fp32 statistics in front of plain GEMMs (slower by ~35 launches per application, and safe for
                     any residual stream) */
} vv_config;

typedef struct vv_engine vv_engine;

VV_API const char* vv_last_error(void);
/* Select the CUDA device of the calling thread for this library (one process per GPU: pass LOCAL_RANK). */
VV_API int vv_set_device(int ordinal);

/* cyclic_4dvar.__init__ -> init_vae_model / init_model_flow (da_4dvar.py:571-603): build both networks. */
VV_API int vv_engine_create(const vv_config* cfg, vv_engine** out);
VV_API void vv_engine_destroy(vv_engine* e);

/* load_state_dict (da_4dvar.py:575-585, 592-601): one fp32 tensor by its reference state_dict name.
 * net: 0 = decoder, 1 = flow.  Copies synchronously. Unknown names are an error; buffers
 * (relative_position_index, attn_mask) are recomputed by the engine and may be skipped by the caller. */
VV_API int vv_set_weight(vv_engine* e, int net, const char* name, const float* data_dev, const int64_t* shape, int ndim);
/* Pack to bf16 (+ transposes for the input-gradient GEMMs), gather the relative-position bias tables. */
VV_API int vv_finalize_weights(vv_engine* e);

/* get_model_mean_std + stdTr (da_4dvar.py:640-647, 1181): host pointers, n_state floats each. */
VV_API int vv_set_constants(vv_engine* e, const float* mean, const float* std, const float* stdTr);

/* Observation operator as an ordered compaction of the dense 0/1 mask H (da_4dvar.py:290-297, 1207):
 * idx_out = flat indices of the non-zeros in ascending order (== torch.nonzero(H.flatten())), y_out = yo[idx],
 * rinv_out = 1 / R[idx].  Outputs must hold n elements; *n_out_host receives the count (synchronises). */
VV_API int vv_compact_mask(const float* H_dev, const float* yo_dev, const float* R_dev, int64_t n, int32_t* idx_out_dev,
                    float* y_out_dev, float* rinv_out_dev, int64_t* n_out_host, void* stream);

/* The tensors the closure captures (da_4dvar.py:1248-1251): xb (C,H,W); yo, H, R (T,C,H,W); obs_coeff. Synchronises. */
VV_API int vv_set_case(vv_engine* e, const float* xb_dev, const float* yo_dev, const float* H_dev, const float* R_dev,
                float obs_coeff, void* stream);
VV_API int vv_num_obs(vv_engine* e, int64_t* n_obs);

/* The same closure on the reference's real geometry (da_4dvar.py:1185-1208 with decoder_hr, nf_model/vae.py:87-90, and
 * integrate(x, flow, 1, True, False), da_4dvar.py:666-681): xb (C,Hh,Wh) and yo, H, R (T,C,Hh,Wh) live on an analysis grid finer than
 * the network grid (721x1440 over 128x256).  Nearest resampling is an index map, so the engine keeps every field on the network grid
 * and composes the maps into the observation indices and into one gather (and its adjoint) between flow steps; J and grad_z are those
 * of the reference's loss.  vv_cost_grad / vv_cost / vv_lbfgs_* then work unchanged.  Synchronises. */
VV_API int vv_set_case_native(vv_engine* e, const float* xb_dev, const float* yo_dev, const float* H_dev, const float* R_dev, int Hh, int Wh,
                              float obs_coeff, void* stream);
/* The real-observation branch of the loss (da_4dvar.py:1196-1206): yo, H, R (T,A,Hh,Wh) live in an AUGMENTED channel space in which
 * observed channel a = sum_j tap_w[a*taps+j] * x[tap_chan[a*taps+j]] at the same grid point (the reference: A = 204 = 4 surface
 * channels + 5 variables x 40 pressure levels, rows of obs_interpolater.interp, da_4dvar.py:62-82, two taps each).  tap_chan / tap_w
 * are HOST arrays of A*taps entries (weight 0 for an unused tap).  The grid may be the network grid or a finer analysis grid; the
 * index maps are composed as for vv_set_case_native.  Synchronises. */
VV_API int vv_set_case_obsop(vv_engine* e, const float* xb_dev, const float* yo_dev, const float* H_dev, const float* R_dev, int Hh, int Wh,
                             int n_obs_channels, int taps, const int32_t* tap_chan_host, const float* tap_w_host, float obs_coeff,
                             void* stream);
/* Analysis on the analysis grid after vv_set_case_native: (decoder_hr(z) stdTr) sigma + xb  (da_4dvar.py:1257-1259, 1301-1306). */
VV_API int vv_decode_native(vv_engine* e, const float* z_dev, float* x_phys_out_dev, void* stream);

/* Health of the folded LayerNorms since the last call (or engine creation): counts[0] = (row, LayerNorm) pairs whose mean is
 * further than 32 standard deviations from the row's stage-input mean (the centred 16-bit operand then carries > 1.5 % rounding
 * error), counts[1] = rows whose centred values approach the fp16 range.  Non-zero counts mean: recreate the engine with
 * no_ln_fold = 1.  Synchronises the engine's stream; resets the counters. */
VV_API int vv_ln_fold_health(vv_engine* e, uint32_t counts_host[2]);

/* closure() (da_4dvar.py:1242-1246): J_out_dev[3] = {J, J_reg, J_obs} (fp64), grad_dev = dJ/dz (same shape as z). */
VV_API int vv_cost_grad(vv_engine* e, const float* z_dev, double* J_out_dev, float* grad_dev, void* stream);
/* cal_loss() (da_4dvar.py:1210-1236): forward sweep only. */
VV_API int vv_cost(vv_engine* e, const float* z_dev, double* J_out_dev, void* stream);
/* xb + D(z) * stdTr * sigma, physical units (da_4dvar.py:1256-1259, 1301-1306). */
VV_API int vv_decode(vv_engine* e, const float* z_dev, float* x_phys_out_dev, void* stream);
/* cyclic_4dvar.integrate(xa, flow_model, steps) (da_4dvar.py:666-681) on the engine grid. */
VV_API int vv_integrate(vv_engine* e, const float* x_phys_in_dev, float* x_phys_out_dev, int steps, void* stream);

/* LGUnet_all.forward (transformer.py:747-752) and its input-VJP, for parity tests.
 * in: (sum in_chans, H, W); out / dout: (kept out channels, H, W); din like in. */
VV_API int vv_net_forward(vv_engine* e, int net, const float* in_dev, float* out_dev, void* stream);
VV_API int vv_net_vjp(vv_engine* e, int net, const float* in_dev, const float* dout_dev, float* din_dev, void* stream);

/* Per-channel diagnostics of the outer loop (da_4dvar.py:1260-1264): out[0..C) = Metrics.WRMSE, out[C..2C) = Metrics.Bias
 * (utils/metrics.py:526-544, 473-474, 282-296, 65-82; latitude weights with the reference's literal 3.1416) of the PHYSICAL
 * fields x and gt (C,H,W), which are normalised with the constants of vv_set_constants first, as the reference does.
 * One fused pass on the device; out is a device array of 2*C doubles. */
VV_API int vv_metrics(vv_engine* e, const float* x_phys_dev, const float* gt_phys_dev, double* out_dev, void* stream);
/* The same diagnostics for (C,H,W) fields on any grid (the analysis grid of vv_set_case_native). */
VV_API int vv_metrics_grid(vv_engine* e, const float* x_phys_dev, const float* gt_phys_dev, int H, int W, double* out_dev, void* stream);

/* ---- native-resolution seams; no engine handle, any field size ---------------------------------
 * F.interpolate(x, (Ho, Wo)) with the default nearest rule (nf_model/vae.py:90 decoder_hr; da_4dvar.py:671, 679) on a (C,Hi,Wi)
 * field, fused with the per-channel (de)normalisation the reference applies on the same side of the seam:
 *   mode 0: out = in[src];  mode 1: out = (in[src] - mean[c]) / std[c]  (da_4dvar.py:667 then :671);
 *   mode 2: out = in[src] * std[c] + mean[c]  (:679 then :681).   Bit-identical to the eager reference. */
VV_API int vv_resample_nearest(const float* in_dev, float* out_dev, int C, int Hi, int Wi, int Ho, int Wo, int mode, const float* mean_dev,
                               const float* std_dev, void* stream);
/* Vector-Jacobian product of vv_resample_nearest (what autograd runs through integrate(..., True, False), da_4dvar.py:1191):
 * din (C,Hi,Wi) from dout (C,Ho,Wo); deterministic ordered sums (ascending output row, then column), bit-identical to the
 * reference's CPU backward.  mode 1: (sum) / std[c];  mode 2: sum of dout * std[c]. */
VV_API int vv_resample_nearest_adjoint(const float* dout_dev, float* din_dev, int C, int Hi, int Wi, int Ho, int Wo, int mode,
                                       const float* std_dev, void* stream);
/* Observation term on a physical-unit field of n_grid elements (da_4dvar.py:1207 with a 0/1 H compacted by vv_compact_mask):
 * J_out_dev[0] = obs_coeff * 1/2 * sum_k rinv[k] (x[idx[k]] - y[k])^2 in double; when grad_dev != NULL it is zero-filled and
 * receives obs_coeff * rinv (x - y) at the observed points.  work_dev holds vv_obs_term_work_doubles() doubles. */
VV_API int64_t vv_obs_term_work_doubles(void);
/* Host-only (no device call): the index tables the seams use between a network grid (H,W) and an analysis grid (Hh,Wh) - source row /
 * column of every up-sampled (Hh / Wh entries) and down-sampled (H / W entries) element, the round trip S = down o up (H / W entries)
 * and, per source row / column of S, the first output that reads it (H + 1 / W + 1 entries).  For tests of the index rule. */
VV_API int vv_debug_seam_tables(int H, int W, int Hh, int Wh, int32_t* up_rows, int32_t* up_cols, int32_t* down_rows, int32_t* down_cols,
                                int32_t* s_row, int32_t* s_col, int32_t* s_row_lo, int32_t* s_col_lo);
VV_API int vv_obs_term(const float* x_dev, const int32_t* idx_dev, const float* y_dev, const float* rinv_dev, int64_t n_obs, float obs_coeff,
                       double* J_out_dev, float* grad_dev, int64_t n_grid, double* work_dev, void* stream);

/* torch.optim.LBFGS([z], history_size, max_iter, line_search_fn="strong_wolfe") + .step(closure)
 * (da_4dvar.py:1240, 1298-1299; torch/optim/lbfgs.py:333-537).  State persists across steps; the vectors never
 * leave the device, the controller reads back O(10) scalars per closure evaluation. */
typedef struct vv_lbfgs vv_lbfgs;
VV_API int vv_lbfgs_create(vv_engine* e, int history_size, int max_iter, vv_lbfgs** out);
VV_API void vv_lbfgs_destroy(vv_lbfgs* o);
/* Forget history, step count and cached evaluations: the state of a new optimiser (one per DA cycle, da_4dvar.py:1240), keeping
 * the device vectors. */
VV_API int vv_lbfgs_reset(vv_lbfgs* o);
/* One optimizer.step(closure) on z_dev (updated in place). info_host[8] = {loss at entry, final loss, n closure evals
 * this step, n_iter total, last step length, |g|_inf, closure evaluations so far, evaluations skipped so far}. Synchronises. */
VV_API int vv_lbfgs_step(vv_lbfgs* o, float* z_dev, double* info_host, void* stream);
/* Losses of every closure evaluation so far (returns the total count; copies at most cap). */
VV_API int vv_lbfgs_history(vv_lbfgs* o, double* loss_out_host, int cap);
/* {J, J_reg, J_obs} at the point the last step() left z on -- what cal_loss(z) (da_4dvar.py:1210-1236, printed by the outer
 * loop at :1265-1269) evaluates again; the optimiser already has it from the accepted line-search trial. */
VV_API int vv_lbfgs_last_cost(vv_lbfgs* o, double* J3_host);
/* torch.optim.LBFGS.step() opens with a closure() at the point the previous step() ended on; its loss and gradient are already
 * held by the optimiser.  With reuse on (the default) that evaluation is skipped when z is bit-identical to the z the previous
 * step left (checked on the device): same iterates, one cost+gradient sweep less per step; it still counts against max_eval.
 * info_host[7] of vv_lbfgs_step = evaluations skipped so far.  on = 0 restores an evaluation per step() entry. */
VV_API int vv_lbfgs_set_reuse(vv_lbfgs* o, int on);
/* Relative rounding noise of the closure's loss that the line search tolerates in its "loss went up" tests (relaxed Armijo
 * condition f(t) <= f(0) + c1 t g'd + f_noise_rel |f(0)|).  Engine-bound optimisers default to 5e-5 (fp16 forward) / 4e-4
 * (bf16 forward); 0 reproduces torch.optim.LBFGS decision for decision. */
VV_API int vv_lbfgs_set_noise(vv_lbfgs* o, double f_noise_rel);
/* Trial step length t of every closure evaluation so far, aligned with vv_lbfgs_history (0 for the evaluation that opens a
 * step()). */
VV_API int vv_lbfgs_steps(vv_lbfgs* o, double* t_out_host, int cap);
/* Same controller on an analytic device function (pairwise Rosenbrock over n floats) -- lets tests compare the
 * optimiser's decisions with torch.optim.LBFGS on an identical objective. */
VV_API int vv_lbfgs_create_testfn(long long n, int history_size, int max_iter, vv_lbfgs** out);

/* Host-only (no device needed): the strong-Wolfe line search's cubic interpolation (torch/optim/lbfgs.py:12-37) with torch's scalar
 * typing -- steps / directional derivatives are 0-dim float32 tensors when *_is_tensor, losses are Python floats.  Identical to torch
 * except where torch's float32 arithmetic overflows to inf / inf = NaN: the bisection step of the negative-discriminant branch is
 * returned instead (lbfgs.cu).  has_bounds = 0: bounds = (min(x1, x2), max(x1, x2)). */
VV_API double vv_debug_cubic_interpolate(double x1, double f1, double g1, double x2, double f2, double g2, int x_is_tensor, int g_is_tensor,
                                  int has_bounds, double lo, double hi);

/* Kernel-level hooks used by tests/ and bench.py (roofline of the dominant kernel). */
/* epi: 0 linear, 1 GELU (aux out = saved gelu'(u)), 2 multiply by aux (aux in = the saved gelu'(u)); | 16: operands, 16-bit outputs
 * and aux are fp16 instead of bf16. */
VV_API int vv_test_gemm(const void* A_16_dev, const void* B_16_dev, const float* bias_dev, const float* res_dev, float* out_f32_dev,
                 void* out_16_dev, void* aux_16_dev, int M, int N, int K, int batch, int epi, void* stream);
/* GEMM with a LayerNorm folded into it (the forward pass's norm1 -> qkv and norm2 -> fc1, swinblock.py:268,305):
 * out = epi(LN(x) W^T + b) evaluated as rstd ((x - shift) W'^T - (mean - shift) s) + c, with A = the 16-bit rows of x - shift (shift =
 * a per-row offset, null = 0), B = W' = W o gamma, cbias = c = b + W beta, colsum = s = W' 1, stats = per-row (mean, M2) partials
 * [batch][parts][M] float2 left by a producer of tile width prod_bn (0 = a single partial over the row), C = row length; rows whose
 * remaining offset endangers the 16-bit operand are counted in health_dev[2] (null = off).
 * If stats_out is given the GEMM is the PRODUCER instead: it emits the partials of the rows of its fp32 output
 * ([batch][parts_out[0]][M] float2, tile width parts_out[1]) and the 16-bit copy out_16 = out - shift, as the proj / fc2 GEMMs do for
 * the LayerNorm that follows them.  res_f32 (optional): fp32 residual [batch][M][N] added before everything else. */
VV_API int vv_test_gemm_ln(const void* A_16_dev, const void* B_16_dev, const float* cbias_dev, const float* colsum_dev, const float* stats_dev,
                    int parts, int C, float eps, void* out_16_dev, void* aux_16_dev, float* out_f32_dev, float* stats_out_dev,
                    int* parts_out_host, int M, int N, int K, int batch, int epi, const float* shift_dev, int prod_bn,
                    uint32_t* health_dev, const float* res_f32_dev, void* stream);
/* Fused tower MLP (the second half of a Swin block, swinblock.py:13-29, 304-307) as ONE kernel per direction; D in {64, 96, 128, 192},
 * rows a multiple of 128.  Forward: out = x1 + fc2(gelu(fc1(normalise(x1)))) where W1 (batch, 4D, D) / b1 (batch, 4D) are fc1 with norm2's
 * gamma / beta folded in (W o gamma, b + W beta), W2 (batch, D, 4D), b2 (batch, D); u_out (batch, rows, 4D) receives gelu'(u) for the
 * backward pass; optional: out_16 = out - shift (shift (batch, rows) or null) and stats_out (batch, rows) float2 = (mean, M2) of the
 * output rows (what the next block's folded norm1 consumes).  16-bit buffers are fp16 if f16 != 0, else bf16. */
VV_API int vv_test_mlp_fwd(const float* x1_dev, const void* W1_16_dev, const void* W2_16_dev, const float* b1_dev, const float* b2_dev, int rows,
                    int batch, int D, int f16, float eps, void* u_out_16_dev, float* out_f32_dev, void* out_16_dev, const float* shift_dev,
                    float* stats_out_dev, void* stream);
/* Its input-VJP: dx = LN^T((dy W2 . u) W1) + dres with dy (batch, rows, D) bf16, u = the saved gelu'(u) (fp16 if f16 != 0), W2T (batch, 4D, D)
 * = fc2.weight^T and W1T (batch, D, 4D) = fc1.weight^T in bf16 (unfolded), gamma (batch, D) = norm2.weight, x1 the forward input, dres
 * the fp32 gradient arriving over the residual branch; writes dx fp32 and its bf16 copy. */
VV_API int vv_test_mlp_bwd(const void* dy_bf16_dev, const void* u_16_dev, const float* x1_dev, const void* W2T_bf16_dev, const void* W1T_bf16_dev,
                    const float* gamma_dev, const float* dres_dev, int rows, int batch, int D, int f16, float eps, float* dx_dev,
                    void* dx_bf16_dev, void* stream);
/* LayerNorm + ONE Linear on the same kernel (norm1 -> qkv of a tower block, swinblock.py:268-269 + :139): out_16 (batch, rows, n_out) =
 * 16bit(normalise(x) W^T + bias) with W (batch, n_out, D) 16-bit carrying gamma and bias (batch, n_out) the folded beta; n_out a
 * multiple of 16. */
VV_API int vv_test_lin_fwd(const float* x_dev, const void* W_16_dev, const float* bias_dev, int rows, int batch, int D, int n_out, int f16,
                    float eps, void* out_16_dev, void* stream);
/* Debug: fused-MLP launches built after this call stamp clock64 values of CTA 0 into trace_dev (128 x uint64: slots 0..63 epilogue warp 0,
 * 64..127 the MMA warp; tools/mlp_trace.py); null = off. */
VV_API int vv_debug_mlp_trace(void* trace_dev);
/* Statistics pass at a stage input: x (rows, C) fp32 -> out_16 = x - mean (16-bit), stats (rows) float2 = (mean, M2), shift = mean. */
VV_API int vv_test_ln_stats(const float* x_dev, void* out_16_dev, float* stats_dev, float* shift_dev, int rows, int C, int f16, void* stream);
/* Debug: GEMM launches built after this call stamp per-CTA clock64 values into trace_dev (64 x uint64 per CTA; layout in
 * tools/gemm_trace.py); null switches tracing off. */
VV_API int vv_debug_gemm_trace(void* trace_dev);
/* Debug: 1 = GEMMs issue their MMAs without loading operands, 2 = GEMMs load operands without issuing MMAs (results are
 * garbage; isolates the tensor-pipe and the TMA-feed rates), 0 = normal. */
VV_API int vv_debug_gemm_mode(int mode);
VV_API int vv_test_layernorm(const float* x_dev, const float* gamma_dev, const float* beta_dev, float* y_dev, const float* dy_dev,
                      float* dx_dev, int rows, int C, float eps, void* stream);
/* f16 = 1: qkv and out are fp16 (dout / dqkv stay bf16). */
VV_API int vv_test_winattn(const void* qkv_16_dev, const float* relbias_dev, void* out_16_dev, const void* dout_bf16_dev,
                    void* dqkv_bf16_dev, int gh, int gw, int heads, int hd, int shift, int f16, void* stream);
/* J_obs and residuals for a given normalised trajectory xn (T,C,H,W) with the case already set. */
VV_API int vv_test_obs(vv_engine* e, const float* xn_dev, double* J_obs_dev, float* grad_xn_dev, void* stream);
/* Steady-state time of every launch of one application plan (app 0 = decoder, >= 1 = flow; bwd = 0/1): each op is run
 * `reps` times between CUDA events. ms_out / kind_out / flop_out hold `cap` entries; returns the number of ops.
 * kind: 0 GEMM, 1 LN fwd, 2 LN bwd, 3 attention fwd, 4 attention bwd, 5 P2T, 6 T2P. */
VV_API int vv_profile_ops(vv_engine* e, int app, int bwd, int reps, float* ms_out, int* kind_out, double* flop_out, int* mnk_out, int cap,
                          int flush_l2 /* 1: every timed launch starts on a cold L2 (256 MiB overwritten outside the timed events) */);
/* Number of kernel launches the last vv_cost_grad enqueued (for bench.py's gpu_launches). */
VV_API int vv_last_launch_count(vv_engine* e);

/* ------------------------------------------------------------------------------------------------------------------------------
 * The forecast network LGUnet_all_1 (networks/LGUnet_all.py:743-777; da_4dvar.py:555 builds it, :1329 applies it once per cycle
 * through integrate(xa, forecast_model, 1), :666-681).  Forward only -- the DA loop never differentiates it (detach=True).
 * Its own handle: the network lives on the reference's native 721x1440 grid with patch (3, 2) / stride 2, three tower levels,
 * SD_attn (2-D RoPE, wh x ww windows, 0 / -inf latitude shift mask; networks/utils/Attention.py:467-664) and a first trunk stage
 * that attends over the whole token grid.
 * ------------------------------------------------------------------------------------------------------------------------------ */
#define VV_NET1_MAX_LEVELS 4
typedef struct {
  int img_h, img_w;              /* img_size; must be covered exactly by the (3, 2) kernel at stride 2 (721x1440 -> 360x720) */
  int n_groups;
  int in_chans[VV_MAX_GROUPS];   /* inchans_list  */
  int out_chans[VV_MAX_GROUPS];  /* outchans_list (mean | std halves per group) */
  int enc_dim, embed_dim;
  int win_h, win_w;              /* window_size */
  int n_levels;                  /* len(enc_depths) */
  int enc_depth[VV_NET1_MAX_LEVELS], enc_heads[VV_NET1_MAX_LEVELS];
  int n_lg;
  int lg_depth[VV_MAX_LG], lg_heads[VV_MAX_LG];
  int keep_out;                  /* leading output channels produced (69 for `model(x)[:, :69]`, da_4dvar.py:674); 0 = all */
} vv_net1_config;
typedef struct vv_net1 vv_net1;

/* init_model_forecast (da_4dvar.py:548-569): LGUnet_all_1(**params), load_state_dict by name, eval(). */
VV_API int vv_net1_create(const vv_net1_config* cfg, vv_net1** out);
VV_API void vv_net1_destroy(vv_net1* n);
VV_API int vv_net1_set_weight(vv_net1* n, const char* name, const float* data_dev, const int64_t* shape, int ndim);
VV_API int vv_net1_finalize(vv_net1* n);
/* LGUnet_all_1.forward (networks/LGUnet_all.py:772-777) on one sample: in (sum in_chans, H, W) -> out (keep_out, H, W), fp32. */
VV_API int vv_net1_forward(vv_net1* n, const float* in_dev, float* out_dev, void* stream);
/* get_model_mean_std (da_4dvar.py:640-647) for vv_net1_integrate: host pointers, sum(in_chans) floats each. */
VV_API int vv_net1_set_constants(vv_net1* n, const float* mean, const float* std);
/* cyclic_4dvar.integrate(xa, forecast_model, steps) (da_4dvar.py:666-681, interpolation=False): physical state in -> physical
 * state out, (x - mean) / std -> model(.)[:, :C] `steps` times -> * std + mean. */
VV_API int vv_net1_integrate(vv_net1* n, const float* x_in_dev, float* x_out_dev, int steps, void* stream);
VV_API int vv_net1_last_launch_count(vv_net1* n);
VV_API long long vv_net1_device_bytes(vv_net1* n);
/* Steady-state time of every launch of the forward plan (each `reps` times between CUDA events).  kind: 0 GEMM, 1 LayerNorm, 7 rope2,
 * 8 SD_attn, 9 patch embedding, 10 ConvTranspose2d head; flop_out: 2 M N K batch (GEMM) / 4 N^2 hd per window and head (attention);
 * mnk_out: 4 ints per op (GEMM: M, N, K, batch; attention: tokens, window tokens, head width, heads x batch).  Returns the op count. */
VV_API int vv_net1_profile_ops(vv_net1* n, int reps, float* ms_out, int* kind_out, double* flop_out, int* mnk_out, int cap);
/* Kernel-level hook: rope2 (positional_encodings.py:255-268; skipped when table_dev is null) + the SD_attn core
 * (Attention.py:560-640) on a packed fp16 qkv buffer [gh * gw][3 * heads * hd] -> out fp16 [gh * gw][heads * hd].
 * use_tc = 1: insist on the tcgen05 kernel of the whole-grid stage (error if the shape is not eligible); 0: the mma.sync kernels. */
VV_API int vv_test_attn1(void* qkv_dev, void* out_dev, const float* table_dev, int gh, int gw, int wh, int ww, int sh, int sw, int heads, int hd,
                         int mask, int use_tc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAEVAR_H_ */
