// Shared pieces of the native-resolution seams (seams.cu) that the engine's cost graph composes (engine.cu).
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace vv {

struct AxisMap {
  int in, out, kind;   // kind 0 identity, 1 halving, 2 scaled
  float scale;
  __device__ int src(int d) const {
    if (kind == 0) return d;
    if (kind == 1) return d >> 1;
    const int s = (int)floorf(__fmul_rn((float)d, scale));
    return s < in - 1 ? s : in - 1;
  }
  // smallest d in [0, out] with src(d) >= s.  src is non-decreasing, so a closed-form guess plus a fix-up walk is exact.
  __device__ int lower(int s) const {
    if (kind == 0) return s < out ? s : out;
    if (kind == 1) return 2 * s < out ? 2 * s : out;
    int d = (int)ceilf((float)s / scale);
    d = d < 0 ? 0 : (d > out ? out : d);
    while (d > 0 && src(d - 1) >= s) --d;
    while (d < out && src(d) < s) ++d;
    return d;
  }
};

inline AxisMap make_axis(int in, int out) {
  AxisMap a;
  a.in = in; a.out = out;
  a.kind = in == out ? 0 : (out == 2 * in ? 1 : 2);
  a.scale = (float)in / (float)out;
  return a;
}

// host copy of AxisMap::src (one correctly rounded float product, like the device)
int host_axis_src(const AxisMap& a, int d);

// Index tables of S = down o up on the network grid (H x W) for an analysis grid (Hh x Wh): the round trip
// F.interpolate(F.interpolate(f, (Hh, Wh)), (H, W)) reads f[row[i], col[j]]  (da_4dvar.py:678-679 followed by :670-671 of the next
// step; NOT the identity at 721x1440 / 128x256).  row_lo / col_lo (H + 1 / W + 1 entries) bracket, for every source row / column, the
// outputs that read it - the ranges the adjoint sums over.
void host_seam_tables(int H, int W, int Hh, int Wh, std::vector<int>& row, std::vector<int>& col, std::vector<int>& row_lo,
                      std::vector<int>& col_lo);
// out[c,i,j] = in[c,row[i],col[j]] (+ add[c,i,j])
void launch_seam_gather(float* out, const float* in, const float* add, const int* row, const int* col, int C, int H, int W, cudaStream_t s);
// din[c,r,q] = sum_{i in [row_lo[r], row_lo[r+1])} sum_{j in [col_lo[q], col_lo[q+1])} dout[c,i,j]   (ordered)
void launch_seam_gather_adjoint(float* din, const float* dout, const int* row_lo, const int* col_lo, int C, int H, int W, cudaStream_t s);
// G[idx[k] - base] += sum of resid over each run of equal idx in [k0, k1)  (idx sorted: one writer per cell, ordered sums)
void launch_obs_adjoint_runs(float* G, const int* idx, const float* resid, long long k0, long long k1, long long base, cudaStream_t s);
// Observations compacted per time level on the analysis grid (flat index within one (C,Hh,Wh) level, ascending) -> composed indices into
// the (T,C,H,W) stack of network-grid fields, stably sorted so equal cells are adjacent; y of level 0 absorbs the background:
// y0 = (y - xb_hr[p]) + mean[c], so that field * sigma + mean - y0 = field * sigma + xb_hr[p] - y.
int native_compose_sort(const int* idx_hr, const float* y, const float* rinv, const long long* off_host, int T, const float* xb_hr,
                        const float* mean, int C, int H, int W, int Hh, int Wh, int* idx_out, float* y_out, float* rinv_out, cudaStream_t s);
// Channel-mixing sparse observation operator (the reference's real-observation branch, da_4dvar.py:1196-1206: observed channel a =
// sum_j tap_w[a][j] * x[tap_chan[a][j]] at the same grid point, K taps per observed channel).  Observations compacted per time level
// in the (A, Hh, Wh) observed space; outputs: per observation K (field index, coefficient = w * sigma) taps into the (T,C,H,W) stack
// and y' = y - sum_j w_j * (xb_hr | mean), so that r = sum_j coef_j F[ia_j] - y'; and the same taps as (cell, observation, coefficient)
// triples stably sorted by cell for the adjoint.
int native_compose_taps(const int* idx_hr, const float* y, const float* rinv, const long long* off_host, int T, const float* xb_hr,
                        const float* mean, const float* sigma, const int* tap_chan, const float* tap_w, int K, int C, int H, int W, int Hh,
                        int Wh, int* tap_ia, float* tap_coef, float* y_out, float* rinv_out, int* pair_cell, int* pair_src, float* pair_coef,
                        cudaStream_t s);
// resid[k] = coeff * rinv[k] * r_k,  partials = per-block sums of rinv r^2 / 2 (double)
void launch_obs_taps_misfit(const float* F, const int* ia, const float* coef, int K, const float* y, const float* rinv, long long n,
                            float coeff, float* resid, double* partials, int nblocks, cudaStream_t s);
// G[cell - base] += sum over each run of equal cells in [q0, q1) of pair_coef * resid[pair_src]
void launch_obs_taps_adjoint(float* G, const int* pair_cell, const int* pair_src, const float* pair_coef, const float* resid, long long q0,
                             long long q1, long long base, cudaStream_t s);
// x_hr = (D[src] * stdTr[c]) * sigma[c] + xb_hr   (da_4dvar.py:1257-1259, 1301-1306 with decoder_hr)
void launch_decode_hr(const float* D, const float* stdTr, const float* sigma, const float* xb_hr, float* out, int C, int H, int W, int Hh,
                      int Wh, cudaStream_t s);
void launch_resample(const float* in, float* out, int C, int Hi, int Wi, int Ho, int Wo, int mode, const float* mean, const float* sd,
                     cudaStream_t s);

}  // namespace vv
