// Native-resolution seams: the nearest-neighbour resampling between the network grid (128x256) and the
// analysis grid (721x1440) that the reference applies with F.interpolate (nf_model/vae.py:87-90 decoder_hr; da_4dvar.py:670-671, 678-679
// integrate(interpolation=True)), its adjoint, and the observation term evaluated directly on a physical-unit field of any size
// (da_4dvar.py:1207).  All HBM-bound: one read and one write per element; one CTA per field row, float4 stores.
//
// Index rule (ATen nearest, legacy floor): src = dst when the sizes agree, dst >> 1 for an exact doubling, otherwise
// min(int(floorf(dst * (float(in) / float(out)))), in - 1).  The (de)normalisation the reference applies on the same side of the seam is
// fused with separate, correctly rounded operations (no FMA contraction), so results are bit-identical to the eager reference.
#include <cmath>
#include <cstdint>

#include "../../include/vaevar.h"
#include <cub/device/device_radix_sort.cuh>

#include "engine.h"
#include "seams.cuh"

namespace vv {
namespace {

__device__ __forceinline__ float seam_value(float v, int mode, float mu, float sd) {
  if (mode == 1) return __fdiv_rn(__fsub_rn(v, mu), sd);
  if (mode == 2) return __fadd_rn(__fmul_rn(v, sd), mu);
  return v;
}

// One CTA per RPB consecutive output rows of one channel: the row decomposition costs one 32-bit division per CTA, the threads stride
// over the columns, and a thread's column indices are computed once for its RPB rows (whose loads are independent).
constexpr int RPB = 4;
template <int VEC>
__global__ void __launch_bounds__(128) resample_kernel(const float* __restrict__ in, float* __restrict__ out, AxisMap rows, AxisMap cols,
                                                       int mode, const float* __restrict__ mean, const float* __restrict__ sd) {
  const int groups = (rows.out + RPB - 1) / RPB;
  const int c = blockIdx.x / groups, i0 = (blockIdx.x - c * groups) * RPB;
  const float mu = mode ? mean[c] : 0.f, s = mode ? sd[c] : 1.f;
  const float* src[RPB];
#pragma unroll
  for (int r = 0; r < RPB; ++r) src[r] = in + ((long long)c * rows.in + rows.src(i0 + r < rows.out ? i0 + r : rows.out - 1)) * cols.in;
  float* dst = out + ((long long)c * rows.out + i0) * cols.out;
  for (int jq = threadIdx.x; jq < cols.out / VEC; jq += blockDim.x) {
    int sj[VEC];
#pragma unroll
    for (int u = 0; u < VEC; ++u) sj[u] = cols.src(jq * VEC + u);
    float v[RPB][VEC];
#pragma unroll
    for (int r = 0; r < RPB; ++r)
#pragma unroll
      for (int u = 0; u < VEC; ++u) v[r][u] = __ldg(src[r] + sj[u]);
#pragma unroll
    for (int r = 0; r < RPB; ++r) {
      if (i0 + r >= rows.out) break;
#pragma unroll
      for (int u = 0; u < VEC; ++u) v[r][u] = seam_value(v[r][u], mode, mu, s);
      float* d = dst + (long long)r * cols.out;
      if constexpr (VEC == 4) *reinterpret_cast<float4*>(d + jq * 4) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
      else d[jq] = v[r][0];
    }
  }
}

// Adjoint when every source element is read by at most one output (down-sampling: the index map is strictly increasing): the
// gradient field is zero-filled by the caller and one thread per OUTPUT element stores its cotangent (0 + g = g, so the result is the
// ordered sum of the general kernel bit for bit).
__global__ void __launch_bounds__(128) resample_adjoint_injective_kernel(const float* __restrict__ dout, float* __restrict__ din, AxisMap rows,
                                                                         AxisMap cols, int mode, const float* __restrict__ sd) {
  const int c = blockIdx.x / rows.out, i = blockIdx.x - c * rows.out;
  const float s = mode ? sd[c] : 1.f;
  const float* g = dout + (long long)blockIdx.x * cols.out;
  float* d = din + ((long long)c * rows.in + rows.src(i)) * cols.in;
  for (int j = threadIdx.x; j < cols.out; j += blockDim.x) {
    const float v = __ldg(g + j);
    d[cols.src(j)] = mode == 2 ? __fmul_rn(v, s) : (mode == 1 ? __fdiv_rn(v, s) : v);
  }
}

// One CTA per source row (c, si), one thread per source element: ordered sum (ascending output row, then column - the order of the
// reference's CPU backward) over the outputs that read it.  mode 1: (sum) / sd[c]  (adjoint of "normalise, then resample");
// mode 2: sum of dout * sd[c]  (adjoint of "resample, then de-normalise").
__global__ void __launch_bounds__(128) resample_adjoint_kernel(const float* __restrict__ dout, float* __restrict__ din, AxisMap rows,
                                                               AxisMap cols, int mode, const float* __restrict__ sd) {
  const int c = blockIdx.x / rows.in, si = blockIdx.x - c * rows.in;
  const int i0 = rows.lower(si), i1 = rows.lower(si + 1);
  const float s = mode ? sd[c] : 1.f;
  for (int sj = threadIdx.x; sj < cols.in; sj += blockDim.x) {
    const int j0 = cols.lower(sj), j1 = cols.lower(sj + 1);
    float acc = 0.f;
    for (int i = i0; i < i1; ++i) {
      const float* g = dout + ((long long)c * rows.out + i) * cols.out;
      for (int j = j0; j < j1; ++j) acc = __fadd_rn(acc, mode == 2 ? __fmul_rn(__ldg(g + j), s) : __ldg(g + j));
    }
    din[(long long)blockIdx.x * cols.in + sj] = mode == 1 ? __fdiv_rn(acc, s) : acc;
  }
}

constexpr int OBS_BLOCKS = 148 * 4;

// J partials (double) and, when grad != nullptr, the scatter of coeff * rinv * (x - y) into the zero-filled gradient field.
__global__ void __launch_bounds__(256) obs_term_kernel(const float* __restrict__ x, const int* __restrict__ idx, const float* __restrict__ y,
                                                       const float* __restrict__ rinv, long long n, float coeff, float* __restrict__ grad,
                                                       double* __restrict__ partials) {
  double acc = 0.0;
  const long long n4 = n >> 2;
  for (long long q = blockIdx.x * 256LL + threadIdx.x; q < n4; q += gridDim.x * 256LL) {
    const int4 id = __ldg(reinterpret_cast<const int4*>(idx) + q);
    const float4 yy = __ldg(reinterpret_cast<const float4*>(y) + q), ri = __ldg(reinterpret_cast<const float4*>(rinv) + q);
    const float r0 = __ldg(x + id.x) - yy.x, r1 = __ldg(x + id.y) - yy.y, r2 = __ldg(x + id.z) - yy.z, r3 = __ldg(x + id.w) - yy.w;
    acc += (double)(ri.x * r0 * r0) + (double)(ri.y * r1 * r1) + (double)(ri.z * r2 * r2) + (double)(ri.w * r3 * r3);
    if (grad) {
      grad[id.x] = coeff * ri.x * r0; grad[id.y] = coeff * ri.y * r1; grad[id.z] = coeff * ri.z * r2; grad[id.w] = coeff * ri.w * r3;
    }
  }
  for (long long k = (n4 << 2) + blockIdx.x * 256LL + threadIdx.x; k < n; k += gridDim.x * 256LL) {
    const int id = idx[k];
    const float r = x[id] - y[k], w = rinv[k];
    acc += (double)(w * r * r);
    if (grad) grad[id] = coeff * w * r;
  }
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

__global__ void obs_term_final_kernel(const double* __restrict__ partials, int nparts, float coeff, double* __restrict__ J) {
  __shared__ double red[256];
  double t = 0.0;
  for (int k = threadIdx.x; k < nparts; k += 256) t += partials[k];
  red[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) J[0] = 0.5 * (double)coeff * red[0];
}

bool strictly_increasing(const AxisMap& a) {
  if (a.out > a.in) return false;
  for (int d = 1; d < a.out; ++d)
    if (host_axis_src(a, d) <= host_axis_src(a, d - 1)) return false;
  return true;
}

int row_threads(int per_row) { return per_row >= 128 ? 128 : (per_row <= 32 ? 32 : (per_row + 31) / 32 * 32); }

}  // namespace
}  // namespace vv

// ---- pieces the engine composes into its cost graph when the analysis grid is finer than the network grid --------------------------
namespace vv {
namespace {

__global__ void __launch_bounds__(256) seam_gather_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ add,
                                                          const int* __restrict__ row, const int* __restrict__ col, int H, int W, long long n) {
  const long long it = blockIdx.x * 256LL + threadIdx.x;
  if (it >= n) return;
  const int j = (int)(it % W);
  const long long ci = it / W;
  const int i = (int)(ci % H);
  const long long c = ci / H;
  const float v = __ldg(in + (c * H + row[i]) * W + col[j]);
  out[it] = add ? v + add[it] : v;
}

__global__ void __launch_bounds__(256) seam_gather_adjoint_kernel(float* __restrict__ din, const float* __restrict__ dout,
                                                                  const int* __restrict__ row_lo, const int* __restrict__ col_lo, int H, int W,
                                                                  long long n) {
  const long long it = blockIdx.x * 256LL + threadIdx.x;
  if (it >= n) return;
  const int q = (int)(it % W);
  const long long ci = it / W;
  const int r = (int)(ci % H);
  const long long c = ci / H;
  float acc = 0.f;
  for (int i = row_lo[r]; i < row_lo[r + 1]; ++i)
    for (int j = col_lo[q]; j < col_lo[q + 1]; ++j) acc += __ldg(dout + (c * H + i) * W + j);
  din[it] = acc;
}

__global__ void __launch_bounds__(256) obs_adjoint_runs_kernel(float* __restrict__ G, const int* __restrict__ idx, const float* __restrict__ resid,
                                                               long long k0, long long k1, long long base) {
  for (long long k = k0 + blockIdx.x * 256LL + threadIdx.x; k < k1; k += gridDim.x * 256LL) {
    const int id = idx[k];
    if (k > k0 && idx[k - 1] == id) continue;          // not the head of its run
    float a = resid[k];
    for (long long m = k + 1; m < k1 && idx[m] == id; ++m) a += resid[m];
    G[id - base] += a;
  }
}

__global__ void __launch_bounds__(256) compose_obs_kernel(const int* __restrict__ idx_hr, const float* __restrict__ y, const float* __restrict__ xb_hr,
                                                          const float* __restrict__ mean, AxisMap rows, AxisMap cols, int HhWh, int Wh, int HW,
                                                          int W, int base, long long n, long long k_off, int* __restrict__ key,
                                                          float* __restrict__ y2, int* __restrict__ iota) {
  const long long k = blockIdx.x * 256LL + threadIdx.x;
  if (k >= n) return;
  const int p = idx_hr[k];
  const int c = p / HhWh, rem = p - c * HhWh;
  const int pi = rem / Wh, pj = rem - pi * Wh;
  key[k] = base + c * HW + rows.src(pi) * W + cols.src(pj);
  y2[k] = xb_hr ? (y[k] - xb_hr[p]) + mean[c] : y[k];
  iota[k] = (int)(k_off + k);
}

__global__ void __launch_bounds__(256) permute_obs_kernel(const int* __restrict__ perm, const float* __restrict__ y2, const float* __restrict__ rinv,
                                                          long long n, float* __restrict__ y_out, float* __restrict__ rinv_out) {
  const long long k = blockIdx.x * 256LL + threadIdx.x;
  if (k >= n) return;
  const int p = perm[k];
  y_out[k] = y2[p];
  rinv_out[k] = rinv[p];
}

__global__ void __launch_bounds__(128) decode_hr_kernel(const float* __restrict__ D, const float* __restrict__ stdTr, const float* __restrict__ sigma,
                                                        const float* __restrict__ xb_hr, float* __restrict__ out, AxisMap rows, AxisMap cols) {
  const int c = blockIdx.x / rows.out, i = blockIdx.x - c * rows.out;
  const float a = stdTr[c], b = sigma[c];
  const float* src = D + ((long long)c * rows.in + rows.src(i)) * cols.in;
  const long long o = (long long)blockIdx.x * cols.out;
  for (int j = threadIdx.x; j < cols.out; j += blockDim.x)
    out[o + j] = __fadd_rn(__fmul_rn(__fmul_rn(__ldg(src + cols.src(j)), a), b), xb_hr[o + j]);
}

__global__ void __launch_bounds__(256) compose_taps_kernel(const int* __restrict__ idx_hr, const float* __restrict__ y, const float* __restrict__ xb_hr,
                                                           const float* __restrict__ mean, const float* __restrict__ sigma,
                                                           const int* __restrict__ tap_chan, const float* __restrict__ tap_w, int K,
                                                           AxisMap rows, AxisMap cols, int HhWh, int Wh, int HW, int W, int base, long long n,
                                                           long long k_off, int* __restrict__ ia, float* __restrict__ coef,
                                                           float* __restrict__ y2, int* __restrict__ iota) {
  const long long k = blockIdx.x * 256LL + threadIdx.x;
  if (k >= n) return;
  const int p = idx_hr[k];
  const int a = p / HhWh, rem = p - a * HhWh;
  const int pi = rem / Wh, pj = rem - pi * Wh;
  const int cell = rows.src(pi) * W + cols.src(pj);
  float yy = y[k];
  for (int j = 0; j < K; ++j) {
    const int c = tap_chan[a * K + j];
    const float w = tap_w[a * K + j];
    const long long q = (k_off + k) * K + j;
    ia[q - k_off * K] = base + c * HW + cell;
    coef[q - k_off * K] = w * sigma[c];
    iota[q - k_off * K] = (int)q;
    if (w != 0.f) yy -= w * (xb_hr ? xb_hr[(long long)c * HhWh + rem] : mean[c]);
  }
  y2[k] = yy;
}

__global__ void __launch_bounds__(256) permute_pairs_kernel(const int* __restrict__ perm, const float* __restrict__ coef, int K, long long n,
                                                            int* __restrict__ src, float* __restrict__ coef_out) {
  const long long q = blockIdx.x * 256LL + threadIdx.x;
  if (q >= n) return;
  const int p = perm[q];
  src[q] = p / K;
  coef_out[q] = coef[p];
}

__global__ void __launch_bounds__(256) obs_taps_misfit_kernel(const float* __restrict__ F, const int* __restrict__ ia, const float* __restrict__ coef,
                                                              int K, const float* __restrict__ y, const float* __restrict__ rinv, long long n,
                                                              float coeff, float* __restrict__ resid, double* __restrict__ partials) {
  double acc = 0.0;
  for (long long k = blockIdx.x * 256LL + threadIdx.x; k < n; k += gridDim.x * 256LL) {
    float r = -y[k];
    for (int j = 0; j < K; ++j) r = fmaf(coef[k * K + j], __ldg(F + ia[k * K + j]), r);
    const float w = rinv[k];
    resid[k] = coeff * w * r;
    acc += 0.5 * (double)(w * r * r);
  }
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) obs_taps_adjoint_kernel(float* __restrict__ G, const int* __restrict__ cell, const int* __restrict__ src,
                                                               const float* __restrict__ coef, const float* __restrict__ resid, long long q0,
                                                               long long q1, long long base) {
  for (long long q = q0 + blockIdx.x * 256LL + threadIdx.x; q < q1; q += gridDim.x * 256LL) {
    const int id = cell[q];
    if (q > q0 && cell[q - 1] == id) continue;
    float a = coef[q] * resid[src[q]];
    for (long long m = q + 1; m < q1 && cell[m] == id; ++m) a = fmaf(coef[m], resid[src[m]], a);
    G[id - base] += a;
  }
}

unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

int host_axis_src(const AxisMap& a, int d) {
  if (a.kind == 0) return d;
  if (a.kind == 1) return d >> 1;
  volatile float p = (float)d * a.scale;
  const int s = (int)floorf(p);
  return s < a.in - 1 ? s : a.in - 1;
}

void host_seam_tables(int H, int W, int Hh, int Wh, std::vector<int>& row, std::vector<int>& col, std::vector<int>& row_lo,
                      std::vector<int>& col_lo) {
  auto one = [](int n, int nh, std::vector<int>& map, std::vector<int>& lo) {
    const AxisMap up = make_axis(n, nh), down = make_axis(nh, n);      // analysis row -> network row; network row -> analysis row
    map.resize(n);
    for (int i = 0; i < n; ++i) map[i] = host_axis_src(up, host_axis_src(down, i));
    lo.assign(n + 1, n);
    int i = 0;
    for (int r = 0; r <= n; ++r) {                                     // map is non-decreasing
      while (i < n && map[i] < r) ++i;
      lo[r] = i;
    }
  };
  one(H, Hh, row, row_lo);
  one(W, Wh, col, col_lo);
}

void launch_seam_gather(float* out, const float* in, const float* add, const int* row, const int* col, int C, int H, int W, cudaStream_t s) {
  const long long n = (long long)C * H * W;
  seam_gather_kernel<<<blocks_for(n), 256, 0, s>>>(out, in, add, row, col, H, W, n);
}
void launch_seam_gather_adjoint(float* din, const float* dout, const int* row_lo, const int* col_lo, int C, int H, int W, cudaStream_t s) {
  const long long n = (long long)C * H * W;
  seam_gather_adjoint_kernel<<<blocks_for(n), 256, 0, s>>>(din, dout, row_lo, col_lo, H, W, n);
}
void launch_obs_adjoint_runs(float* G, const int* idx, const float* resid, long long k0, long long k1, long long base, cudaStream_t s) {
  if (k1 <= k0) return;
  const long long want = (k1 - k0 + 255) / 256;
  obs_adjoint_runs_kernel<<<(unsigned)(want < 1184 ? want : 1184), 256, 0, s>>>(G, idx, resid, k0, k1, base);
}
void launch_decode_hr(const float* D, const float* stdTr, const float* sigma, const float* xb_hr, float* out, int C, int H, int W, int Hh,
                      int Wh, cudaStream_t s) {
  decode_hr_kernel<<<C * Hh, 128, 0, s>>>(D, stdTr, sigma, xb_hr, out, make_axis(H, Hh), make_axis(W, Wh));
}

void launch_obs_taps_misfit(const float* F, const int* ia, const float* coef, int K, const float* y, const float* rinv, long long n,
                            float coeff, float* resid, double* partials, int nblocks, cudaStream_t s) {
  obs_taps_misfit_kernel<<<nblocks, 256, 0, s>>>(F, ia, coef, K, y, rinv, n, coeff, resid, partials);
}
void launch_obs_taps_adjoint(float* G, const int* pair_cell, const int* pair_src, const float* pair_coef, const float* resid, long long q0,
                             long long q1, long long base, cudaStream_t s) {
  if (q1 <= q0) return;
  const long long want = (q1 - q0 + 255) / 256;
  obs_taps_adjoint_kernel<<<(unsigned)(want < 1184 ? want : 1184), 256, 0, s>>>(G, pair_cell, pair_src, pair_coef, resid, q0, q1, base);
}

int native_compose_taps(const int* idx_hr, const float* y, const float* rinv, const long long* off, int T, const float* xb_hr, const float* mean,
                        const float* sigma, const int* tap_chan, const float* tap_w, int K, int C, int H, int W, int Hh, int Wh, int* tap_ia,
                        float* tap_coef, float* y_out, float* rinv_out, int* pair_cell, int* pair_src, float* pair_coef, cudaStream_t s) {
  const long long total = off[T], pairs = total * K;
  if (total == 0) return 0;
  int *iota = nullptr, *perm = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, tap_ia, pair_cell, iota, perm, (int)pairs, 0, 31, s);
  bool ok = cudaMalloc(&iota, pairs * sizeof(int)) == cudaSuccess && cudaMalloc(&perm, pairs * sizeof(int)) == cudaSuccess &&
            cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1) == cudaSuccess;
  if (ok) {
    const AxisMap rows = make_axis(H, Hh), cols = make_axis(W, Wh);
    for (int t = 0; t < T; ++t) {
      const long long n = off[t + 1] - off[t];
      if (n <= 0) continue;
      compose_taps_kernel<<<blocks_for(n), 256, 0, s>>>(idx_hr + off[t], y + off[t], t == 0 ? xb_hr : nullptr, mean, sigma, tap_chan, tap_w, K, rows,
                                                       cols, Hh * Wh, Wh, H * W, W, (int)((long long)t * C * H * W), n, off[t], tap_ia + off[t] * K,
                                                       tap_coef + off[t] * K, y_out + off[t], iota + off[t] * K);
    }
    cudaMemcpyAsync(rinv_out, rinv, total * sizeof(float), cudaMemcpyDeviceToDevice, s);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, tap_ia, pair_cell, iota, perm, (int)pairs, 0, 31, s);
    permute_pairs_kernel<<<blocks_for(pairs), 256, 0, s>>>(perm, tap_coef, K, pairs, pair_src, pair_coef);
    ok = cudaStreamSynchronize(s) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  }
  cudaFree(iota); cudaFree(perm); cudaFree(tmp);
  if (!ok) { set_error("native_compose_taps: %s", cudaGetErrorString(cudaGetLastError())); return -1; }
  return 0;
}

int native_compose_sort(const int* idx_hr, const float* y, const float* rinv, const long long* off, int T, const float* xb_hr, const float* mean,
                        int C, int H, int W, int Hh, int Wh, int* idx_out, float* y_out, float* rinv_out, cudaStream_t s) {
  const long long total = off[T];
  if (total == 0) return 0;
  int *key = nullptr, *iota = nullptr, *perm = nullptr;
  float* y2 = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, idx_out, iota, perm, (int)total, 0, 31, s);
  bool ok = cudaMalloc(&key, total * sizeof(int)) == cudaSuccess && cudaMalloc(&iota, total * sizeof(int)) == cudaSuccess &&
            cudaMalloc(&perm, total * sizeof(int)) == cudaSuccess && cudaMalloc(&y2, total * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1) == cudaSuccess;
  if (ok) {
    const AxisMap rows = make_axis(H, Hh), cols = make_axis(W, Wh);
    for (int t = 0; t < T; ++t) {
      const long long n = off[t + 1] - off[t];
      if (n <= 0) continue;
      compose_obs_kernel<<<blocks_for(n), 256, 0, s>>>(idx_hr + off[t], y + off[t], t == 0 ? xb_hr : nullptr, mean, rows, cols, Hh * Wh, Wh,
                                                      H * W, W, (int)((long long)t * C * H * W), n, off[t], key + off[t], y2 + off[t],
                                                      iota + off[t]);
    }
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, idx_out, iota, perm, (int)total, 0, 31, s);
    permute_obs_kernel<<<blocks_for(total), 256, 0, s>>>(perm, y2, rinv, total, y_out, rinv_out);
    ok = cudaStreamSynchronize(s) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  }
  cudaFree(key); cudaFree(iota); cudaFree(perm); cudaFree(y2); cudaFree(tmp);
  if (!ok) { set_error("native_compose_sort: %s", cudaGetErrorString(cudaGetLastError())); return -1; }
  return 0;
}

}  // namespace vv

using namespace vv;

#define SEAM_CHECK(cond, ...) do { if (!(cond)) { set_error(__VA_ARGS__); return -2; } } while (0)

namespace vv {
void launch_resample(const float* in, float* out, int C, int Hi, int Wi, int Ho, int Wo, int mode, const float* mean, const float* sd,
                     cudaStream_t s) {
  const AxisMap rows = make_axis(Hi, Ho), cols = make_axis(Wi, Wo);
  const bool vec = (Wo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int grid = C * ((Ho + RPB - 1) / RPB);
  if (vec) resample_kernel<4><<<grid, row_threads(Wo / 4), 0, s>>>(in, out, rows, cols, mode, mean, sd);
  else resample_kernel<1><<<grid, row_threads(Wo), 0, s>>>(in, out, rows, cols, mode, mean, sd);
}
}  // namespace vv

extern "C" {

VV_API int vv_resample_nearest(const float* in_dev, float* out_dev, int C, int Hi, int Wi, int Ho, int Wo, int mode, const float* mean_dev,
                               const float* std_dev, void* stream) {
  SEAM_CHECK(in_dev && out_dev && C >= 1 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1, "vv_resample_nearest: bad argument");
  SEAM_CHECK(mode >= 0 && mode <= 2 && (mode == 0 || (mean_dev && std_dev)), "vv_resample_nearest: mode %d needs mean and std", mode);
  SEAM_CHECK((long long)C * Ho < (1LL << 31) && (long long)C * Hi < (1LL << 31), "vv_resample_nearest: too many rows");
  launch_resample(in_dev, out_dev, C, Hi, Wi, Ho, Wo, mode, mean_dev, std_dev, (cudaStream_t)stream);
  const cudaError_t err = cudaGetLastError();
  SEAM_CHECK(err == cudaSuccess, "vv_resample_nearest: %s", cudaGetErrorString(err));
  return 0;
}

VV_API int vv_resample_nearest_adjoint(const float* dout_dev, float* din_dev, int C, int Hi, int Wi, int Ho, int Wo, int mode,
                                       const float* std_dev, void* stream) {
  SEAM_CHECK(dout_dev && din_dev && C >= 1 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1, "vv_resample_nearest_adjoint: bad argument");
  SEAM_CHECK(mode >= 0 && mode <= 2 && (mode == 0 || std_dev), "vv_resample_nearest_adjoint: mode %d needs std", mode);
  cudaStream_t s = (cudaStream_t)stream;
  SEAM_CHECK((long long)C * Ho < (1LL << 31) && (long long)C * Hi < (1LL << 31), "vv_resample_nearest_adjoint: too many rows");
  const AxisMap rows = make_axis(Hi, Ho), cols = make_axis(Wi, Wo);
  if ((Hi > Ho || Wi > Wo) && strictly_increasing(rows) && strictly_increasing(cols)) {
    const cudaError_t e0 = cudaMemsetAsync(din_dev, 0, (size_t)C * Hi * Wi * sizeof(float), s);
    SEAM_CHECK(e0 == cudaSuccess, "vv_resample_nearest_adjoint: %s", cudaGetErrorString(e0));
    resample_adjoint_injective_kernel<<<C * Ho, row_threads(Wo), 0, s>>>(dout_dev, din_dev, rows, cols, mode, std_dev);
  } else {
    resample_adjoint_kernel<<<C * Hi, row_threads(Wi), 0, s>>>(dout_dev, din_dev, rows, cols, mode, std_dev);
  }
  const cudaError_t err = cudaGetLastError();
  SEAM_CHECK(err == cudaSuccess, "vv_resample_nearest_adjoint: %s", cudaGetErrorString(err));
  return 0;
}

VV_API int vv_debug_seam_tables(int H, int W, int Hh, int Wh, int32_t* up_rows, int32_t* up_cols, int32_t* down_rows, int32_t* down_cols,
                                int32_t* s_row, int32_t* s_col, int32_t* s_row_lo, int32_t* s_col_lo) {
  SEAM_CHECK(H >= 1 && W >= 1 && Hh >= 1 && Wh >= 1 && up_rows && up_cols && down_rows && down_cols && s_row && s_col && s_row_lo && s_col_lo,
             "vv_debug_seam_tables: bad argument");
  const AxisMap ur = make_axis(H, Hh), uc = make_axis(W, Wh), dr = make_axis(Hh, H), dc = make_axis(Wh, W);
  for (int i = 0; i < Hh; ++i) up_rows[i] = host_axis_src(ur, i);
  for (int j = 0; j < Wh; ++j) up_cols[j] = host_axis_src(uc, j);
  for (int i = 0; i < H; ++i) down_rows[i] = host_axis_src(dr, i);
  for (int j = 0; j < W; ++j) down_cols[j] = host_axis_src(dc, j);
  std::vector<int> row, col, row_lo, col_lo;
  host_seam_tables(H, W, Hh, Wh, row, col, row_lo, col_lo);
  for (int i = 0; i < H; ++i) s_row[i] = row[i];
  for (int j = 0; j < W; ++j) s_col[j] = col[j];
  for (int i = 0; i <= H; ++i) s_row_lo[i] = row_lo[i];
  for (int j = 0; j <= W; ++j) s_col_lo[j] = col_lo[j];
  return 0;
}

VV_API int64_t vv_obs_term_work_doubles(void) { return OBS_BLOCKS; }

VV_API int vv_obs_term(const float* x_dev, const int32_t* idx_dev, const float* y_dev, const float* rinv_dev, int64_t n_obs, float coeff,
                       double* J_out_dev, float* grad_dev, int64_t n_grid, double* work_dev, void* stream) {
  SEAM_CHECK(x_dev && J_out_dev && work_dev && n_obs >= 0 && (n_obs == 0 || (idx_dev && y_dev && rinv_dev)), "vv_obs_term: bad argument");
  SEAM_CHECK(!grad_dev || n_grid > 0, "vv_obs_term: the gradient field needs its size");
  SEAM_CHECK(((reinterpret_cast<uintptr_t>(idx_dev) | reinterpret_cast<uintptr_t>(y_dev) | reinterpret_cast<uintptr_t>(rinv_dev)) & 15) == 0,
             "vv_obs_term: idx / y / rinv must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (grad_dev) {
    const cudaError_t e0 = cudaMemsetAsync(grad_dev, 0, (size_t)n_grid * sizeof(float), s);
    SEAM_CHECK(e0 == cudaSuccess, "vv_obs_term: %s", cudaGetErrorString(e0));
  }
  obs_term_kernel<<<OBS_BLOCKS, 256, 0, s>>>(x_dev, idx_dev, y_dev, rinv_dev, n_obs, coeff, grad_dev, work_dev);
  obs_term_final_kernel<<<1, 256, 0, s>>>(work_dev, OBS_BLOCKS, coeff, J_out_dev);
  const cudaError_t err = cudaGetLastError();
  SEAM_CHECK(err == cudaSuccess, "vv_obs_term: %s", cudaGetErrorString(err));
  return 0;
}

}  // extern "C"
