// HBM-bound kernels of the cost function outside the networks:
//  * observation operator H as an ordered compaction of the 0/1 mask (indices == torch.nonzero order),
//    fused gather + (de)normalisation + R^-1-weighted misfit with warp-shuffle reductions, and its adjoint scatter
//    (da_4dvar.py:1195,1207; 667,681);
//  * the vector algebra of the L-BFGS two-loop recursion with device-resident scalars (torch/optim/lbfgs.py:428-443).
#include "ops.h"

namespace vv {

constexpr int RED_BLOCKS = 296;   // 2 x 148 SMs
constexpr int RED_THREADS = 256;
int reduce_blocks() { return RED_BLOCKS; }

VV_DEVINL double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide fp64 sum; result valid in thread 0.
VV_DEVINL double block_sum_d(double v) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    v = warp_sum_d(v);
  }
  return v;
}

// ---- ordered compaction ---------------------------------------------------------------------
constexpr int CHUNK = 1024;   // elements per block: 256 threads x 4 consecutive elements

__global__ void __launch_bounds__(256) compact_count_kernel(const float* H, long long n, int* counts) {
  const long long base = (long long)blockIdx.x * CHUNK + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (base + k < n && H[base + k] != 0.f) ++c;
  __shared__ int sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    counts[blockIdx.x] = t;
  }
}

// In-place exclusive scan of counts[0..n) by one block; counts[n] receives the total.
__global__ void __launch_bounds__(1024) compact_scan_kernel(int* counts, long long n) {
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < n; base += 1024) {
    const long long i = base + threadIdx.x;
    const int v = i < n ? counts[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {       // Hillis-Steele inclusive scan
      const int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const int incl = sh[threadIdx.x];
    const int c0 = carry;
    if (i < n) counts[i] = c0 + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c0 + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[n] = carry;
}

__global__ void __launch_bounds__(256) compact_write_kernel(const float* H, const float* yo, const float* R, long long n,
                                                            const int* offsets, int* idx, float* y, float* rinv) {
  const long long base = (long long)blockIdx.x * CHUNK + threadIdx.x * 4;
  bool nz[4];
  float hv[4];
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    hv[k] = base + k < n ? H[base + k] : 0.f;
    nz[k] = hv[k] != 0.f;
    c += nz[k];
  }
  // exclusive scan of c over the block (thread order == index order)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __shared__ int wsum[8];
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  int woff = 0;
  for (int k = 0; k < w; ++k) woff += wsum[k];
  int pos = offsets[blockIdx.x] + woff + incl - c;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (nz[k]) {
      idx[pos] = (int)(base + k);
      y[pos] = yo[base + k];
      rinv[pos] = hv[k] / R[base + k];      // H / R: the reference's sum(H (x - yo)^2 / R) for ANY weight H (da_4dvar.py:1207); == 1 / R bit for bit when H = 1
      ++pos;
    }
  }
}

void launch_compact_count(const float* H, long long n, int* counts, cudaStream_t s) {
  compact_count_kernel<<<(unsigned)((n + CHUNK - 1) / CHUNK), 256, 0, s>>>(H, n, counts);
}
void launch_compact_scan(int* counts, long long nchunks, cudaStream_t s) { compact_scan_kernel<<<1, 1024, 0, s>>>(counts, nchunks); }
void launch_compact_write(const float* H, const float* yo, const float* R, long long n, const int* offsets, int* idx, float* y,
                          float* rinv, cudaStream_t s) {
  compact_write_kernel<<<(unsigned)((n + CHUNK - 1) / CHUNK), 256, 0, s>>>(H, yo, R, n, offsets, idx, y, rinv);
}

// ---- misfit + adjoint -----------------------------------------------------------------------
// Four consecutive observations per thread and iteration: idx / y / 1/R are read and the residuals written as 16-byte vectors
// (coalesced), the four gathers of x are independent and in flight together (the gather chain idx -> x is the latency that
// bounds this kernel), the channel comes from a 32-bit division.
__global__ void __launch_bounds__(RED_THREADS) obs_misfit_kernel(const float* __restrict__ xn, const int* __restrict__ idx,
                                                                 const float* __restrict__ y, const float* __restrict__ rinv,
                                                                 const float* __restrict__ sigma, const float* __restrict__ mu,
                                                                 long long n_obs, long long HW, int C, float coeff,
                                                                 float* __restrict__ resid, double* __restrict__ partials) {
  double acc = 0.0;
  const unsigned hw = (unsigned)HW, uc = (unsigned)C;
  const long long n4 = n_obs >> 2;
  for (long long k4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; k4 < n4; k4 += (long long)gridDim.x * blockDim.x) {
    const int4 id = __ldg(reinterpret_cast<const int4*>(idx) + k4);
    const float4 yy = __ldg(reinterpret_cast<const float4*>(y) + k4);
    const float4 ri = __ldg(reinterpret_cast<const float4*>(rinv) + k4);
    const float x0 = __ldg(xn + id.x), x1 = __ldg(xn + id.y), x2 = __ldg(xn + id.z), x3 = __ldg(xn + id.w);
    const unsigned c0 = ((unsigned)id.x / hw) % uc, c1 = ((unsigned)id.y / hw) % uc, c2 = ((unsigned)id.z / hw) % uc, c3 = ((unsigned)id.w / hw) % uc;
    const float s0 = sigma[c0], s1 = sigma[c1], s2 = sigma[c2], s3 = sigma[c3];
    const float r0 = fmaf(x0, s0, mu[c0]) - yy.x, r1 = fmaf(x1, s1, mu[c1]) - yy.y;     // de-normalise (da_4dvar.py:681), misfit
    const float r2 = fmaf(x2, s2, mu[c2]) - yy.z, r3 = fmaf(x3, s3, mu[c3]) - yy.w;
    float4 out;
    out.x = coeff * s0 * ri.x * r0; out.y = coeff * s1 * ri.y * r1;                      // d(coeff*J_obs)/d(xn)
    out.z = coeff * s2 * ri.z * r2; out.w = coeff * s3 * ri.w * r3;
    reinterpret_cast<float4*>(resid)[k4] = out;
    acc += 0.5 * ((double)(ri.x * r0 * r0) + (double)(ri.y * r1 * r1) + (double)(ri.z * r2 * r2) + (double)(ri.w * r3 * r3));
  }
  for (long long k = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_obs; k += (long long)gridDim.x * blockDim.x) {
    const int id = idx[k];
    const unsigned c = ((unsigned)id / hw) % uc;
    const float sg = sigma[c];
    const float r = fmaf(xn[id], sg, mu[c]) - y[k];
    const float rr = rinv[k];
    resid[k] = coeff * sg * rr * r;
    acc += 0.5 * (double)(rr * r * r);
  }
  acc = block_sum_d(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
void launch_obs_misfit(const float* xn_all, const int* idx, const float* y, const float* rinv, const float* sigma, const float* mu,
                       long long n_obs, long long HW, int C, float coeff, float* resid, double* block_partials, int nblocks,
                       cudaStream_t s) {
  obs_misfit_kernel<<<nblocks, RED_THREADS, 0, s>>>(xn_all, idx, y, rinv, sigma, mu, n_obs, HW, C, coeff, resid, block_partials);
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* partials, int n, double* out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
  acc = block_sum_d(acc);
  if (threadIdx.x == 0) out[0] = acc;
}
void launch_reduce_partials(const double* partials, int n, double* out, cudaStream_t s) {
  reduce_partials_kernel<<<1, 256, 0, s>>>(partials, n, out);
}

__global__ void __launch_bounds__(256) obs_adjoint_kernel(float* G, const int* idx, const float* resid, long long k0, long long k1,
                                                          long long base) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long k = k0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; k + stride < k1; k += 2 * stride) {               // indices are unique: plain read-modify-write, two in flight
    const int i0 = idx[k], i1 = idx[k + stride];
    const float a0 = resid[k], a1 = resid[k + stride];
    float* g0 = G + (i0 - base); float* g1 = G + (i1 - base);
    const float v0 = *g0, v1 = *g1;
    *g0 = v0 + a0; *g1 = v1 + a1;
  }
  if (k < k1) G[idx[k] - base] += resid[k];
}
void launch_obs_adjoint(float* G, const int* idx, const float* resid, long long k0, long long k1, long long base, cudaStream_t s) {
  if (k1 <= k0) return;
  const long long n = k1 - k0;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  obs_adjoint_kernel<<<blocks, 256, 0, s>>>(G, idx, resid, k0, k1, base);
}

// ---- vector algebra -------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) multi_dot_kernel(const DotPairs p, long long n, double* scratch) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < p.n_pairs) acc[j] += (double)(p.a[j][i] * p.b[j][i]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < p.n_pairs) {
      const double v = block_sum_d(acc[j]);
      if (threadIdx.x == 0) scratch[j * RED_BLOCKS + blockIdx.x] = v;
    }
  }
}
__global__ void __launch_bounds__(256) multi_dot_final_kernel(const double* scratch, int n_pairs, double* out) {
  for (int j = 0; j < n_pairs; ++j) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < RED_BLOCKS; i += blockDim.x) acc += scratch[j * RED_BLOCKS + i];
    acc = block_sum_d(acc);
    if (threadIdx.x == 0) out[j] = acc;
  }
}
void launch_multi_dot(const DotPairs& p, long long n, double* out, double* scratch, cudaStream_t s) {
  multi_dot_kernel<<<RED_BLOCKS, RED_THREADS, 0, s>>>(p, n, scratch);
  multi_dot_final_kernel<<<1, 256, 0, s>>>(scratch, p.n_pairs, out);
}

__global__ void __launch_bounds__(256) axpby_kernel(float* y, const float* x, const double* alpha_dev, double alpha_host,
                                                    const double* beta_dev, double beta_host, long long n) {
  const float al = (float)((alpha_dev ? *alpha_dev : 1.0) * alpha_host);
  const float be = (float)((beta_dev ? *beta_dev : 1.0) * beta_host);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float yv = be == 0.f ? 0.f : be * y[i];
    y[i] = fmaf(al, x[i], yv);
  }
}
void launch_axpby(float* y, const float* x, const double* alpha_dev, double alpha_host, const double* beta_dev, double beta_host,
                  long long n, cudaStream_t s) {
  axpby_kernel<<<1184, 256, 0, s>>>(y, x, alpha_dev, alpha_host, beta_dev, beta_host, n);
}

__global__ void __launch_bounds__(256) axpy_diff_kernel(float* y, const float* x, const double* a1, const double* a2, double scale,
                                                        long long n) {
  const float al = (float)(scale * (*a1 - (a2 ? *a2 : 0.0)));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(al, x[i], y[i]);
}
void launch_axpy_diff(float* y, const float* x, const double* a1_dev, const double* a2_dev, double scale, long long n, cudaStream_t s) {
  axpy_diff_kernel<<<1184, 256, 0, s>>>(y, x, a1_dev, a2_dev, scale, n);
}

__global__ void __launch_bounds__(RED_THREADS) absmax_l1_kernel(const float* x, long long n, double* scratch) {
  double mx = 0.0, l1 = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = fabs((double)x[i]);
    mx = v > mx ? v : mx;
    l1 += v;
  }
  l1 = block_sum_d(l1);
  // block max via shared memory
  __shared__ double shm[RED_THREADS];
  shm[threadIdx.x] = mx;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) shm[threadIdx.x] = fmax(shm[threadIdx.x], shm[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    scratch[blockIdx.x] = shm[0];
    scratch[RED_BLOCKS + blockIdx.x] = l1;
  }
}
__global__ void __launch_bounds__(256) absmax_l1_final_kernel(const double* scratch, double* out) {
  __shared__ double shm[256];
  double mx = 0.0, l1 = 0.0;
  for (int i = threadIdx.x; i < RED_BLOCKS; i += blockDim.x) {
    mx = fmax(mx, scratch[i]);
    l1 += scratch[RED_BLOCKS + i];
  }
  l1 = block_sum_d(l1);
  shm[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) shm[threadIdx.x] = fmax(shm[threadIdx.x], shm[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = shm[0];
    out[1] = l1;
  }
}
void launch_absmax_l1(const float* x, long long n, double* out, double* scratch, cudaStream_t s) {
  absmax_l1_kernel<<<RED_BLOCKS, RED_THREADS, 0, s>>>(x, n, scratch);
  absmax_l1_final_kernel<<<1, 256, 0, s>>>(scratch, out);
}

// ---- diagnostics: latitude-weighted RMSE and bias of every channel in one pass -----------------------------------
// utils/metrics.py:282-296 (weighted_rmse_torch_channels), :65-82 (type_weighted_bias_torch_channels, "all"),
// Metrics.WRMSE / Metrics.Bias :526-544, :473-474, applied the way da_4dvar.py:1260-1264 does: both fields are
// normalised with (mean, std) first, the per-channel result is multiplied by std.
//   w_j = H cos(3.1416/180 lat_j) / sum_j cos(3.1416/180 lat_j),  lat_j = 90 - 180 j / (H - 1)     (metrics.py:5-10)
constexpr int MET_SPLIT = 8;    // blocks per channel

__global__ void lat_weight_kernel(float* w, int H) {           // one block
  __shared__ float red[32];
  float s = 0.f;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const float lat = 90.0f - (float)j * 180.0f / (float)(H - 1);
    const float c = cosf(3.1416f / 180.0f * lat);
    w[j] = c;
    s += c;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const float tot = red[0];
  for (int j = threadIdx.x; j < H; j += blockDim.x) w[j] = (float)H * w[j] / tot;
}

__global__ void __launch_bounds__(256) metrics_partial_kernel(const float* x, const float* gt, const float* mean, const float* sigma,
                                                              const float* w, int H, int W, double* partials) {
  const int c = blockIdx.y;
  const long long HW = (long long)H * W, base = (long long)c * HW;
  const float mu = mean[c], sd = sigma[c];
  double s1 = 0.0, s2 = 0.0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const float d = (x[base + p] - mu) / sd - (gt[base + p] - mu) / sd;     // the reference subtracts the NORMALISED fields
    const float wj = w[p / W];
    s1 += (double)(wj * d);
    s2 += (double)(wj * (d * d));
  }
  s1 = block_sum_d(s1);
  __syncthreads();
  s2 = block_sum_d(s2);
  if (threadIdx.x == 0) {
    partials[((long long)c * gridDim.x + blockIdx.x) * 2] = s1;
    partials[((long long)c * gridDim.x + blockIdx.x) * 2 + 1] = s2;
  }
}

__global__ void metrics_final_kernel(const double* partials, const float* sigma, int C, int nsplit, double inv_hw, double* out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < nsplit; ++k) { s1 += partials[((long long)c * nsplit + k) * 2]; s2 += partials[((long long)c * nsplit + k) * 2 + 1]; }
  out[c] = sqrt(s2 * inv_hw) * (double)sigma[c];           // WRMSE
  out[C + c] = s1 * inv_hw * (double)sigma[c];              // Bias
}

int metrics_scratch_doubles(int C) { return C * MET_SPLIT * 2; }
void launch_metrics(const float* x, const float* gt, const float* mean, const float* sigma, int C, int H, int W, float* w_scratch,
                    double* partials, double* out, cudaStream_t s) {
  lat_weight_kernel<<<1, 256, 0, s>>>(w_scratch, H);
  metrics_partial_kernel<<<dim3(MET_SPLIT, C), 256, 0, s>>>(x, gt, mean, sigma, w_scratch, H, W, partials);
  metrics_final_kernel<<<(C + 127) / 128, 128, 0, s>>>(partials, sigma, C, MET_SPLIT, 1.0 / ((double)H * W), out);
}

}  // namespace vv
