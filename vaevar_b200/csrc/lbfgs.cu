// Device-resident L-BFGS with strong-Wolfe line search, semantics of torch.optim.LBFGS as the reference uses it
// (da_4dvar.py:1240: history_size=10, max_iter=10, line_search_fn="strong_wolfe" -> lr=1, max_eval=max_iter*5/4,
// tolerance_grad=1e-7, tolerance_change=1e-9; torch/optim/lbfgs.py:333-537, _strong_wolfe :40, _cubic_interpolate :12).
// All n-vectors (z, g, d, history) stay in HBM; the two-loop recursion runs as fused dot / axpy launches whose
// scalars stay in device memory; the branchy controller reads back a handful of doubles per closure evaluation.
#include <math.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "engine.h"

using namespace vv;

// closure(z, Jdev[>=1], grad_out, stream): enqueue J(z) -> Jdev[0] and dJ/dz -> grad_out on `stream`
typedef std::function<int(const float*, double*, float*, cudaStream_t)> Closure;

struct vv_lbfgs {
  vv_engine* e = nullptr;
  Closure closure;
  cudaStream_t own_stream = nullptr;     // test-function optimisers have no engine stream
  std::vector<double> hist_loss, hist_t; // every closure evaluation: loss and trial step
  int hist = 10, max_iter = 10, max_eval = 12;
  double lr = 1.0, tol_grad = 1e-7, tol_change = 1e-9;
  long long n = 0;
  // device vectors
  float *g = nullptr, *g_prev = nullptr, *d = nullptr, *x_init = nullptr, *g_new = nullptr, *bg[2] = {nullptr, nullptr}, *ls_gprev = nullptr;
  std::vector<float*> ys_pool;     // 2*hist + 2 vectors; old_dirs (y) / old_stps (s) are views into it
  std::vector<float*> old_y, old_s;
  std::vector<float*> free_vecs;
  std::vector<double> ro;
  double H_diag = 1.0, t = 1.0, prev_loss = 0.0;
  bool have_prev = false;
  long long n_iter_total = 0, func_evals = 0;
  // device scalars + pinned read-back
  double *dsc = nullptr, *dscratch = nullptr, *Jdev = nullptr;
  double* pinned = nullptr;
  std::vector<void*> allocs;
};

namespace {

#define LB_CUDA(x)                                                                       \
  do {                                                                                   \
    cudaError_t _e = (x);                                                                \
    if (_e != cudaSuccess) {                                                             \
      set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -1;                                                                         \
    }                                                                                    \
  } while (0)

template <typename T>
T* lalloc(vv_lbfgs* o, size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
  o->allocs.push_back(p);
  return static_cast<T*>(p);
}

int readback(vv_lbfgs* o, const double* dev, int n, cudaStream_t s) {
  LB_CUDA(cudaMemcpyAsync(o->pinned, dev, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  LB_CUDA(cudaStreamSynchronize(s));
  return 0;
}
int copy_vec(vv_lbfgs* o, float* dst, const float* src, cudaStream_t s) {
  LB_CUDA(cudaMemcpyAsync(dst, src, o->n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}
// dot(a,b) -> host
int dot_host(vv_lbfgs* o, const float* a, const float* b, double* out, cudaStream_t s) {
  DotPairs p{}; p.a[0] = a; p.b[0] = b; p.n_pairs = 1;
  launch_multi_dot(p, o->n, o->dsc, o->dscratch, s);
  if (readback(o, o->dsc, 1, s)) return -1;
  *out = o->pinned[0];
  return 0;
}
int absmax_l1_host(vv_lbfgs* o, const float* x, double* mx, double* l1, cudaStream_t s) {
  launch_absmax_l1(x, o->n, o->dsc, o->dscratch, s);
  if (readback(o, o->dsc, 2, s)) return -1;
  *mx = o->pinned[0];
  if (l1) *l1 = o->pinned[1];
  return 0;
}
// closure at z: loss -> *f, gradient -> gout, and gout . dvec -> *gtd (when dvec != null), one read-back
int eval(vv_lbfgs* o, const float* z, float* gout, const float* dvec, double* f, double* gtd, cudaStream_t s) {
  int rc = o->closure(z, o->Jdev, gout, s);
  if (rc) return rc;
  if (dvec) {
    DotPairs p{}; p.a[0] = gout; p.b[0] = dvec; p.n_pairs = 1;
    launch_multi_dot(p, o->n, o->Jdev + 3, o->dscratch, s);
  }
  if (readback(o, o->Jdev, 4, s)) return -1;
  *f = o->pinned[0];
  if (gtd) *gtd = o->pinned[3];
  o->func_evals++;
  o->hist_loss.push_back(*f);
  return 0;
}

double cubic_interpolate(double x1, double f1, double g1, double x2, double f2, double g2, bool has_bounds, double lo, double hi) {
  double xmin = has_bounds ? lo : std::min(x1, x2), xmax = has_bounds ? hi : std::max(x1, x2);
  const double d1 = g1 + g2 - 3.0 * (f1 - f2) / (x1 - x2);
  const double d2sq = d1 * d1 - g1 * g2;
  if (d2sq >= 0.0) {
    const double d2 = sqrt(d2sq);
    const double mp = x1 <= x2 ? x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2.0 * d2))
                               : x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2.0 * d2));
    return std::min(std::max(mp, xmin), xmax);
  }
  return (xmin + xmax) / 2.0;
}

// lbfgs.py:40-209.  On return: *f_out / o->g hold the loss / gradient of the lowest bracket end, *t_out its step.
int strong_wolfe(vv_lbfgs* o, float* z, double t, double f, double gtd, int max_ls, double* f_out, double* t_out, int* evals,
                 cudaStream_t s) {
  const double c1 = 1e-4, c2 = 0.9;
  double d_norm;
  if (absmax_l1_host(o, o->d, &d_norm, nullptr, s)) return -1;
  auto trial = [&](double tt, double* fn, double* gn) -> int {        // _directional_evaluate, lbfgs.py:325-331
    if (copy_vec(o, z, o->x_init, s)) return -1;
    launch_axpby(z, o->d, nullptr, tt, nullptr, 1.0, o->n, s);
    o->hist_t.push_back(tt);
    return eval(o, z, o->g_new, o->d, fn, gn, s);
  };
  double f_new, gtd_new;
  if (trial(t, &f_new, &gtd_new)) return -1;
  int ls_evals = 1;
  double t_prev = 0.0, f_prev = f, gtd_prev = gtd;
  if (copy_vec(o, o->ls_gprev, o->g, s)) return -1;                   // g_prev = g
  bool done = false;
  int ls_iter = 0;
  double br[2] = {0, 0}, br_f[2] = {0, 0}, br_gtd[2] = {0, 0};
  bool single = false;   // bracket collapsed to one point (Wolfe satisfied in the bracketing phase)
  bool have_bracket = false;
  while (ls_iter < max_ls) {
    if (f_new > (f + c1 * t * gtd) || (ls_iter > 1 && f_new >= f_prev)) {
      br[0] = t_prev; br[1] = t; br_f[0] = f_prev; br_f[1] = f_new; br_gtd[0] = gtd_prev; br_gtd[1] = gtd_new;
      if (copy_vec(o, o->bg[0], o->ls_gprev, s) || copy_vec(o, o->bg[1], o->g_new, s)) return -1;
      have_bracket = true;
      break;
    }
    if (fabs(gtd_new) <= -c2 * gtd) {
      br[0] = t; br_f[0] = f_new;
      if (copy_vec(o, o->bg[0], o->g_new, s)) return -1;
      single = true; done = true; have_bracket = true;
      break;
    }
    if (gtd_new >= 0) {
      br[0] = t_prev; br[1] = t; br_f[0] = f_prev; br_f[1] = f_new; br_gtd[0] = gtd_prev; br_gtd[1] = gtd_new;
      if (copy_vec(o, o->bg[0], o->ls_gprev, s) || copy_vec(o, o->bg[1], o->g_new, s)) return -1;
      have_bracket = true;
      break;
    }
    const double min_step = t + 0.01 * (t - t_prev), max_step = t * 10.0, tmp = t;
    t = cubic_interpolate(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, true, min_step, max_step);
    t_prev = tmp; f_prev = f_new; gtd_prev = gtd_new;
    if (copy_vec(o, o->ls_gprev, o->g_new, s)) return -1;
    if (trial(t, &f_new, &gtd_new)) return -1;
    ++ls_evals; ++ls_iter;
  }
  if (!have_bracket) {                                                // reached max_ls (lbfgs.py:96-100)
    br[0] = 0.0; br[1] = t; br_f[0] = f; br_f[1] = f_new;
    if (copy_vec(o, o->bg[0], o->g, s) || copy_vec(o, o->bg[1], o->g_new, s)) return -1;
    br_gtd[0] = gtd; br_gtd[1] = gtd_new;
  }
  bool insuf = false;
  int low = 0, high = 1;
  if (!single) { if (!(br_f[0] <= br_f[1])) { low = 1; high = 0; } }
  while (!done && ls_iter < max_ls) {
    if (fabs(br[1] - br[0]) * d_norm < o->tol_change) break;
    t = cubic_interpolate(br[0], br_f[0], br_gtd[0], br[1], br_f[1], br_gtd[1], false, 0, 0);
    const double bmax = std::max(br[0], br[1]), bmin = std::min(br[0], br[1]);
    const double eps = 0.1 * (bmax - bmin);
    if (std::min(bmax - t, t - bmin) < eps) {
      if (insuf || t >= bmax || t <= bmin) {
        t = fabs(t - bmax) < fabs(t - bmin) ? bmax - eps : bmin + eps;
        insuf = false;
      } else {
        insuf = true;
      }
    } else {
      insuf = false;
    }
    if (trial(t, &f_new, &gtd_new)) return -1;
    ++ls_evals; ++ls_iter;
    if (f_new > (f + c1 * t * gtd) || f_new >= br_f[low]) {
      br[high] = t; br_f[high] = f_new; br_gtd[high] = gtd_new;
      if (copy_vec(o, o->bg[high], o->g_new, s)) return -1;
      if (br_f[0] <= br_f[1]) { low = 0; high = 1; } else { low = 1; high = 0; }
    } else {
      if (fabs(gtd_new) <= -c2 * gtd) {
        done = true;
      } else if (gtd_new * (br[high] - br[low]) >= 0) {
        br[high] = br[low]; br_f[high] = br_f[low]; br_gtd[high] = br_gtd[low];
        if (copy_vec(o, o->bg[high], o->bg[low], s)) return -1;
      }
      br[low] = t; br_f[low] = f_new; br_gtd[low] = gtd_new;
      if (copy_vec(o, o->bg[low], o->g_new, s)) return -1;
    }
  }
  const int pick = single ? 0 : low;
  *t_out = br[pick];
  *f_out = br_f[pick];
  if (copy_vec(o, o->g, o->bg[pick], s)) return -1;
  *evals = ls_evals;
  return 0;
}

}  // namespace

extern "C" {

static int lbfgs_alloc(vv_lbfgs* o, long long nn, int history_size, int max_iter, vv_lbfgs** out);

VV_API int vv_lbfgs_create(vv_engine* e, int history_size, int max_iter, vv_lbfgs** out) {
  if (!e || !out || history_size < 1 || max_iter < 1) { set_error("vv_lbfgs_create: bad argument"); return -2; }
  vv_lbfgs* o = new vv_lbfgs();
  o->e = e;
  o->closure = [e](const float* z, double* Jdev, float* g, cudaStream_t s) { return engine_cost_grad(e, z, Jdev, g, s); };
  return lbfgs_alloc(o, (long long)e->Zc * e->HW, history_size, max_iter, out);
}

// Analytic closure for testing the controller against torch.optim.LBFGS on the same function:
//   f(x) = sum_i [ 100 (x_{2i+1} - x_{2i}^2)^2 + (1 - x_{2i})^2 ]   (pairwise Rosenbrock), evaluated in fp32 on the device.
__global__ void rosenbrock_kernel(const float* x, long long n, float* g, double* partial) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n / 2; i += (long long)gridDim.x * blockDim.x) {
    const float a = x[2 * i], b = x[2 * i + 1];
    const float r = b - a * a, q = 1.0f - a;
    acc += (double)(100.0f * r * r + q * q);
    g[2 * i] = -400.0f * a * r - 2.0f * q;
    g[2 * i + 1] = 200.0f * r;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

VV_API int vv_lbfgs_create_testfn(long long n, int history_size, int max_iter, vv_lbfgs** out) {
  if (!out || n < 2 || (n & 1) || history_size < 1 || max_iter < 1) { set_error("vv_lbfgs_create_testfn: bad argument"); return -2; }
  vv_lbfgs* o = new vv_lbfgs();
  if (cudaStreamCreateWithFlags(&o->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete o; set_error("stream"); return -1; }
  double* part = nullptr;
  if (cudaMalloc(&part, 64 * sizeof(double)) != cudaSuccess) { delete o; set_error("oom"); return -1; }
  o->allocs.push_back(part);
  o->closure = [n, part](const float* z, double* Jdev, float* g, cudaStream_t s) {
    rosenbrock_kernel<<<64, 256, 0, s>>>(z, n, g, part);
    launch_reduce_partials(part, 64, Jdev, s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  };
  return lbfgs_alloc(o, n, history_size, max_iter, out);
}

VV_API int vv_lbfgs_history(vv_lbfgs* o, double* loss_out, int cap) {
  if (!o) return 0;
  const int k = (int)std::min<size_t>(o->hist_loss.size(), (size_t)cap);
  for (int i = 0; i < k && loss_out; ++i) loss_out[i] = o->hist_loss[i];
  return (int)o->hist_loss.size();
}

static int lbfgs_alloc(vv_lbfgs* o, long long nn, int history_size, int max_iter, vv_lbfgs** out) {
  o->hist = history_size; o->max_iter = max_iter; o->max_eval = max_iter * 5 / 4;
  o->n = nn;
  const size_t n = (size_t)o->n;
  float** vecs[] = {&o->g, &o->g_prev, &o->d, &o->x_init, &o->g_new, &o->bg[0], &o->bg[1], &o->ls_gprev};
  bool ok = true;
  for (auto v : vecs) ok = ok && (*v = lalloc<float>(o, n));
  for (int i = 0; i < 2 * history_size + 2 && ok; ++i) {
    float* p = lalloc<float>(o, n);
    ok = ok && p;
    o->free_vecs.push_back(p);
  }
  o->dsc = lalloc<double>(o, 8 + 2 * history_size);
  o->dscratch = lalloc<double>(o, 4 * reduce_blocks());
  o->Jdev = lalloc<double>(o, 8);
  ok = ok && o->dsc && o->dscratch && o->Jdev && cudaMallocHost(&o->pinned, 16 * sizeof(double)) == cudaSuccess;
  if (!ok) { set_error("vv_lbfgs_create: out of memory"); vv_lbfgs_destroy(o); return -1; }
  *out = o;
  return 0;
}

VV_API void vv_lbfgs_destroy(vv_lbfgs* o) {
  if (!o) return;
  cudaDeviceSynchronize();
  for (void* p : o->allocs) cudaFree(p);
  if (o->pinned) cudaFreeHost(o->pinned);
  if (o->own_stream) cudaStreamDestroy(o->own_stream);
  delete o;
}

VV_API int vv_lbfgs_step(vv_lbfgs* o, float* z, double* info, void* stream) {
  if (!o || !z) { set_error("vv_lbfgs_step: null argument"); return -2; }
  cudaStream_t user = (cudaStream_t)stream;
  cudaStream_t s = o->e ? o->e->stream : o->own_stream;   // everything runs on a private capturable stream
  if (o->e) { if (fence_in(o->e, user)) return -1; }
  else LB_CUDA(cudaStreamSynchronize(user));
  const long long n = o->n;
  double loss, gmax, gl1;
  int rc = eval(o, z, o->g, nullptr, &loss, nullptr, s);             // lbfgs.py:361-366
  if (rc) return rc;
  const double orig_loss = loss;
  int current_evals = 1;
  if (absmax_l1_host(o, o->g, &gmax, &gl1, s)) return -1;
  int n_iter = 0;
  if (gmax > o->tol_grad) {
    while (n_iter < o->max_iter) {
      ++n_iter; ++o->n_iter_total;
      if (o->n_iter_total == 1) {                                      // lbfgs.py:395-401
        launch_axpby(o->d, o->g, nullptr, -1.0, nullptr, 0.0, n, s);
        for (float* p : o->old_y) o->free_vecs.push_back(p);
        for (float* p : o->old_s) o->free_vecs.push_back(p);
        o->old_y.clear(); o->old_s.clear(); o->ro.clear();
        o->H_diag = 1.0;
      } else {
        float* y = o->free_vecs.back(); o->free_vecs.pop_back();
        float* sv = o->free_vecs.back(); o->free_vecs.pop_back();
        if (copy_vec(o, y, o->g, s)) return -1;
        launch_axpby(y, o->g_prev, nullptr, -1.0, nullptr, 1.0, n, s);               // y = g - g_prev
        launch_axpby(sv, o->d, nullptr, o->t, nullptr, 0.0, n, s);                    // s = t d
        DotPairs p{}; p.a[0] = y; p.b[0] = sv; p.a[1] = y; p.b[1] = y; p.n_pairs = 2;
        launch_multi_dot(p, n, o->dsc, o->dscratch, s);
        if (readback(o, o->dsc, 2, s)) return -1;
        const double ys = o->pinned[0], yy = o->pinned[1];
        if (ys > 1e-10) {                                                             // lbfgs.py:407-421
          if ((int)o->old_y.size() == o->hist) {
            o->free_vecs.push_back(o->old_y.front()); o->free_vecs.push_back(o->old_s.front());
            o->old_y.erase(o->old_y.begin()); o->old_s.erase(o->old_s.begin()); o->ro.erase(o->ro.begin());
          }
          o->old_y.push_back(y); o->old_s.push_back(sv); o->ro.push_back(1.0 / ys);
          o->H_diag = ys / yy;
        } else {
          o->free_vecs.push_back(y); o->free_vecs.push_back(sv);
        }
        const int k = (int)o->old_y.size();
        double* al = o->dsc + 8;                                                      // al[i] kept on the device (unscaled by ro)
        double* be = o->dsc + 4;
        launch_axpby(o->d, o->g, nullptr, -1.0, nullptr, 0.0, n, s);                  // q = -g
        for (int i = k - 1; i >= 0; --i) {                                            // lbfgs.py:431-435
          DotPairs q{}; q.a[0] = o->old_s[i]; q.b[0] = o->d; q.n_pairs = 1;
          launch_multi_dot(q, n, al + i, o->dscratch, s);
          launch_axpby(o->d, o->old_y[i], al + i, -o->ro[i], nullptr, 1.0, n, s);     // q -= (ro_i s_i.q) y_i
        }
        launch_axpby(o->d, o->d, nullptr, 0.0, nullptr, o->H_diag, n, s);             // r = q H_diag
        for (int i = 0; i < k; ++i) {                                                 // lbfgs.py:439-443
          DotPairs q{}; q.a[0] = o->old_y[i]; q.b[0] = o->d; q.n_pairs = 1;
          launch_multi_dot(q, n, be, o->dscratch, s);
          launch_axpy_diff(o->d, o->old_s[i], al + i, be, o->ro[i], n, s);            // r += ro_i (s_i.q - y_i.r) s_i
        }
      }
      if (copy_vec(o, o->g_prev, o->g, s)) return -1;
      o->have_prev = true;
      o->prev_loss = loss;
      if (o->n_iter_total == 1) {
        if (absmax_l1_host(o, o->g, &gmax, &gl1, s)) return -1;
        o->t = std::min(1.0, 1.0 / gl1) * o->lr;                                      // lbfgs.py:454-457
      } else {
        o->t = o->lr;
      }
      double gtd;
      if (dot_host(o, o->g, o->d, &gtd, s)) return -1;
      if (gtd > -o->tol_change) break;                                                // lbfgs.py:460-464
      if (copy_vec(o, o->x_init, z, s)) return -1;
      int ls_evals = 0;
      double t_new, f_new;
      if (strong_wolfe(o, z, o->t, loss, gtd, o->max_eval - current_evals, &f_new, &t_new, &ls_evals, s)) return -1;
      loss = f_new; o->t = t_new;
      if (copy_vec(o, z, o->x_init, s)) return -1;
      launch_axpby(z, o->d, nullptr, o->t, nullptr, 1.0, n, s);                       // z += t d  (lbfgs.py:488)
      if (absmax_l1_host(o, o->g, &gmax, nullptr, s)) return -1;
      current_evals += ls_evals;
      if (n_iter == o->max_iter) break;
      if (current_evals >= o->max_eval) break;
      if (gmax <= o->tol_grad) break;
      double dmax;
      if (absmax_l1_host(o, o->d, &dmax, nullptr, s)) return -1;
      if (dmax * fabs(o->t) <= o->tol_change) break;
      if (fabs(loss - o->prev_loss) < o->tol_change) break;
    }
  }
  LB_CUDA(cudaStreamSynchronize(s));
  if (info) {
    info[0] = orig_loss; info[1] = loss; info[2] = current_evals; info[3] = (double)o->n_iter_total;
    info[4] = o->t; info[5] = gmax; info[6] = (double)o->func_evals; info[7] = 0;
  }
  return 0;
}

}  // extern "C"
