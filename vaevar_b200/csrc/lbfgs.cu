// Device-resident L-BFGS with strong-Wolfe line search, semantics of torch.optim.LBFGS as the reference uses it
// (da_4dvar.py:1240: history_size=10, max_iter=10, line_search_fn="strong_wolfe" -> lr=1, max_eval=max_iter*5/4,
// tolerance_grad=1e-7, tolerance_change=1e-9; torch/optim/lbfgs.py:333-537, _strong_wolfe :40, _cubic_interpolate :12).
// All n-vectors (z, g, d, history) stay in HBM; the two-loop recursion runs as fused dot / axpy launches whose
// scalars stay in device memory; the branchy controller reads back a handful of doubles per closure evaluation.
#include <math.h>

#include <algorithm>
#include <array>
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
  double f_noise_rel = 0.0;              // relative rounding noise of the closure's loss (0 = exact torch.optim.LBFGS tests)
  long long n = 0;
  // device vectors
  float *g = nullptr, *g_prev = nullptr, *d = nullptr, *x_init = nullptr, *g_new = nullptr, *bg[2] = {nullptr, nullptr}, *ls_gprev = nullptr;
  std::vector<float*> ys_pool;     // 2*hist + 2 vectors; old_dirs (y) / old_stps (s) are views into it
  std::vector<float*> old_y, old_s;
  std::vector<float*> free_vecs;
  std::vector<double> ro;
  double H_diag = 1.0, t = 1.0, prev_loss = 0.0;
  bool t_is_tensor = false;
  bool have_prev = false;
  long long n_iter_total = 0, func_evals = 0;
  // torch.optim.LBFGS.step() opens with closure() at the point the previous step() ended on (lbfgs.py:361-366) -- a point whose
  // loss and gradient this object already holds (strong_wolfe leaves them in last_loss / g).  The engine is deterministic, so
  // when z is bit-identical to the copy taken at the end of the previous step the evaluation is skipped and the stored values
  // are used: same trajectory, one network sweep less per step.  The skipped evaluation still counts against max_eval.
  double eval_J3[3] = {0, 0, 0};          // {J, J_reg, J_obs} of the most recent closure evaluation (engine closures)
  double acc_J3[3] = {0, 0, 0};           // ... of the point z currently sits on (what cal_loss(z) would return)
  std::vector<std::array<double, 4>> ls_log;   // (t, J, J_reg, J_obs) of the current line search's trials
  bool reuse_entry = true, have_last = false;
  unsigned long long last_generation = 0;   // engine case / constants generation the stored evaluation belongs to
  double last_loss = 0.0;
  long long skipped_evals = 0;
  float* z_last = nullptr;
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
  *f = (double)(float)o->pinned[0];          // float(closure()): the reference's loss is a float32 tensor
  o->eval_J3[0] = o->pinned[0]; o->eval_J3[1] = o->pinned[1]; o->eval_J3[2] = o->pinned[2];
  if (gtd) *gtd = o->pinned[3];
  o->func_evals++;
  o->hist_loss.push_back(*f);
  if (!std::isfinite(*f) && o->e) {
    static bool probed = false;
    if (!probed && getenv("VV_NAN_PROBE")) { probed = true; engine_nan_probe(o->e, z); }
  }
  return 0;
}

// ---- scalar arithmetic exactly as the reference performs it -------------------------------------------------------
// In torch.optim.LBFGS the step length t, the directional derivatives gtd and everything derived from them are 0-dim
// float32 tensors (flat_grad.dot(d), 1/flat_grad.abs().sum(), ...) while the losses are Python floats.  A binary op
// with a tensor operand is carried out in float32 and yields a tensor; Python-float-only ops stay double.  The line
// search is sensitive to this (d1^2 - g1 g2 cancels catastrophically in float32), so the controller mirrors it.
struct Sc {
  double v;
  bool t;   // true: 0-dim float32 tensor
};
inline Sc py(double v) { return Sc{v, false}; }
inline Sc tn(double v) { return Sc{(double)(float)v, true}; }
inline Sc op(Sc a, char o, Sc b) {
  if (a.t || b.t) {
    const float x = (float)a.v, y = (float)b.v;
    float r = 0.f;
    switch (o) { case '+': r = x + y; break; case '-': r = x - y; break; case '*': r = x * y; break; default: r = x / y; }
    return Sc{(double)r, true};
  }
  double r = 0.0;
  switch (o) { case '+': r = a.v + b.v; break; case '-': r = a.v - b.v; break; case '*': r = a.v * b.v; break; default: r = a.v / b.v; }
  return Sc{r, false};
}
inline bool lt(Sc a, Sc b) { return (a.t || b.t) ? (float)a.v < (float)b.v : a.v < b.v; }
inline bool le(Sc a, Sc b) { return (a.t || b.t) ? (float)a.v <= (float)b.v : a.v <= b.v; }
inline bool gt(Sc a, Sc b) { return lt(b, a); }
inline bool ge(Sc a, Sc b) { return le(b, a); }
inline Sc sabs(Sc a) { return Sc{fabs(a.v), a.t}; }
inline Sc pymin(Sc a, Sc b) { return lt(b, a) ? b : a; }   // Python's min(a, b)
inline Sc pymax(Sc a, Sc b) { return gt(b, a) ? b : a; }   // Python's max(a, b)

// lbfgs.py:12-37.  eps_f > 0: when the two function values differ by less than the closure's rounding noise, their difference
// carries no information; it is replaced by the trapezoid rule (f1 - f2 = (x1 - x2)(g1 + g2)/2, exact for a quadratic), which
// turns the cubic step into the secant step on the directional derivative.
Sc cubic_interpolate(Sc x1, double f1, Sc g1, Sc x2, double f2, Sc g2, bool has_bounds, Sc lo, Sc hi, double eps_f) {
  Sc xmin = lo, xmax = hi;
  if (!has_bounds) {
    if (le(x1, x2)) { xmin = x1; xmax = x2; } else { xmin = x2; xmax = x1; }
  }
  const bool noisy = eps_f > 0.0 && fabs(f1 - f2) <= 2.0 * eps_f;
  const Sc slope3 = noisy ? op(py(1.5), '*', op(g1, '+', g2)) : op(py(3.0 * (f1 - f2)), '/', op(x1, '-', x2));
  const Sc d1 = op(op(g1, '+', g2), '-', slope3);
  const Sc d2sq = op(op(d1, '*', d1), '-', op(g1, '*', g2));
  if (ge(d2sq, py(0.0))) {
    const Sc d2 = d2sq.t ? Sc{(double)sqrtf((float)d2sq.v), true} : Sc{sqrt(d2sq.v), false};
    Sc mp;
    if (le(x1, x2))
      mp = op(x2, '-', op(op(x2, '-', x1), '*', op(op(op(g2, '+', d2), '-', d1), '/', op(op(g2, '-', g1), '+', op(py(2.0), '*', d2)))));
    else
      mp = op(x1, '-', op(op(x1, '-', x2), '*', op(op(op(g1, '+', d2), '-', d1), '/', op(op(g1, '-', g2), '+', op(py(2.0), '*', d2)))));
    // The one deliberate departure from lbfgs.py:12-37: in float32 d1 * d1 overflows once |3 (f1 - f2) / (x1 - x2)| exceeds 1.8e19 (losses
    // of 1e10 over steps of 1e-10 -- the first trial steps of a cycle whose background is far from the observations); torch then gets
    // inf / inf = NaN for the step and the NaN reaches z.  A non-finite minimiser is treated like a negative discriminant: bisection.
    if (std::isfinite(mp.v)) return pymin(pymax(mp, xmin), xmax);
  }
  return op(op(xmin, '+', xmax), '/', py(2.0));
}

// lbfgs.py:40-209.  On return: *f_out / o->g hold the loss / gradient of the lowest bracket end, *t_out its step.
int strong_wolfe(vv_lbfgs* o, float* z, Sc t, double f, Sc gtd, int max_ls, double* f_out, Sc* t_out, int* evals, cudaStream_t s) {
  const Sc c1 = py(1e-4), mc2 = py(-0.9);
  double dn;
  if (absmax_l1_host(o, o->d, &dn, nullptr, s)) return -1;
  const Sc d_norm = tn(dn);
  auto trial = [&](Sc tt, double* fn, Sc* gn) -> int {                // _directional_evaluate, lbfgs.py:325-331
    if (copy_vec(o, z, o->x_init, s)) return -1;
    launch_axpby(z, o->d, nullptr, tt.v, nullptr, 1.0, o->n, s);
    o->hist_t.push_back(tt.v);
    double g = 0.0;
    int rc = eval(o, z, o->g_new, o->d, fn, &g, s);
    *gn = tn(g);
    o->ls_log.push_back({tt.v, o->eval_J3[0], o->eval_J3[1], o->eval_J3[2]});
    return rc;
  };
  // Noise-tolerant sufficient-decrease test: the closure's loss carries relative rounding noise eps (16-bit activations), and
  // the first trial steps of a cycle (t ~ 1/|g|_1, lbfgs.py:454-457) change J by less than that.  Every "did the loss go up"
  // decision therefore gets the slack eps_f = f_noise_rel |f(0)| (relaxed Armijo of noise-tolerant quasi-Newton methods);
  // the curvature test uses gradients only and is unchanged.  f_noise_rel = 0 is torch.optim.LBFGS to the letter.
  const double eps_f = o->f_noise_rel * fabs(f);
  auto armijo_fails = [&](double fnew, Sc tt) {
    return gt(py(fnew), op(op(py(f), '+', op(op(c1, '*', tt), '*', gtd)), '+', py(eps_f)));
  };
  double f_new; Sc gtd_new;
  if (trial(t, &f_new, &gtd_new)) return -1;
  int ls_evals = 1;
  Sc t_prev = py(0.0), gtd_prev = gtd;
  double f_prev = f;
  if (copy_vec(o, o->ls_gprev, o->g, s)) return -1;                   // g_prev = g
  bool done = false, single = false, have_bracket = false;
  int ls_iter = 0;
  Sc br[2] = {py(0), py(0)}, br_gtd[2] = {py(0), py(0)};
  double br_f[2] = {0, 0};
  while (ls_iter < max_ls) {
    if (armijo_fails(f_new, t) || (ls_iter > 1 && f_new >= f_prev + eps_f)) {
      br[0] = t_prev; br[1] = t; br_f[0] = f_prev; br_f[1] = f_new; br_gtd[0] = gtd_prev; br_gtd[1] = gtd_new;
      if (copy_vec(o, o->bg[0], o->ls_gprev, s) || copy_vec(o, o->bg[1], o->g_new, s)) return -1;
      have_bracket = true;
      break;
    }
    if (le(sabs(gtd_new), op(mc2, '*', gtd))) {
      br[0] = t; br_f[0] = f_new;
      if (copy_vec(o, o->bg[0], o->g_new, s)) return -1;
      single = true; done = true; have_bracket = true;
      break;
    }
    if (ge(gtd_new, py(0.0))) {
      br[0] = t_prev; br[1] = t; br_f[0] = f_prev; br_f[1] = f_new; br_gtd[0] = gtd_prev; br_gtd[1] = gtd_new;
      if (copy_vec(o, o->bg[0], o->ls_gprev, s) || copy_vec(o, o->bg[1], o->g_new, s)) return -1;
      have_bracket = true;
      break;
    }
    const Sc min_step = op(t, '+', op(py(0.01), '*', op(t, '-', t_prev)));
    const Sc max_step = op(t, '*', py(10.0));
    const Sc tmp = t;
    t = cubic_interpolate(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, true, min_step, max_step, eps_f);
    t_prev = tmp; f_prev = f_new; gtd_prev = gtd_new;
    if (copy_vec(o, o->ls_gprev, o->g_new, s)) return -1;
    if (trial(t, &f_new, &gtd_new)) return -1;
    ++ls_evals; ++ls_iter;
  }
  if (!have_bracket) {                                                // reached max_ls (lbfgs.py:96-100)
    br[0] = py(0.0); br[1] = t; br_f[0] = f; br_f[1] = f_new;
    if (copy_vec(o, o->bg[0], o->g, s) || copy_vec(o, o->bg[1], o->g_new, s)) return -1;
    br_gtd[0] = gtd; br_gtd[1] = gtd_new;
  }
  bool insuf = false;
  int low = 0, high = 1;
  if (!single && !(br_f[0] <= br_f[1])) { low = 1; high = 0; }
  while (!done && ls_iter < max_ls) {
    if (lt(op(sabs(op(br[1], '-', br[0])), '*', d_norm), py(o->tol_change))) break;
    t = cubic_interpolate(br[0], br_f[0], br_gtd[0], br[1], br_f[1], br_gtd[1], false, py(0), py(0), eps_f);
    const Sc bmax = pymax(br[0], br[1]), bmin = pymin(br[0], br[1]);
    const Sc eps = op(py(0.1), '*', op(bmax, '-', bmin));
    if (lt(pymin(op(bmax, '-', t), op(t, '-', bmin)), eps)) {
      if (insuf || ge(t, bmax) || le(t, bmin)) {
        t = lt(sabs(op(t, '-', bmax)), sabs(op(t, '-', bmin))) ? op(bmax, '-', eps) : op(bmin, '+', eps);
        insuf = false;
      } else {
        insuf = true;
      }
    } else {
      insuf = false;
    }
    if (trial(t, &f_new, &gtd_new)) return -1;
    ++ls_evals; ++ls_iter;
    if (armijo_fails(f_new, t) || f_new >= br_f[low] + eps_f) {
      br[high] = t; br_f[high] = f_new; br_gtd[high] = gtd_new;
      if (copy_vec(o, o->bg[high], o->g_new, s)) return -1;
      if (br_f[0] <= br_f[1]) { low = 0; high = 1; } else { low = 1; high = 0; }
    } else {
      if (le(sabs(gtd_new), op(mc2, '*', gtd))) {
        done = true;
      } else if (ge(op(gtd_new, '*', op(br[high], '-', br[low])), py(0.0))) {
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
  o->f_noise_rel = e->cfg.forward_fp16 ? 5e-5 : 4e-4;   // ~ measured |J_engine(z + dz) - J_engine(z) - dJ| / J for tiny dz
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

// Host-only hook (no device needed): the line search's cubic interpolation with torch's scalar typing (x / g are 0-dim float32 tensors
// when *_is_tensor, losses are Python floats), for the CPU test that pins it to torch.optim.lbfgs._cubic_interpolate and checks the
// overflow guard.  has_bounds = 0: bounds default to (min(x1, x2), max(x1, x2)).
VV_API double vv_debug_cubic_interpolate(double x1, double f1, double g1, double x2, double f2, double g2, int x_is_tensor, int g_is_tensor,
                                         int has_bounds, double lo, double hi) {
  const Sc X1 = x_is_tensor ? tn(x1) : py(x1), X2 = x_is_tensor ? tn(x2) : py(x2);
  const Sc G1 = g_is_tensor ? tn(g1) : py(g1), G2 = g_is_tensor ? tn(g2) : py(g2);
  return cubic_interpolate(X1, f1, G1, X2, f2, G2, has_bounds != 0, x_is_tensor ? tn(lo) : py(lo), x_is_tensor ? tn(hi) : py(hi), 0.0).v;
}

VV_API int vv_lbfgs_history(vv_lbfgs* o, double* loss_out, int cap) {
  if (!o) return 0;
  const int k = (int)std::min<size_t>(o->hist_loss.size(), (size_t)cap);
  for (int i = 0; i < k && loss_out; ++i) loss_out[i] = o->hist_loss[i];
  return (int)o->hist_loss.size();
}

// Back to the state of a freshly constructed optimiser (the reference builds a new torch.optim.LBFGS per cycle,
// da_4dvar.py:1240) without giving back its ~30 device vectors and pinned buffer.
VV_API int vv_lbfgs_reset(vv_lbfgs* o) {
  if (!o) { set_error("vv_lbfgs_reset: null argument"); return -2; }
  for (float* p : o->old_y) o->free_vecs.push_back(p);
  for (float* p : o->old_s) o->free_vecs.push_back(p);
  o->old_y.clear(); o->old_s.clear(); o->ro.clear();
  o->hist_loss.clear(); o->hist_t.clear(); o->ls_log.clear();
  o->H_diag = 1.0; o->t = 1.0; o->prev_loss = 0.0;
  o->t_is_tensor = false; o->have_prev = false; o->have_last = false;
  o->n_iter_total = 0; o->func_evals = 0; o->skipped_evals = 0; o->last_loss = 0.0;
  for (int k = 0; k < 3; ++k) o->eval_J3[k] = o->acc_J3[k] = 0.0;
  return 0;
}

VV_API int vv_lbfgs_last_cost(vv_lbfgs* o, double* J3_host) {
  if (!o || !J3_host) { set_error("vv_lbfgs_last_cost: null argument"); return -2; }
  if (!o->have_last) { set_error("vv_lbfgs_last_cost: no step has been taken yet"); return -2; }
  for (int k = 0; k < 3; ++k) J3_host[k] = o->acc_J3[k];
  return 0;
}

VV_API int vv_lbfgs_set_reuse(vv_lbfgs* o, int on) {
  if (!o) { set_error("vv_lbfgs_set_reuse: null argument"); return -2; }
  o->reuse_entry = on != 0;
  return 0;
}

VV_API int vv_lbfgs_set_noise(vv_lbfgs* o, double f_noise_rel) {
  if (!o || !(f_noise_rel >= 0.0)) { set_error("vv_lbfgs_set_noise: bad argument"); return -2; }
  o->f_noise_rel = f_noise_rel;
  return 0;
}

VV_API int vv_lbfgs_steps(vv_lbfgs* o, double* t_out, int cap) {
  if (!o) return 0;
  const int k = (int)std::min<size_t>(o->hist_t.size(), (size_t)cap);
  for (int i = 0; i < k && t_out; ++i) t_out[i] = o->hist_t[i];
  return (int)o->hist_t.size();
}

static int lbfgs_alloc(vv_lbfgs* o, long long nn, int history_size, int max_iter, vv_lbfgs** out) {
  o->hist = history_size; o->max_iter = max_iter; o->max_eval = max_iter * 5 / 4;
  o->n = nn;
  const size_t n = (size_t)o->n;
  float** vecs[] = {&o->g, &o->g_prev, &o->d, &o->x_init, &o->g_new, &o->bg[0], &o->bg[1], &o->ls_gprev, &o->z_last};
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
  o->hist_t.push_back(0.0);
  bool reused = false;
  if (o->reuse_entry && o->have_last && (!o->e || o->e->generation == o->last_generation)) {
    launch_axpby(o->x_init, z, nullptr, 1.0, nullptr, 0.0, n, s);                     // x_init is scratch here: z - z_last
    launch_axpby(o->x_init, o->z_last, nullptr, -1.0, nullptr, 1.0, n, s);
    double dmax0;
    if (absmax_l1_host(o, o->x_init, &dmax0, nullptr, s)) return -1;
    if (dmax0 == 0.0) { loss = o->last_loss; o->hist_loss.push_back(loss); o->skipped_evals++; reused = true; }
  }
  if (!reused) {
    int rc = eval(o, z, o->g, nullptr, &loss, nullptr, s);           // lbfgs.py:361-366
    if (rc) return rc;
    for (int k = 0; k < 3; ++k) o->acc_J3[k] = o->eval_J3[k];
  }
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
        const double ys = (double)(float)o->pinned[0], yy = (double)(float)o->pinned[1];
        if (ys > 1e-10) {                                                             // lbfgs.py:407-421
          if ((int)o->old_y.size() == o->hist) {
            o->free_vecs.push_back(o->old_y.front()); o->free_vecs.push_back(o->old_s.front());
            o->old_y.erase(o->old_y.begin()); o->old_s.erase(o->old_s.begin()); o->ro.erase(o->ro.begin());
          }
          o->old_y.push_back(y); o->old_s.push_back(sv); o->ro.push_back((double)(1.0f / (float)ys));
          o->H_diag = (double)((float)ys / (float)yy);
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
      Sc t0;
      if (o->n_iter_total == 1) {
        if (absmax_l1_host(o, o->g, &gmax, &gl1, s)) return -1;
        const Sc inv = op(py(1.0), '/', tn(gl1));                                     // 1. / flat_grad.abs().sum(): a tensor
        t0 = op(pymin(py(1.0), inv), '*', py(o->lr));                                 // lbfgs.py:454-457
      } else {
        t0 = py(o->lr);
      }
      double gtd_d;
      if (dot_host(o, o->g, o->d, &gtd_d, s)) return -1;
      const Sc gtd = tn(gtd_d);
      if (gt(gtd, py(-o->tol_change))) break;                                         // lbfgs.py:460-464
      if (copy_vec(o, o->x_init, z, s)) return -1;
      int ls_evals = 0;
      Sc t_new; double f_new;
      o->ls_log.clear();
      if (strong_wolfe(o, z, t0, loss, gtd, o->max_eval - current_evals, &f_new, &t_new, &ls_evals, s)) return -1;
      for (const auto& r : o->ls_log)                                                 // the trial the search settled on (t = 0: keep)
        if (r[0] == t_new.v) { o->acc_J3[0] = r[1]; o->acc_J3[1] = r[2]; o->acc_J3[2] = r[3]; }
      loss = f_new; o->t = t_new.v; o->t_is_tensor = t_new.t;
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
  if (copy_vec(o, o->z_last, z, s)) return -1;
  o->last_loss = loss; o->have_last = true;
  o->last_generation = o->e ? o->e->generation : 0;
  LB_CUDA(cudaStreamSynchronize(s));
  if (info) {
    info[0] = orig_loss; info[1] = loss; info[2] = current_evals; info[3] = (double)o->n_iter_total;
    info[4] = o->t; info[5] = gmax; info[6] = (double)o->func_evals; info[7] = (double)o->skipped_evals;
  }
  return 0;
}

}  // extern "C"
