// Host side of the tcgen05 GEMM: TMA descriptor encoding and tile-shape dispatch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <tuple>

#include "ops.h"

namespace vv {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D map over a [batch][rows][cols] tensor of 2-byte (fp16 / bf16 storage) or 4-byte elements, box = (box_cols x box_rows x 1).
// The box row must span 128 bytes (128B swizzle), 64 bytes (64B swizzle) or 32 bytes (no swizzle); OOB -> zeros on loads / clipped on stores.
static const char* encode_map_t(CUtensorMap* tm, const void* base, bool f32, long long K, long long rows, long long batch,
                                long long ld, long long bs, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return "cuTensorMapEncodeTiled entry point not available";
  const int es = f32 ? 4 : 2;
  const int row_bytes = box_cols * es;
  if (row_bytes != 128 && row_bytes != 64 && row_bytes != 32) return "GEMM tensor-map box row must be 32, 64 or 128 bytes";
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * es) & 15) || (batch > 1 && ((bs * es) & 15))) return "GEMM operand not 16-byte aligned";
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * es, (cuuint64_t)(batch > 1 ? bs : ld * rows) * es};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  // the 16-bit payload is opaque to TMA (no arithmetic, OOB fill is zero bits): fp16 data uses the same map type as bf16
  CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static thread_local char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): K=%lld rows=%lld batch=%lld ld=%lld bs=%lld box_rows=%d", (int)r,
             K, rows, batch, ld, bs, box_rows);
    return buf;
  }
  return nullptr;
}

// 2-D map over a row-major [rows][cols] matrix of 16-bit elements with row stride ld (elements), 128B-swizzled boxes of
// box_cols (= 64) x box_rows; used by the tcgen05 attention of the forecast network (net1_kernels.cu).
const char* encode_tma_2d_16(CUtensorMap* tm, const void* base, long long cols, long long rows, long long ld, int box_cols, int box_rows) {
  return encode_map_t(tm, base, false, cols, rows, 1, ld, 0, box_cols, box_rows);
}

// 3-D map over a [batch][rows][cols] tensor of 16-bit elements (row stride ld, batch stride bs, in elements); the box row decides
// the swizzle: 128 B -> 128B, 64 B -> 64B, 32 B -> none (dense rows).  Used by the fused tower MLP (mlp_fused.cu).
const char* encode_tma_3d_16(CUtensorMap* tm, const void* base, long long cols, long long rows, long long batch, long long ld, long long bs,
                             int box_cols, int box_rows) {
  return encode_map_t(tm, base, false, cols, rows, batch, ld, bs, box_cols, box_rows);
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// 256 x BN tiles, BN a multiple of 64 (tiles may overhang N: TMA zero-fills the loads and clips the stores).  The kernel is
// persistent over num_sms/2 CTA pairs, so the cost of a tiling is its number of waves times the per-tile time, which grows
// with BN; among the candidates pick the one with the least (waves x BN), ties to the larger BN (less operand traffic per MAC).
static int pick_bn(int M, int N, int batch) {
  const char* env = getenv("VV_GEMM_BN");
  if (env && atoi(env) > 0) return atoi(env);
  static const int cand[4] = {256, 192, 128, 64};
  const long long pair_rows = (M + 255) / 256;
  const long long pairs = num_sms() / 2;
  int best = 128; long long best_cost = -1;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    const long long tiles = (long long)((N + bn - 1) / bn) * pair_rows * batch;
    const long long waves = (tiles + pairs - 1) / pairs;
    const long long cost = waves * (bn + 32);       // +32: per-tile fixed cost (pipeline fill, accumulator hand-over)
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

static unsigned long long* g_gemm_trace = nullptr;
static int g_gemm_debug_mode = 0;
void set_gemm_trace(unsigned long long* dev_ptr) { g_gemm_trace = dev_ptr; }
void set_gemm_debug_mode(int mode) { g_gemm_debug_mode = mode; }

static const char* make_gemm_desc_bn(GemmDesc* d, const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb, long long b_bs,
                                     const GemmArgs& args, int force_bn);

// Tile width by measurement: the first time a (shape, epilogue) signature is seen, every candidate BN is timed on the real
// operands (5 launches each, CUDA events) and the fastest is cached.  The cost model of pick_bn cannot see the epilogue's
// instruction cost or the L2 feed limit, which decide the small-K tower GEMMs and the N = 1152 trunk GEMMs.
struct TuneKey {
  int M, N, K, batch, epi, f16, flags;
  bool operator<(const TuneKey& o) const {
    return std::tie(M, N, K, batch, epi, f16, flags) < std::tie(o.M, o.N, o.K, o.batch, o.epi, o.f16, o.flags);
  }
};
static std::map<TuneKey, int> g_tuned;
static bool autotune_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VV_GEMM_AUTOTUNE");      // opt-in: on B200 it confirmed pick_bn's choices (tools: VV_GEMM_AUTOTUNE_LOG=1),
    v = (e && e[0] == '1') ? 1 : 0;                   // and a measured choice would make the kernel selection run-dependent
  }
  return v == 1 && !getenv("VV_GEMM_BN");
}
static int tuned_bn(const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb, long long b_bs, const GemmArgs& args) {
  const int flags = (args.res ? 1 : 0) | (args.out_f32 ? 2 : 0) | (args.out_bf16 ? 4 : 0) | (args.aux_out ? 8 : 0) | (args.aux_in ? 16 : 0) |
                    (args.ln_stats ? 32 : 0) | (args.stats_out ? 64 : 0) | (args.split_n > 0 ? 128 : 0) | (args.bias ? 256 : 0);
  const TuneKey key{args.M, args.N, args.K, args.batch, args.epi, args.f16, flags};
  auto it = g_tuned.find(key);
  if (it != g_tuned.end()) return it->second;
  static const int cand[4] = {64, 128, 192, 256};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int best = 0; float best_ms = 1e30f;
  for (int i = 0; i < 4; ++i) {
    GemmDesc d;
    GemmArgs a2 = args;
    if (a2.stats_out) a2.stats_out_bs = 2LL * 2 * ((a2.N + cand[i] - 1) / cand[i]) * a2.M;
    if (make_gemm_desc_bn(&d, A, lda, a_bs, B, ldb, b_bs, a2, cand[i])) continue;
    d.a.pf_ptr = nullptr; d.a.pf_bytes = 0; d.a.pf2_ptr = nullptr; d.a.pf2_bytes = 0; d.a.trace = nullptr; d.a.debug_mode = 0;
    d.a.stats_out_bs = a2.stats_out_bs;
    launch_gemm(d, 0); launch_gemm(d, 0);
    cudaEventRecord(e0, 0);
    for (int r = 0; r < 5; ++r) launch_gemm(d, 0);
    cudaEventRecord(e1, 0);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaGetLastError(); continue; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best_ms) { best_ms = ms; best = cand[i]; }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (getenv("VV_GEMM_AUTOTUNE_LOG"))
    fprintf(stderr, "[vaevar] gemm %dx%dx%dx%d epi %d f16 %d flags 0x%x -> BN %d (%.1f us)\n", args.M, args.N, args.K, args.batch, args.epi,
            args.f16, flags, best, best_ms * 200.f);
  g_tuned[key] = best;
  return best;
}

const char* make_gemm_desc(GemmDesc* d, const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb, long long b_bs,
                           const GemmArgs& args) {
  int bn = 0;
  if (autotune_enabled() && args.M >= 1024) bn = tuned_bn(A, lda, a_bs, B, ldb, b_bs, args);
  return make_gemm_desc_bn(d, A, lda, a_bs, B, ldb, b_bs, args, bn);
}

static const char* make_gemm_desc_bn(GemmDesc* d, const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb, long long b_bs,
                                     const GemmArgs& args, int force_bn) {
  if (args.N % 8) return "GEMM N must be a multiple of 8";
  if (args.K % 8) return "GEMM K must be a multiple of 8";
  if (args.split_n > 0 && args.split_n % GEMM_EC) return "GEMM split_n must be a multiple of 32";
  if (args.epi == EPI_GELU && args.aux_out && args.out_f32) return "GEMM: fp32 output together with a saved pre-activation is not supported";
  if (args.epi == EPI_DGELU && args.res) return "GEMM: GELU' epilogue with a residual is not supported";
  if (args.epi == EPI_DGELU && (args.ln_stats || args.stats_out)) return "GEMM: GELU' epilogue with LayerNorm folding is not supported";
  if (args.ln_stats && args.stats_out) return "GEMM: a LayerNorm consumer cannot also be a statistics producer";
  if (args.stats_out && args.epi != EPI_LINEAR) return "GEMM: statistics are produced by linear epilogues only";
  if (args.stats_out && args.N % GEMM_EC) return "GEMM: a statistics producer needs N to be a multiple of 32";
  d->a = args;
  const bool b_contig = ldb == args.K && (args.batch == 1 || b_bs == (long long)args.N * args.K);
  d->b_ptr = b_contig ? B : nullptr;
  d->b_bytes = b_contig ? (unsigned long long)args.N * args.K * args.batch * 2 : 0;
  d->a.pf_ptr = nullptr; d->a.pf_bytes = 0; d->a.pf2_ptr = nullptr; d->a.pf2_bytes = 0;
  d->a.trace = g_gemm_trace;
  d->a.debug_mode = g_gemm_debug_mode;
  d->bn = force_bn > 0 ? force_bn : pick_bn(args.M, args.N, args.batch);
  const char* e = encode_map_t(&d->tmA, A, false, args.K, args.M, args.batch, lda, a_bs, GEMM_BK, GEMM_BM);
  if (e) return e;
  e = encode_map_t(&d->tmB, B, false, args.K, args.N, args.batch, ldb, b_bs, GEMM_BK, d->bn / 2);
  if (e) return e;
  memset(&d->sm, 0, sizeof d->sm);
  if (args.out_f32 && (e = encode_map_t(&d->sm.f32, args.out_f32, true, args.N, args.M, args.batch, args.ld_f32, args.f32_bs, GEMM_EC, 32))) return e;
  if (args.out_bf16) {
    if (args.split_n > 0)
      e = encode_map_t(&d->sm.bf16, args.out_bf16, false, args.split_n, args.M, args.N / args.split_n, args.ld_bf16, args.split_stride, GEMM_EC, 32);
    else
      e = encode_map_t(&d->sm.bf16, args.out_bf16, false, args.N, args.M, args.batch, args.ld_bf16, args.bf16_bs, GEMM_EC, 32);
    if (e) return e;
  }
  if (args.epi == EPI_GELU && args.aux_out &&
      (e = encode_map_t(&d->sm.aux, args.aux_out, false, args.N, args.M, args.batch, args.ld_aux, args.aux_bs, GEMM_EC, 32)))
    return e;
  if (args.res && (e = encode_map_t(&d->sm.res, args.res, true, args.N, args.M, args.batch, args.ld_res, args.res_bs, GEMM_EC, 32))) return e;
  if (args.epi == EPI_DGELU && args.aux_in &&
      (e = encode_map_t(&d->sm.aux_in, args.aux_in, false, args.N, args.M, args.batch, args.ld_aux, args.aux_bs, GEMM_EC, 32)))
    return e;
  return nullptr;
}

template <int BN, int STAGES, bool F16, int LNX, int EPI>
static void launch_pair(const GemmDesc& d, cudaStream_t s) {
  using L = GemmSmem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_pair_kernel<BN, STAGES, F16, LNX, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    attr_set = true;
  }
  const long long tiles = (long long)((d.a.N + BN - 1) / BN) * ((d.a.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM)) * d.a.batch;
  int pairs = (int)std::min<long long>(tiles, num_sms() / 2);
  if (const char* cap = getenv("VV_GEMM_MAXPAIRS")) pairs = std::max(1, std::min(pairs, atoi(cap)));   // experiments only
  launch_kernel(gemm_pair_kernel<BN, STAGES, F16, LNX, EPI>, dim3(2 * pairs), dim3(GEMM_THREADS), L::TOTAL, s, d.tmA, d.tmB, d.sm, d.a);
}

template <bool F16, int LNX, int EPI>
static void launch_f(const GemmDesc& d, cudaStream_t s) {
  switch (d.bn) {
    case 64: launch_pair<64, 6, F16, LNX, EPI>(d, s); break;
    case 192: launch_pair<192, 4, F16, LNX, EPI>(d, s); break;
    case 256: launch_pair<256, 4, F16, LNX, EPI>(d, s); break;
    default: launch_pair<128, 5, F16, LNX, EPI>(d, s); break;
  }
}

template <bool F16>
static void launch_e(const GemmDesc& d, cudaStream_t s) {
  // forward pass only: a GEMM either applies a folded LayerNorm (qkv, fc1) or emits the statistics of its output (proj, fc2, seams)
  if (d.a.ln_stats) {
    if (d.a.epi == EPI_GELU) launch_f<F16, LN_CONSUME, EPI_GELU>(d, s);
    else launch_f<F16, LN_CONSUME, EPI_LINEAR>(d, s);
  } else if (d.a.stats_out) {
    launch_f<F16, LN_PRODUCE, EPI_LINEAR>(d, s);
  } else {
    if (d.a.epi == EPI_GELU) launch_f<F16, LN_NONE, EPI_GELU>(d, s);
    else if (d.a.epi == EPI_DGELU) launch_f<F16, LN_NONE, EPI_DGELU>(d, s);
    else launch_f<F16, LN_NONE, EPI_LINEAR>(d, s);
  }
}

void launch_gemm(const GemmDesc& d, cudaStream_t s) {
  if (d.a.f16) launch_e<true>(d, s);
  else launch_e<false>(d, s);
}

}  // namespace vv
