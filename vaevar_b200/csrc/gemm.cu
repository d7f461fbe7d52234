// Host side of the tcgen05 GEMM: TMA descriptor encoding and tile-shape dispatch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

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

// 3-D map over a [batch][rows][cols] tensor (bf16 or fp32), box = (box_cols x box_rows x 1) with 128-byte rows,
// 128B swizzle, OOB -> zeros on loads / clipped on stores.
static const char* encode_map_t(CUtensorMap* tm, const void* base, bool f32, long long K, long long rows, long long batch,
                                long long ld, long long bs, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return "cuTensorMapEncodeTiled entry point not available";
  const int es = f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * es) & 15) || (batch > 1 && ((bs * es) & 15))) return "GEMM operand not 16-byte aligned";
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * es, (cuuint64_t)(batch > 1 ? bs : ld * rows) * es};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static thread_local char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): K=%lld rows=%lld batch=%lld ld=%lld bs=%lld box_rows=%d", (int)r,
             K, rows, batch, ld, bs, box_rows);
    return buf;
  }
  return nullptr;
}
static const char* encode_map(CUtensorMap* tm, const bf16* base, long long K, long long rows, long long batch, long long ld,
                              long long bs, int box_rows) {
  return encode_map_t(tm, base, false, K, rows, batch, ld, bs, GEMM_BK, box_rows);
}

static bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '1';
}
static bool use_direct_epilogue() {
  static int v = -1;
  if (v < 0) v = env_flag("VV_GEMM_DIRECT_EPI") ? 1 : 0;
  return v == 1;
}

static bool use_nonpersistent() {
  static int v = -1;
  if (v < 0) v = env_flag("VV_GEMM_NONPERSISTENT") ? 1 : 0;
  return v == 1;
}
static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static bool use_1cta() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VV_GEMM_1CTA"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

// Single-CTA kernel (fallback / A-B comparison): 128 x BN tiles.
static int pick_bn_1cta(int N) {
  if (N <= 64) return 64;
  if (N % 128 == 0) return 128;
  if (N % 96 == 0) return 96;
  if (N % 64 == 0 && N < 256) return 64;
  return 128;
}
// CTA-pair kernel: 256 x BN tiles, BN a multiple of 64 (the TMA-store epilogue works on 64-column slabs; tiles may overhang N,
// TMA zero-fills the loads and clips the stores).  Least column padding first; then the largest BN that still gives about
// one CTA per SM, otherwise the BN with the most CTAs.
static int pick_bn_2cta(int M, int N, int batch) {
  static const int cand[4] = {256, 192, 128, 64};
  const long long pair_rows = (M + 255) / 256;
  int min_waste = 1 << 30;
  for (int i = 0; i < 4; ++i) min_waste = std::min(min_waste, (N + cand[i] - 1) / cand[i] * cand[i] - N);
  int best = 0; long long best_ctas = -1;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    const int tiles = (N + bn - 1) / bn;
    if (tiles * bn - N != min_waste) continue;
    const long long ctas = 2 * pair_rows * tiles * batch;
    if (ctas >= 140) return bn;
    if (ctas > best_ctas) { best_ctas = ctas; best = bn; }
  }
  return best ? best : 128;
}

const char* make_gemm_desc(GemmDesc* d, const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb, long long b_bs,
                           const GemmArgs& args) {
  if (args.N % 8) return "GEMM N must be a multiple of 8";
  if (args.K % 8) return "GEMM K must be a multiple of 8";
  if (args.split_n > 0 && args.split_n % 8) return "GEMM split_n must be a multiple of 8";
  d->a = args;
  d->two_cta = use_1cta() ? 0 : 1;
  d->bn = d->two_cta ? pick_bn_2cta(args.M, args.N, args.batch) : pick_bn_1cta(args.N);
  d->a.tma_store = (d->two_cta && !use_direct_epilogue()) ? 1 : 0;
  d->persist = 0;
  if (args.split_n > 0 && args.split_n % 64) d->a.tma_store = 0;
  const char* e = encode_map(&d->tmA, A, args.K, args.M, args.batch, lda, a_bs, GEMM_BM);
  if (e) return e;
  e = encode_map(&d->tmB, B, args.K, args.N, args.batch, ldb, b_bs, d->two_cta ? d->bn / 2 : d->bn);
  if (e) return e;
  memset(&d->sm, 0, sizeof d->sm);
  if (d->a.tma_store) {
    if (args.out_f32 && (e = encode_map_t(&d->sm.f32, args.out_f32, true, args.N, args.M, args.batch, args.ld_f32, args.f32_bs, 32, 32))) return e;
    if (args.out_bf16) {
      if (args.split_n > 0)
        e = encode_map_t(&d->sm.bf16, args.out_bf16, false, args.split_n, args.M, args.N / args.split_n, args.ld_bf16, args.split_stride, 64, 32);
      else
        e = encode_map_t(&d->sm.bf16, args.out_bf16, false, args.N, args.M, args.batch, args.ld_bf16, args.bf16_bs, 64, 32);
      if (e) return e;
    }
    if (args.epi == EPI_GELU && args.aux_out &&
        (e = encode_map_t(&d->sm.aux, args.aux_out, false, args.N, args.M, args.batch, args.ld_aux, args.aux_bs, 64, 32)))
      return e;
    d->persist = use_nonpersistent() ? 0 : 1;
    if (d->persist) {
      if (args.res && (e = encode_map_t(&d->sm.res, args.res, true, args.N, args.M, args.batch, args.ld_res, args.res_bs, 32, 32))) return e;
      if (args.epi == EPI_DGELU && args.aux_in &&
          (e = encode_map_t(&d->sm.aux_in, args.aux_in, false, args.N, args.M, args.batch, args.ld_aux, args.aux_bs, 64, 32)))
        return e;
    }
  }
  return nullptr;
}

template <int BN, int STAGES>
static void launch_t(const GemmDesc& d, cudaStream_t s) {
  using L = GemmSmem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_tn_tcgen05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    attr_set = true;
  }
  dim3 grid((d.a.N + BN - 1) / BN, (d.a.M + GEMM_BM - 1) / GEMM_BM, d.a.batch);
  gemm_tn_tcgen05_kernel<BN, STAGES><<<grid, GEMM_THREADS, L::TOTAL, s>>>(d.tmA, d.tmB, d.a);
}

template <int BN, int STAGES>
static void launch_2cta(const GemmDesc& d, cudaStream_t s) {
  using L = Gemm2Smem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_tn_2cta_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    attr_set = true;
  }
  const int mtiles = (d.a.M + GEMM_BM - 1) / GEMM_BM;
  dim3 grid(((mtiles + 1) / 2) * 2, (d.a.N + BN - 1) / BN, d.a.batch);     // grid.x even: CTA pairs along M
  gemm_tn_2cta_kernel<BN, STAGES><<<grid, GEMM_THREADS, L::TOTAL, s>>>(d.tmA, d.tmB, d.sm, d.a);
}

template <int BN, int STAGES>
static void launch_persist(const GemmDesc& d, cudaStream_t s) {
  using L = GemmPSmem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_tn_persist_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    attr_set = true;
  }
  const long long tiles = (long long)((d.a.N + BN - 1) / BN) * ((d.a.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM)) * d.a.batch;
  const int pairs = (int)std::min<long long>(tiles, num_sms() / 2);
  gemm_tn_persist_kernel<BN, STAGES><<<2 * pairs, GEMMP_THREADS, L::TOTAL, s>>>(d.tmA, d.tmB, d.sm, d.a);
}

void launch_gemm(const GemmDesc& d, cudaStream_t s) {
  if (d.persist) {
    switch (d.bn) {
      case 64: launch_persist<64, 6>(d, s); break;
      case 192: launch_persist<192, 4>(d, s); break;
      case 256: launch_persist<256, 4>(d, s); break;
      default: launch_persist<128, 5>(d, s); break;
    }
    return;
  }
  if (d.two_cta) {
    switch (d.bn) {
      case 64: launch_2cta<64, 4>(d, s); break;
      case 192: launch_2cta<192, 3>(d, s); break;
      case 256: launch_2cta<256, 3>(d, s); break;
      default: launch_2cta<128, 4>(d, s); break;
    }
    return;
  }
  switch (d.bn) {
    case 64: launch_t<64, 4>(d, s); break;
    case 96: launch_t<96, 3>(d, s); break;
    default: launch_t<128, 3>(d, s); break;
  }
}

}  // namespace vv
