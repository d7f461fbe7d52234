// Host side of the tcgen05 GEMM: TMA descriptor encoding and tile-shape dispatch.
#include <stdio.h>
#include <string.h>

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

// 3-D map over a [batch][rows][K] bf16 tensor, box = (64 x box_rows x 1), 128B swizzle, OOB -> zeros.
static const char* encode_map(CUtensorMap* tm, const bf16* base, long long K, long long rows, long long batch, long long ld,
                              long long bs, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return "cuTensorMapEncodeTiled entry point not available";
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 7) || (batch > 1 && (bs & 7))) return "GEMM operand not 16-byte aligned";
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? bs : ld * rows) * 2};
  cuuint32_t box[3] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static thread_local char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): K=%lld rows=%lld batch=%lld ld=%lld bs=%lld box_rows=%d", (int)r,
             K, rows, batch, ld, bs, box_rows);
    return buf;
  }
  return nullptr;
}

static int pick_bn(int M, int N, int batch) {
  (void)M; (void)batch;
  if (N <= 64) return 64;
  if (N % 128 == 0) return 128;
  if (N % 96 == 0) return 96;
  if (N % 64 == 0 && N < 256) return 64;
  return 128;
}

const char* make_gemm_desc(GemmDesc* d, const bf16* A, long long lda, long long a_bs, const bf16* B, long long ldb, long long b_bs,
                           const GemmArgs& args) {
  if (args.N % 8) return "GEMM N must be a multiple of 8";
  if (args.K % 8) return "GEMM K must be a multiple of 8";
  if (args.split_n > 0 && args.split_n % 8) return "GEMM split_n must be a multiple of 8";
  d->a = args;
  d->bn = pick_bn(args.M, args.N, args.batch);
  const char* e = encode_map(&d->tmA, A, args.K, args.M, args.batch, lda, a_bs, GEMM_BM);
  if (e) return e;
  return encode_map(&d->tmB, B, args.K, args.N, args.batch, ldb, b_bs, d->bn);
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

void launch_gemm(const GemmDesc& d, cudaStream_t s) {
  switch (d.bn) {
    case 64: launch_t<64, 4>(d, s); break;
    case 96: launch_t<96, 3>(d, s); break;
    default: launch_t<128, 3>(d, s); break;
  }
}

}  // namespace vv
