// Batched TN GEMM on the 5th-gen tensor cores:  C[b] = epilogue( A[b] (M x K, K-major bf16) * B[b]^T (N x K, K-major bf16) )
//
// Every Linear of the Swin blocks (swinblock.py:18,20,105,115) and of the U-Net seams
// (transformer.py:73,103,435,552,596) -- forward y = x W^T + b and input-gradient dx = dy W (with W^T packed once) --
// goes through this one kernel.  fp32 accumulation in TMEM, operands staged by TMA into 128B-swizzled shared memory,
// tcgen05.mma issued by a single elected thread, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer
// (+ TMEM owner), warps 2..5 = epilogue (TMEM -> registers -> fused bias / GELU / GELU' / residual -> global).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace vv {

enum GemmEpi : int {
  EPI_LINEAR = 0,  // v = acc + bias
  EPI_GELU = 1,    // u = acc + bias ; aux_out = bf16(u) ; v = gelu(u)
  EPI_DGELU = 2,   // v = acc * gelu'(aux_in)
};

struct GemmArgs {
  int M, N, K, batch;
  int epi;
  const float* bias;          // [batch][N] or null
  long long bias_bs;
  const float* res;           // fp32 residual added to v, [batch][M][ld_res] or null
  long long ld_res, res_bs;
  const __nv_bfloat16* aux_in;  // EPI_DGELU: saved pre-activation u
  __nv_bfloat16* aux_out;       // EPI_GELU: where to save u
  long long ld_aux, aux_bs;
  float* out_f32;             // optional fp32 output
  long long ld_f32, f32_bs;
  __nv_bfloat16* out_bf16;    // optional bf16 output
  long long ld_bf16, bf16_bs;
  int split_n;                // >0: bf16 output column n goes to block n / split_n (stride split_stride), column n % split_n
  long long split_stride;
  int tma_store;              // CTA-pair kernel: 1 = TMA-store epilogue, 0 = direct per-thread stores
};

// Output tensor maps of the TMA-store epilogue (CTA-pair kernel): fp32 boxes are 32 cols x 32 rows, bf16 boxes 64 cols x
// 32 rows, both 128-byte rows with the 128B swizzle.  A map whose pointer in GemmArgs is null is unused.
struct alignas(64) GemmStoreMaps {
  CUtensorMap f32, bf16, aux;
  CUtensorMap res, aux_in;      // persistent kernel: residual (fp32) and saved pre-activation (bf16) are TMA-loaded too
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;  // +1024: manual alignment slack
};

// Epilogue of one 128-row accumulator tile held in this CTA's TMEM: executed by warps 2..5 (warp w may touch TMEM lanes
// [(w%4)*32, +32)); thread = row, 32 fp32 columns per tcgen05.ld, fused bias / GELU / GELU' / residual, direct global stores.
template <int BN>
VV_DEVINL void gemm_epilogue(const GemmArgs& p, uint32_t tmem_base, uint64_t* tmem_full_bar, int warp, int lane, int m0, int n0, int b) {
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int m = m0 + row;
  const bool row_ok = m < p.M;
  mbar_wait(tmem_full_bar, 0);
  tc_fence_after();
  const float* bias = p.bias ? p.bias + (long long)b * p.bias_bs : nullptr;
  const float* res = p.res ? p.res + (long long)b * p.res_bs + (long long)m * p.ld_res : nullptr;
  const __nv_bfloat16* aux_in = p.aux_in ? p.aux_in + (long long)b * p.aux_bs + (long long)m * p.ld_aux : nullptr;
  __nv_bfloat16* aux_out = p.aux_out ? p.aux_out + (long long)b * p.aux_bs + (long long)m * p.ld_aux : nullptr;
  float* of = p.out_f32 ? p.out_f32 + (long long)b * p.f32_bs + (long long)m * p.ld_f32 : nullptr;
  __nv_bfloat16* ob = p.out_bf16 ? p.out_bf16 + (long long)b * p.bf16_bs + (long long)m * p.ld_bf16 : nullptr;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t r[32];
    __syncwarp();
    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {            // 4 groups of 8 columns
      const int n = n0 + c0 + g * 8;
      if (!row_ok || n + 8 > p.N) continue;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
      if (bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (p.epi == EPI_GELU) {
        if (aux_out) {
          uint4 w;
          w.x = pack_bf16(v[0], v[1]); w.y = pack_bf16(v[2], v[3]);
          w.z = pack_bf16(v[4], v[5]); w.w = pack_bf16(v[6], v[7]);
          *reinterpret_cast<uint4*>(aux_out + n) = w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
      } else if (p.epi == EPI_DGELU) {
        const uint4 w = *reinterpret_cast<const uint4*>(aux_in + n);
        const float2 u0 = unpack_bf16(w.x), u1 = unpack_bf16(w.y), u2 = unpack_bf16(w.z), u3 = unpack_bf16(w.w);
        v[0] *= gelu_erf_grad(u0.x); v[1] *= gelu_erf_grad(u0.y);
        v[2] *= gelu_erf_grad(u1.x); v[3] *= gelu_erf_grad(u1.y);
        v[4] *= gelu_erf_grad(u2.x); v[5] *= gelu_erf_grad(u2.y);
        v[6] *= gelu_erf_grad(u3.x); v[7] *= gelu_erf_grad(u3.y);
      }
      if (res) {
        const float4 r0 = *reinterpret_cast<const float4*>(res + n);
        const float4 r1 = *reinterpret_cast<const float4*>(res + n + 4);
        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
        v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
      }
      if (of) {
        *reinterpret_cast<float4*>(of + n) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(of + n + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (ob) {
        uint4 w;
        w.x = pack_bf16(v[0], v[1]); w.y = pack_bf16(v[2], v[3]);
        w.z = pack_bf16(v[4], v[5]); w.w = pack_bf16(v[6], v[7]);
        long long off = n;
        if (p.split_n > 0) off = (long long)(n / p.split_n) * p.split_stride + (n % p.split_n);
        *reinterpret_cast<uint4*>(ob + off) = w;
      }
    }
  }
  }

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  using L = GemmSmem<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * GEMM_BM;
  const int b = blockIdx.z;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        tma_load_3d(sa, &tmA, &full_bar[s], kb * GEMM_BK, m0, b);
        tma_load_3d(sa + L::A_BYTES, &tmB, &full_bar[s], kb * GEMM_BK, n0, b);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t da = make_smem_desc_sw128(sa);
        const uint64_t db = make_smem_desc_sw128(sa + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);          // frees the smem slot once these MMAs retire
      }
      umma_commit(tmem_full_bar);            // accumulator complete
    }
  } else {
    gemm_epilogue<BN>(p, tmem_base, tmem_full_bar, warp, lane, m0, n0, b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// TMA-store epilogue: every epilogue warp owns 32 accumulator rows.  Per 64 output columns it loads the fp32 accumulators
// from TMEM, applies the fused bias / GELU / GELU' / residual, stages fp32 / bf16 / aux results in 128B-swizzled 4 KB slabs
// (the main-loop stage buffers, idle once the accumulator is complete) and lets one elected lane issue bulk tensor
// stores: fully coalesced writes, clipping at the tensor edge for free, stores of slab set i overlap the TMEM load and
// math of set i+1.
template <int BN>
VV_DEVINL void gemm_epilogue_tma(const GemmArgs& p, const GemmStoreMaps& sm, uint8_t* smem, uint32_t tmem_base,
                                 uint64_t* tmem_full_bar, int warp, int lane, int m0, int n0, int b) {
  static_assert(BN % 64 == 0, "TMA-store epilogue works on 64-column slabs");
  const int q = warp & 3;
  const int m = m0 + q * 32 + lane;
  const bool row_ok = m < p.M;
  mbar_wait(tmem_full_bar, 0);
  tc_fence_after();
  uint8_t* slab = smem + q * 16384;                       // 4 slabs of 4 KB per warp
  const uint32_t sw = static_cast<uint32_t>(lane & 7);    // 128B swizzle: 16-byte chunk c of row r lives at chunk c ^ (r & 7)
  const float* bias = p.bias ? p.bias + (long long)b * p.bias_bs : nullptr;
  const float* res = (p.res && row_ok) ? p.res + (long long)b * p.res_bs + (long long)m * p.ld_res : nullptr;
  const __nv_bfloat16* aux_in = (p.aux_in && row_ok) ? p.aux_in + (long long)b * p.aux_bs + (long long)m * p.ld_aux : nullptr;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 64) {
    const int nb = n0 + c0;
    if (nb >= p.N) break;                                  // warp-uniform: nothing of this slab is inside the tensor
    uint32_t r[64];
    __syncwarp();
    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
    tmem_ld_wait();
    uint32_t ub[32];                                       // bf16 copy of the pre-activation (EPI_GELU)
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int n = nb + g * 8;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
      if (n + 8 <= p.N) {
        if (bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n + 4));
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        if (p.epi == EPI_GELU) {
#pragma unroll
          for (int i = 0; i < 4; ++i) ub[g * 4 + i] = pack_bf16(v[2 * i], v[2 * i + 1]);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
        } else if (p.epi == EPI_DGELU) {
          if (aux_in) {
            const uint4 w = *reinterpret_cast<const uint4*>(aux_in + n);
            const float2 u0 = unpack_bf16(w.x), u1 = unpack_bf16(w.y), u2 = unpack_bf16(w.z), u3 = unpack_bf16(w.w);
            v[0] *= gelu_erf_grad(u0.x); v[1] *= gelu_erf_grad(u0.y);
            v[2] *= gelu_erf_grad(u1.x); v[3] *= gelu_erf_grad(u1.y);
            v[4] *= gelu_erf_grad(u2.x); v[5] *= gelu_erf_grad(u2.y);
            v[6] *= gelu_erf_grad(u3.x); v[7] *= gelu_erf_grad(u3.y);
          }
        }
        if (res) {
          const float4 r0 = *reinterpret_cast<const float4*>(res + n);
          const float4 r1 = *reinterpret_cast<const float4*>(res + n + 4);
          v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
          v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) r[g * 8 + i] = __float_as_uint(v[i]);
    }
    // the previous slab set must have been read by the TMA engine before it is overwritten
    if (lane == 0) tma_store_wait_read0();
    __syncwarp();
    const uint32_t rowoff = static_cast<uint32_t>(lane) * 128;
    if (p.out_f32) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {                         // two fp32 slabs of 32 columns
        uint8_t* s = slab + h * 4096 + rowoff;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(s + ((c ^ sw) << 4)) =
              make_uint4(r[h * 32 + c * 4], r[h * 32 + c * 4 + 1], r[h * 32 + c * 4 + 2], r[h * 32 + c * 4 + 3]);
      }
    }
    if (p.out_bf16) {
      uint8_t* s = slab + 2 * 4096 + rowoff;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(r[c * 8]), __uint_as_float(r[c * 8 + 1]));
        w.y = pack_bf16(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
        w.z = pack_bf16(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5]));
        w.w = pack_bf16(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
        *reinterpret_cast<uint4*>(s + ((c ^ sw) << 4)) = w;
      }
    }
    if (p.epi == EPI_GELU && p.aux_out) {
      uint8_t* s = slab + 3 * 4096 + rowoff;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(s + ((c ^ sw) << 4)) = make_uint4(ub[c * 4], ub[c * 4 + 1], ub[c * 4 + 2], ub[c * 4 + 3]);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      const int mrow = m0 + q * 32;
      if (p.out_f32) {
        tma_store_3d(&sm.f32, slab, nb, mrow, b);
        if (nb + 32 < p.N) tma_store_3d(&sm.f32, slab + 4096, nb + 32, mrow, b);
      }
      if (p.out_bf16) {
        if (p.split_n > 0) tma_store_3d(&sm.bf16, slab + 2 * 4096, nb % p.split_n, mrow, nb / p.split_n);
        else tma_store_3d(&sm.bf16, slab + 2 * 4096, nb, mrow, b);
      }
      if (p.epi == EPI_GELU && p.aux_out) tma_store_3d(&sm.aux, slab + 3 * 4096, nb, mrow, b);
      tma_store_commit();
    }
  }
  if (lane == 0) tma_store_wait_read0();
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant: two CTAs of a (2,1,1) cluster compute one 256 x BN tile with tcgen05.mma.cta_group::2.  Each CTA
// stages its own 128 rows of A and HALF of the B tile (BN/2 rows), so the L2 -> shared-memory traffic per MAC drops to
// (128 + BN/2) / (128 * BN) of an operand row (BN = 256: half of the single-CTA 128 x 128 tile).  The leader CTA (rank 0)
// issues the MMAs; both CTAs' TMA loads credit the leader's "full" barrier; tcgen05.commit multicasts the "slot free"
// and "accumulator ready" arrivals to both CTAs; each CTA drains its own 128 accumulator rows.
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct Gemm2Smem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ GemmStoreMaps sm, const GemmArgs p) {
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN");
  using L = Gemm2Smem<BN, STAGES>;
  static_assert(STAGES * L::STAGE_BYTES >= 4 * 16384, "epilogue slabs live in the stage buffers");
  constexpr uint32_t TMEM_COLS = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int n0 = blockIdx.y * BN;
  const int m0 = blockIdx.x * GEMM_BM;              // this CTA's 128 rows (pair = blockIdx.x >> 1)
  const int b = blockIdx.z;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();                               // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);     // bytes of both CTAs land on the leader's barrier
        const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[s]), 0);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        tma_load_3d_2cta(sa, &tmA, leader_full, kb * GEMM_BK, m0, b);
        tma_load_3d_2cta(sa + L::A_BYTES, &tmB, leader_full, kb * GEMM_BK, n0 + (int)rank * (BN / 2), b);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t da = make_smem_desc_sw128(sa);
        const uint64_t db = make_smem_desc_sw128(sa + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_2cta(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        umma_commit_2cta(&empty_bar[s], 3);         // frees this slot in BOTH CTAs
      }
      umma_commit_2cta(tmem_full_bar, 3);           // accumulator complete in both CTAs
    }
  } else {
    if (p.tma_store) gemm_epilogue_tma<BN>(p, sm, smem, tmem_base, tmem_full_bar, warp, lane, m0, n0, b);
    else gemm_epilogue<BN>(p, tmem_base, tmem_full_bar, warp, lane, m0, n0, b);
  }
  tc_fence_before();
  cluster_sync_all();                               // neither CTA may retire while its peer can still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent CTA-pair kernel.  One cluster of two CTAs per SM pair walks over the 256 x BN tiles of the (batched) GEMM:
//   warp 0      TMA producer (both CTAs; STAGES-deep ring that keeps running across tiles)
//   warp 1      tcgen05.mma issuer (leader CTA) and TMEM owner; the fp32 accumulator is DOUBLE-BUFFERED in TMEM
//               (2 x BN columns), so the MMAs of tile i+1 run while tile i is drained
//   warps 2..9  eight epilogue warps: warp w drains TMEM lane quarter (w & 3); the two warps of a quarter take alternate
//               64-column slabs.  Residual / saved pre-activation are TMA-LOADED into the 128B-swizzled slab the result
//               is then written back to in place, and TMA-STORED: every global access of the epilogue is a full-line
//               bulk transfer, none is issued by the LSU.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GEMMP_THREADS = 320;
constexpr int GEMMP_EPI_WARPS = 8;
constexpr int GEMMP_SLAB_BYTES = 3 * 4096;     // per epilogue warp: S0, S1 (fp32 halves / aux-out) and S2 (bf16)

template <int BN, int STAGES>
struct GemmPSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLAB_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = SLAB_OFF + GEMMP_EPI_WARPS * GEMMP_SLAB_BYTES;
  static constexpr int NBAR = 2 * STAGES + 4 + GEMMP_EPI_WARPS;
  static constexpr int TOTAL = BAR_OFF + NBAR * 8 + 16 + 1024;
};

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMMP_THREADS, 1)
gemm_tn_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ GemmStoreMaps io, const GemmArgs p) {
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN");
  using L = GemmPSmem<BN, STAGES>;
  static_assert(L::TOTAL <= 232448, "shared memory budget");
  constexpr uint32_t TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2], leader's copy is the one that counts
  uint64_t* epi_ld_bar = tmem_empty_bar + 2;         // [8]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(epi_ld_bar + GEMMP_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int pair_rows = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int total_tiles = n_tiles * pair_rows * p.batch;
  const int num_pairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 2 * GEMMP_EPI_WARPS);       // 8 epilogue warps in each of the two CTAs
    mbar_init(&tmem_empty_bar[1], 2 * GEMMP_EPI_WARPS);
#pragma unroll
    for (int w = 0; w < GEMMP_EPI_WARPS; ++w) mbar_init(&epi_ld_bar[w], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        const int nt = tile % n_tiles, rest = tile / n_tiles, mp = rest % pair_rows, b = rest / pair_rows;
        const int m0 = (2 * mp + (int)rank) * GEMM_BM, n0 = nt * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[s]), 0);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          tma_load_3d_2cta(sa, &tmA, leader_full, kb * GEMM_BK, m0, b);
          tma_load_3d_2cta(sa + L::A_BYTES, &tmB, leader_full, kb * GEMM_BK, n0, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN);
      uint32_t it = 0, ti = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs, ++ti) {
        const uint32_t acc = ti & 1;
        mbar_wait_cluster(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);       // both CTAs drained this accumulator stage
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_2cta(tacc, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit_2cta(&empty_bar[s], 3);
        }
        umma_commit_2cta(&tmem_full_bar[acc], 3);
      }
    }
  } else {
    // ===== epilogue warps =====
    const int ew = warp - 2;
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    const int half = ew >> 2;                                // which of the two warps of the quarter: alternate 64-col slabs
    uint8_t* S0 = smem + L::SLAB_OFF + ew * GEMMP_SLAB_BYTES;
    uint8_t* S1 = S0 + 4096;
    uint8_t* S2 = S0 + 8192;
    uint64_t* ldbar = &epi_ld_bar[ew];
    uint32_t ld_phase = 0;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const uint32_t rowoff = static_cast<uint32_t>(lane) * 128;
    const bool has_res = p.res != nullptr, has_auxin = (p.epi == EPI_DGELU) && p.aux_in != nullptr;
    const bool has_auxout = (p.epi == EPI_GELU) && p.aux_out != nullptr;
    uint32_t ti = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs, ++ti) {
      const int nt = tile % n_tiles, rest = tile / n_tiles, mp = rest % pair_rows, b = rest / pair_rows;
      const int mrow = (2 * mp + (int)rank) * GEMM_BM + q * 32;        // first of this warp's 32 rows
      const int n0 = nt * BN;
      const uint32_t acc = ti & 1;
      const float* bias = p.bias ? p.bias + (long long)b * p.bias_bs : nullptr;
      mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * 64; c0 < BN; c0 += 128) {
        const int nb = n0 + c0;
        if (nb >= p.N || mrow >= p.M) break;                            // warp-uniform: slab entirely outside the tensor
        const bool second = nb + 32 < p.N;
        // (1) bulk-load the residual / saved pre-activation into the slabs the result will overwrite in place
        if (has_res || has_auxin) {
          if (lane == 0) {
            tma_store_wait_read0();                                     // earlier stores have finished reading the slabs
            uint32_t bytes = 0;
            if (has_res) bytes += second ? 8192u : 4096u;
            if (has_auxin) bytes += 4096u;
            mbar_expect_tx(ldbar, bytes);
            if (has_res) {
              tma_load_3d(S0, &io.res, ldbar, nb, mrow, b);
              if (second) tma_load_3d(S1, &io.res, ldbar, nb + 32, mrow, b);
            }
            if (has_auxin) tma_load_3d(S2, &io.aux_in, ldbar, nb, mrow, b);
          }
        }
        // (2) accumulators: 64 columns of this thread's row
        uint32_t r[64];
        __syncwarp();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c0;
        tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_ld_wait();
        if (has_res || has_auxin) {
          mbar_wait(ldbar, ld_phase);
          ld_phase ^= 1;
        }
        // (3) fused math
        uint32_t ub[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int n = nb + g * 8;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
          if (n + 8 <= p.N) {
            if (bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (p.epi == EPI_GELU) {
#pragma unroll
              for (int i = 0; i < 4; ++i) ub[g * 4 + i] = pack_bf16(v[2 * i], v[2 * i + 1]);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
            } else if (has_auxin) {
              const uint4 w = *reinterpret_cast<const uint4*>(S2 + rowoff + ((static_cast<uint32_t>(g) ^ sw) << 4));
              const float2 u0 = unpack_bf16(w.x), u1 = unpack_bf16(w.y), u2 = unpack_bf16(w.z), u3 = unpack_bf16(w.w);
              v[0] *= gelu_erf_grad(u0.x); v[1] *= gelu_erf_grad(u0.y);
              v[2] *= gelu_erf_grad(u1.x); v[3] *= gelu_erf_grad(u1.y);
              v[4] *= gelu_erf_grad(u2.x); v[5] *= gelu_erf_grad(u2.y);
              v[6] *= gelu_erf_grad(u3.x); v[7] *= gelu_erf_grad(u3.y);
            }
            if (has_res) {
              const uint8_t* rs = (g < 4 ? S0 : S1) + rowoff;
              const uint32_t c = static_cast<uint32_t>((g & 3) * 2);
              const float4 r0 = *reinterpret_cast<const float4*>(rs + ((c ^ sw) << 4));
              const float4 r1 = *reinterpret_cast<const float4*>(rs + (((c + 1) ^ sw) << 4));
              v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
              v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) r[g * 8 + i] = __float_as_uint(v[i]);
        }
        // (4) stage the results (in place over the loaded operands) and bulk-store them
        if (!(has_res || has_auxin)) {
          if (lane == 0) tma_store_wait_read0();
        }
        __syncwarp();
        if (p.out_f32) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint8_t* s = (h ? S1 : S0) + rowoff;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(s + ((static_cast<uint32_t>(c) ^ sw) << 4)) =
                  make_uint4(r[h * 32 + c * 4], r[h * 32 + c * 4 + 1], r[h * 32 + c * 4 + 2], r[h * 32 + c * 4 + 3]);
          }
        }
        if (p.out_bf16) {
          uint8_t* s = S2 + rowoff;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(r[c * 8]), __uint_as_float(r[c * 8 + 1]));
            w.y = pack_bf16(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
            w.z = pack_bf16(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5]));
            w.w = pack_bf16(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
            *reinterpret_cast<uint4*>(s + ((static_cast<uint32_t>(c) ^ sw) << 4)) = w;
          }
        }
        if (has_auxout) {
          uint8_t* s = S0 + rowoff;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(s + ((static_cast<uint32_t>(c) ^ sw) << 4)) = make_uint4(ub[c * 4], ub[c * 4 + 1], ub[c * 4 + 2], ub[c * 4 + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (p.out_f32) {
            tma_store_3d(&io.f32, S0, nb, mrow, b);
            if (second) tma_store_3d(&io.f32, S1, nb + 32, mrow, b);
          }
          if (p.out_bf16) {
            if (p.split_n > 0) tma_store_3d(&io.bf16, S2, nb % p.split_n, mrow, nb / p.split_n);
            else tma_store_3d(&io.bf16, S2, nb, mrow, b);
          }
          if (has_auxout) tma_store_3d(&io.aux, S0, nb, mrow, b);
          tma_store_commit();
        }
      }
      // this warp no longer needs accumulator stage `acc`: tell the leader's MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[acc], 0);
    }
    if (lane == 0) tma_store_wait_read0();
    __syncwarp();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

}  // namespace vv
