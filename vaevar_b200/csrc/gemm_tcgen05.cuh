// Batched TN GEMM on the 5th-gen tensor cores:  C[b] = epilogue( A[b] (M x K, K-major) * B[b]^T (N x K, K-major) )
//
// Every Linear of the Swin blocks (swinblock.py:18,20,105,115) and of the U-Net seams
// (transformer.py:73,103,435,552,596) -- forward y = x W^T + b and input-gradient dx = dy W (with W^T packed once) --
// goes through this one kernel.  16-bit operands (IEEE fp16 in the forward pass, bf16 for gradients), fp32 accumulation
// in TMEM, operands staged by TMA into 128B-swizzled shared memory, tcgen05.mma.cta_group::2 issued by a single thread.
//
// Persistent CTA-pair kernel.  One cluster of two CTAs per SM pair walks over the 256 x BN tiles of the (batched) GEMM:
//   warp 0      TMA producer (both CTAs; STAGES-deep ring that keeps running across tiles).  Each CTA stages its own
//               128 rows of A and HALF of the B tile, so the L2 -> shared-memory traffic per MAC is that of a 256 x BN tile.
//   warp 1      tcgen05.mma issuer (leader CTA) and TMEM owner; the fp32 accumulator is DOUBLE-BUFFERED in TMEM
//               (2 x BN columns), so the MMAs of tile i+1 run while tile i is drained
//   warps 2..9  eight epilogue warps: warp w drains TMEM lane quarter (w & 3); the two warps of a quarter take alternate
//               32-column chunks.  Per chunk: tcgen05.ld -> registers -> fused bias / GELU / GELU' / residual -> 128B- or
//               64B-swizzled staging slab -> TMA store.  The residual and the saved gelu'(u) are TMA-LOADED into the
//               slab the result then overwrites in place; slabs are double-buffered and the loads of chunk i+1 are issued
//               while chunk i is processed, so no global access of the epilogue is issued by the LSU except the bias.
//
// Producer and MMA warps run warp-uniform loops with the TMA / tcgen05 instructions under elect.sync (a divergent lane-0
// branch costs an ELECT + R2UR waterfall per instruction: 170 cycles per MMA).  The epilogue arithmetic is packed fp32
// (FFMA2 / FMUL2 / FADD2).  Template parameters beyond the tile shape: F16 (operand format), LNX (plain / folds a LayerNorm
// into the epilogue / emits LayerNorm statistics of its output), EPI (epilogue variant) -- one variant per instantiation keeps
// each kernel inside the instruction cache.  The idle producer warp prefetches the NEXT GEMM's weights into L2.
// Measurements behind these choices: DESIGN.md section 4; tools/gemm_trace.py, tools/gemm_boundary.py.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace vv {

enum GemmLn : int { LN_NONE = 0, LN_CONSUME = 1, LN_PRODUCE = 2 };

enum GemmEpi : int {
  EPI_LINEAR = 0,  // v = acc + bias
  EPI_GELU = 1,    // u = acc + bias ; aux_out = 16bit(gelu'(u)) ; v = gelu(u)
  EPI_DGELU = 2,   // v = acc * aux_in   (aux_in = the gelu'(u) the forward GEMM saved)
};

struct GemmArgs {
  int M, N, K, batch;
  int epi;
  int f16;                    // 1: A, B and every 16-bit output are IEEE fp16 (forward pass); 0: bfloat16 (gradients)
  int aux_f16;                // format of aux_in (the forward pass saves the pre-activation as fp16 when it runs in fp16)
  const float* bias;          // [batch][N] or null
  long long bias_bs;
  const float* res;           // fp32 residual added to v, [batch][M][ld_res] or null
  long long ld_res, res_bs;
  const __nv_bfloat16* aux_in;  // EPI_DGELU: saved gelu'(u) (16-bit storage, format aux_f16)
  __nv_bfloat16* aux_out;       // EPI_GELU: where to save gelu'(u) (16-bit storage, format f16)
  long long ld_aux, aux_bs;
  float* out_f32;             // optional fp32 output
  long long ld_f32, f32_bs;
  __nv_bfloat16* out_bf16;    // optional 16-bit output (storage type only: the format follows `f16`)
  long long ld_bf16, bf16_bs;
  int split_n;                // >0: 16-bit output column n goes to block n / split_n (stride split_stride), column n % split_n
  long long split_stride;
  // LayerNorm folded into the GEMM (forward pass).  With W' = W o gamma, s_n = sum_k W'[n,k], c_n = b_n + sum_k beta_k W[n,k]:
  //   LN(x) W^T + b = rstd_r (x W'^T - mean_r s_n) + c_n
  // so the GEMM runs on a 16-bit copy of x and the epilogue applies the per-row (mean, rstd) it derives from the statistics
  // partials a PRODUCER GEMM (or the stats kernel) left in ln_stats; c_n travels as `bias`.
  // Numerics (what makes the fold safe for residual streams with |mean| >> sigma):
  //  * the 16-bit copy holds x - shift_r, shift_r = the row's mean at the input of the stage (ln_shift, written by the stats
  //    kernel), so the operand rounding error is relative to the row's spread, not to its offset; the consumer then applies
  //    mean_r - shift_r;
  //  * a partial is (mean_p, M2_p) = (mean, sum of squared deviations from it) of the columns it covers, accumulated around
  //    the partial's first value and combined with Chan's pairwise update -- no E[x^2] - mean^2 cancellation anywhere;
  //  * rows whose remaining offset or magnitude still endangers the 16-bit operand are counted in ln_health.
  const float* ln_stats;      // consumer: [batch][ln_parts][M] float2 (mean_p, M2_p) partials of the A rows; null = plain GEMM
  int ln_parts;
  int ln_prod_bn;             // consumer: tile width of the GEMM that produced the partials (partial 2 nt + h covers the 32-column
                              // chunks h, h + 2, ... of n-tile nt); 0 = one partial over the whole row (stats kernel)
  int ln_c;                   // consumer: row length C of the normalised rows
  long long ln_stats_bs;      // batch stride in floats
  const float* ln_colsum;     // consumer: s_n, [batch][N] (batch stride = bias_bs)
  float ln_inv_c, ln_eps;     // 1 / C, epsilon
  const float* ln_shift;      // consumer and producer: per-row shift [batch][M] (batch stride M); null = 0
  unsigned int* ln_health;    // consumer: [0] += rows with |mean - shift| > 32 sigma, [1] += rows near fp16 saturation; null = off
  float* stats_out;           // producer: [batch][2 * n_tiles][M] float2 partials of the rows of the fp32 output; null = none
  long long stats_out_bs;     // batch stride in floats
  const void* pf_ptr;         // weights of the NEXT GEMM in the plan: prefetched into L2 by the idle producer warp (null = none)
  unsigned long long pf_bytes;
  const void* pf2_ptr;        // a second region to prefetch (the weights of a fused tower MLP launched next); null = none
  unsigned long long pf2_bytes;
  int debug_mode;             // debug: 0 normal; 1 MMA only (no TMA, operands = whatever is in smem); 2 TMA only (no MMA)
  unsigned long long* trace;  // debug: per-CTA clock64 stamps (64 slots per CTA, see tools/gemm_trace.py); null in production
};

// Tensor maps of the epilogue: fp32 boxes are 32 cols x 32 rows (128-byte rows, 128B swizzle), 16-bit boxes 32 cols x 32 rows
// (64-byte rows, 64B swizzle).  A map whose pointer in GemmArgs is null is unused.
struct alignas(64) GemmStoreMaps {
  CUtensorMap f32, bf16, aux;   // stores
  CUtensorMap res, aux_in;      // loads
};

#ifndef VV_EPI_GW
#define VV_EPI_GW 8                 // columns of an epilogue chunk processed together (8, 16 or 32)
#endif
constexpr int GEMM_BM = 128;        // rows per CTA (256 per pair)
constexpr int GEMM_BK = 64;         // K elements per pipeline stage = one 128-byte swizzle row
constexpr int GEMM_THREADS = 320;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_EC = 32;         // epilogue chunk: columns per tcgen05.ld / staging slab
constexpr int GEMM_SLOT_A = 4096;   // fp32 chunk (32 rows x 128 B) -- residual in, fp32 out; or the saved pre-activation out
constexpr int GEMM_SLOT_B = 2048;   // 16-bit chunk (32 rows x 64 B) -- pre-activation in, 16-bit out
constexpr int GEMM_BUF = GEMM_SLOT_A + GEMM_SLOT_B;
constexpr int GEMM_WARP_SLAB = 2 * GEMM_BUF;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLAB_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = SLAB_OFF + GEMM_EPI_WARPS * GEMM_WARP_SLAB;
  static constexpr int NBAR = 2 * STAGES + 4 + 2 * GEMM_EPI_WARPS;
  static constexpr int TOTAL = BAR_OFF + NBAR * 8 + 16 + 1024;   // +1024: manual alignment slack
};

// Two fp32 -> packed 16-bit pair (a in the low half).  fp16 conversions saturate to +-65504 instead of producing inf: an
// activation that leaves fp16's range (a line-search trial far outside the basin) then yields a large finite J the line
// search backs away from, as the fp32 reference would, instead of NaN.
template <bool F16>
VV_DEVINL uint32_t pack16(float a, float b) {
  if (F16) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
  return pack_bf16(a, b);
}
VV_DEVINL float2 unpack16(uint32_t w, bool f16) {
  if (f16) return __half22float2(*reinterpret_cast<__half2*>(&w));
  return unpack_bf16(w);
}

// One 32-row x 32-column chunk of the epilogue, executed by one warp (thread = row).  r[] holds the fp32 accumulators,
// bv[] the bias of the 32 columns; slot A / slot B are this warp's staging slabs (see GEMM_SLOT_*).
// ln_a scales the accumulator (the row's rstd when a LayerNorm is folded into this GEMM, else 1).  When this GEMM produces
// LayerNorm statistics for its consumer, rs / rq accumulate sum (v - piv) and sum (v - piv)^2 of the final fp32 values around
// the pivot `piv` (the first value this thread saw in the tile; set here when `first`), and the 16-bit copy holds v - shift.
template <int EPI, bool F16, int LNX>
VV_DEVINL void epilogue_chunk(const GemmArgs& p, uint32_t (&r)[32], const float (&bv)[32], uint8_t* SA, uint8_t* SB, int lane,
                              bool has_res, bool has_auxin, bool has_auxout, float ln_a, float& rs, float& rq, float& piv,
                              bool first, float shift) {
  const uint32_t sw128 = static_cast<uint32_t>(lane & 7);          // 128B swizzle: 16-byte chunk c of row r lives at c ^ (r & 7)
  const uint32_t sw64 = static_cast<uint32_t>((lane >> 1) & 3);    // 64B swizzle: chunk c of row r lives at c ^ ((r >> 1) & 3)
  uint8_t* rowA = SA + lane * 128;
  uint8_t* rowA16 = SA + lane * 64;                                // slot A used as a 16-bit slab (saved pre-activation)
  uint8_t* rowB = SB + lane * 64;
  // GW columns are in flight together: the dependent chains of the epilogue math (GELU: ~10 deep, two MUFU each) are hidden by
  // the GW / 2 independent column pairs of a group, so wider groups stall less on fixed-latency dependencies (VV_EPI_GW).
  constexpr int GW = VV_EPI_GW, NS = GW / 8;
#pragma unroll
  for (int g0 = 0; g0 < 32 / GW; ++g0) {
    float v[GW];
#pragma unroll
    for (int i = 0; i < GW; i += 2) {                                // packed fp32: two columns per instruction
      const float2 a2 = make_float2(__uint_as_float(r[g0 * GW + i]), __uint_as_float(r[g0 * GW + i + 1]));
      const float2 b2 = make_float2(bv[g0 * GW + i], bv[g0 * GW + i + 1]);
      const float2 o2 = LNX == LN_CONSUME ? fma2(splat2(ln_a), a2, b2) : add2(a2, b2);
      v[i] = o2.x; v[i + 1] = o2.y;
    }
    uint4 u16[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) u16[s] = make_uint4(0, 0, 0, 0);
    if (EPI == EPI_GELU) {
      if (has_auxout) {                      // gelu and gelu' share their transcendental work; the backward pass needs only gelu'(u)
        float d[GW];
#pragma unroll
        for (int i = 0; i < GW; i += 2) {
          float2 y2, d2;
#ifdef VV_EXP_NOGELU
          y2 = make_float2(v[i], v[i + 1]); d2 = y2;
#else
          gelu_erf_both2(make_float2(v[i], v[i + 1]), &y2, &d2);
#endif
          v[i] = y2.x; v[i + 1] = y2.y; d[i] = d2.x; d[i + 1] = d2.y;
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          u16[s].x = pack16<F16>(d[8 * s], d[8 * s + 1]); u16[s].y = pack16<F16>(d[8 * s + 2], d[8 * s + 3]);
          u16[s].z = pack16<F16>(d[8 * s + 4], d[8 * s + 5]); u16[s].w = pack16<F16>(d[8 * s + 6], d[8 * s + 7]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < GW; i += 2) {
          const float2 y2 = gelu_erf2(make_float2(v[i], v[i + 1]));
          v[i] = y2.x; v[i + 1] = y2.y;
        }
      }
    } else if (EPI == EPI_DGELU) {
      if (has_auxin) {                       // aux = gelu'(u) saved by the forward GEMM
        const bool af = p.aux_f16 != 0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const int g = g0 * NS + s;
          const uint4 w = *reinterpret_cast<const uint4*>(rowB + ((static_cast<uint32_t>(g) ^ sw64) << 4));
          const float2 u0 = unpack16(w.x, af), u1 = unpack16(w.y, af), u2 = unpack16(w.z, af), u3 = unpack16(w.w, af);
          float* vs = v + 8 * s;
          const float2 m0 = mul2(make_float2(vs[0], vs[1]), u0), m1 = mul2(make_float2(vs[2], vs[3]), u1);
          const float2 m2 = mul2(make_float2(vs[4], vs[5]), u2), m3 = mul2(make_float2(vs[6], vs[7]), u3);
          vs[0] = m0.x; vs[1] = m0.y; vs[2] = m1.x; vs[3] = m1.y; vs[4] = m2.x; vs[5] = m2.y; vs[6] = m3.x; vs[7] = m3.y;
        }
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int g = g0 * NS + s;                                     // 8-column subgroup = one 16-byte chunk of the 16-bit slabs
      float* vs = v + 8 * s;
      const uint32_t cA0 = (static_cast<uint32_t>(2 * g) ^ sw128) << 4, cA1 = (static_cast<uint32_t>(2 * g + 1) ^ sw128) << 4;
      if (has_res) {
        const float4 r0 = *reinterpret_cast<const float4*>(rowA + cA0);
        const float4 r1 = *reinterpret_cast<const float4*>(rowA + cA1);
        vs[0] += r0.x; vs[1] += r0.y; vs[2] += r0.z; vs[3] += r0.w;
        vs[4] += r1.x; vs[5] += r1.y; vs[6] += r1.z; vs[7] += r1.w;
      }
      if (LNX == LN_PRODUCE && p.stats_out) {   // deviations from the pivot; two independent partial chains per statistic
        if (first && g == 0) piv = vs[0];
        float dv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dv[i] = vs[i] - piv;
        rs += (dv[0] + dv[1]) + (dv[2] + dv[3]) + ((dv[4] + dv[5]) + (dv[6] + dv[7]));
        rq += fmaf(dv[0], dv[0], dv[1] * dv[1]) + fmaf(dv[2], dv[2], dv[3] * dv[3]) + (fmaf(dv[4], dv[4], dv[5] * dv[5]) + fmaf(dv[6], dv[6], dv[7] * dv[7]));
      }
      if (p.out_f32) {
        *reinterpret_cast<float4*>(rowA + cA0) = make_float4(vs[0], vs[1], vs[2], vs[3]);
        *reinterpret_cast<float4*>(rowA + cA1) = make_float4(vs[4], vs[5], vs[6], vs[7]);
      }
      if (p.out_bf16) {
        if (LNX == LN_PRODUCE) {                 // the copy a folded LayerNorm consumes is centred on the row's shift
#pragma unroll
          for (int i = 0; i < 8; ++i) vs[i] -= shift;
        }
        uint4 w;
        w.x = pack16<F16>(vs[0], vs[1]); w.y = pack16<F16>(vs[2], vs[3]);
        w.z = pack16<F16>(vs[4], vs[5]); w.w = pack16<F16>(vs[6], vs[7]);
        *reinterpret_cast<uint4*>(rowB + ((static_cast<uint32_t>(g) ^ sw64) << 4)) = w;
      }
      if (EPI == EPI_GELU && has_auxout) *reinterpret_cast<uint4*>(rowA16 + ((static_cast<uint32_t>(g) ^ sw64) << 4)) = u16[s];
    }
  }
}

// LNX: LN_CONSUME = the instantiation that folds a LayerNorm into the epilogue (ln_stats), LN_PRODUCE = the one that emits the statistics of
// its fp32 output (producer: stats_out); the plain instantiation carries none of that code.
// EPI is a compile-time parameter as well: one epilogue variant per instantiation keeps the kernel around 40 KB of SASS (the
// three roles' code sits far apart and the all-variants kernel overflowed the instruction cache: "no instruction" stalls).
template <int BN, int STAGES, bool F16, int LNX, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ GemmStoreMaps io, const GemmArgs p) {
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN");
  using L = GemmSmem<BN, STAGES>;
  static_assert(L::TOTAL <= 232448, "shared memory budget");
  constexpr uint32_t TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 64 + 11] = globaltimer_ns();     // kernel entry
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2], the leader's copy is the one that counts
  uint64_t* epi_ld_bar = tmem_empty_bar + 2;         // [8 warps][2 buffers]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(epi_ld_bar + 2 * GEMM_EPI_WARPS);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);     // warp-uniform for the compiler (uniform datapath)
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
    mbar_init(&tmem_empty_bar[0], 2 * GEMM_EPI_WARPS);       // 8 epilogue warps in each of the two CTAs
    mbar_init(&tmem_empty_bar[1], 2 * GEMM_EPI_WARPS);
#pragma unroll
    for (int w = 0; w < 2 * GEMM_EPI_WARPS; ++w) mbar_init(&epi_ld_bar[w], 1);
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
  // Programmatic dependent launch: everything above overlaps the tail of the previous kernel in the stream; nothing below
  // may touch global memory before the previous kernel has completed.
  if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 64 + 12] = globaltimer_ns();     // prologue done
  pdl_launch_dependents();
  pdl_wait();
  unsigned long long* trc = p.trace ? p.trace + (size_t)blockIdx.x * 64 : nullptr;
  if (trc && threadIdx.x == 0) { trc[0] = clock64(); trc[10] = globaltimer_ns(); }

  // The producer and the MMA warp run their loops WARP-UNIFORMLY (all 32 lanes take the same path; one elected lane issues
  // the TMA / tcgen05 instructions): descriptors, coordinates and barrier addresses then live in uniform registers.  Under a
  // divergent `lane == 0` branch the compiler has to wrap every TMA / MMA instruction in an ELECT + R2UR waterfall loop, which
  // alone costs ~170 cycles per tcgen05.mma (measured) -- more than the MMA itself.
  if (warp == 0) {
    // ===== TMA producer =====
    if (p.debug_mode != 1) {
      uint32_t it = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        const int nt = tile % n_tiles, rest = tile / n_tiles, mp = rest % pair_rows, b = rest / pair_rows;
        const int m0 = (2 * mp + (int)rank) * GEMM_BM, n0 = nt * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (trc && it < 24 && lane == 0) trc[16 + it] = clock64();
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[s]), 0);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);   // bytes of both CTAs land on the leader's barrier
            tma_load_3d_2cta(sa, &tmA, leader_full, kb * GEMM_BK, m0, b);
            tma_load_3d_2cta(sa + L::A_BYTES, &tmB, leader_full, kb * GEMM_BK, n0, b);
          }
          __syncwarp();
        }
      }
    }
    // All operand loads of this CTA are in flight: use the idle producer warp to pull this CTA's slice of the NEXT GEMM's
    // weights from HBM into L2 (the 0.43 GB of weights per network never stay in the 126 MB L2 from one evaluation to the
    // next, so without this every GEMM starts on a DRAM-latency-bound pipeline fill).
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
      const void* pf = r ? p.pf2_ptr : p.pf_ptr;
      const unsigned long long pfb = r ? p.pf2_bytes : p.pf_bytes;
      if (!pf) continue;
      constexpr unsigned long long CH = 4096;
      const unsigned long long per = ((pfb + gridDim.x - 1) / gridDim.x + CH - 1) / CH * CH;
      const unsigned long long beg = (unsigned long long)blockIdx.x * per;
      const unsigned long long end = beg + per < pfb ? beg + per : pfb;
      for (unsigned long long o = beg + lane * CH; o < end; o += 32 * CH) {
        const unsigned long long n = end - o < CH ? (end - o) & ~15ull : CH;
        if (n) l2_prefetch_bulk(static_cast<const uint8_t*>(pf) + o, (uint32_t)n);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    if (rank == 0) {
      const uint32_t idesc = make_idesc_16(2 * GEMM_BM, BN, F16);
      const uint32_t smem_base = smem_u32(smem);
      uint32_t it = 0, ti = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs, ++ti) {
        const uint32_t acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);          // both CTAs drained this accumulator stage
        tc_fence_after();
        if (trc && ti < 2 && lane == 0) trc[2 + ti] = clock64();
        const uint32_t tacc = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (p.debug_mode != 1) mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (trc && it < 24 && lane == 0) trc[40 + it] = clock64();
          if (p.debug_mode == 2) {                                       // feed only: hand the slot straight back
            if (elect_one()) {
              mbar_arrive(&empty_bar[s]);
              mbar_arrive_cluster(&empty_bar[s], 1);
            }
            __syncwarp();
            continue;
          }
          const uint32_t sa = smem_base + s * L::STAGE_BYTES;
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + L::A_BYTES);
          const int krem = p.K - kb * GEMM_BK;                           // a short last block issues only the MMAs it needs
          const int nk = krem >= GEMM_BK ? GEMM_BK / 16 : (krem + 15) / 16;
          if (elect_one()) {
            // advance 16 elements = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            if (nk == GEMM_BK / 16) {
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k) umma_16_2cta(tacc, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
            } else {
              for (int k = 0; k < nk; ++k) umma_16_2cta(tacc, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
            }
            umma_commit_2cta(&empty_bar[s], 3);       // frees this slot in BOTH CTAs once these MMAs retire
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_2cta(&tmem_full_bar[acc], 3);     // accumulator complete in both CTAs
        __syncwarp();
        if (trc && ti < 2 && lane == 0) trc[4 + ti] = clock64();
      }
    }
  } else {
    // ===== epilogue warps =====
    const int ew = warp - 2;
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    const int half = ew >> 2;                                // which of the two warps of the quarter: alternate 32-col chunks
    uint8_t* slab = smem + L::SLAB_OFF + ew * GEMM_WARP_SLAB;
    uint64_t* ldbar = &epi_ld_bar[2 * ew];
    uint32_t ld_phase = 0;                                   // bit b: phase of ldbar[b]
    const bool has_res = p.res != nullptr, has_auxin = (EPI == EPI_DGELU) && p.aux_in != nullptr;
    const bool has_auxout = (EPI == EPI_GELU) && p.aux_out != nullptr;
    const bool has_loads = has_res || has_auxin;
    const uint32_t load_bytes = (has_res ? GEMM_SLOT_A : 0) + (has_auxin ? GEMM_SLOT_B : 0);

    // Tile coordinates advance incrementally (the persistent stride num_pairs decomposed once into (nt, mp, b) steps): no
    // integer divisions per tile -- the small-K tower GEMMs run ~13 tiles of ~1 us per CTA, where they were a visible cost.
    struct Coord { int tile, nt, mp, b; };
    const int step_nt = num_pairs % n_tiles, step_mp = (num_pairs / n_tiles) % pair_rows, step_b = num_pairs / (n_tiles * pair_rows);
    auto advance = [&](Coord& c) {
      c.tile += num_pairs;
      c.nt += step_nt;
      int carry = 0;
      if (c.nt >= n_tiles) { c.nt -= n_tiles; carry = 1; }
      c.mp += step_mp + carry;
      carry = 0;
      if (c.mp >= pair_rows) { c.mp -= pair_rows; carry = 1; }
      c.b += step_b + carry;
    };
    auto row0_of = [&](const Coord& c) { return (2 * c.mp + (int)rank) * GEMM_BM + q * 32; };
    // number of 32-column chunks of the tile this warp's rows take part in
    auto nch_of = [&](const Coord& c) {
      const int ncol = min(BN, p.N - c.nt * BN);
      return row0_of(c) < p.M ? (ncol + GEMM_EC - 1) / GEMM_EC : 0;
    };
    auto issue_loads = [&](int col, int mrow, int b, int buf) {        // the elected lane only
      tma_store_wait_read0();                                          // the store that last read this buffer is done with it
      mbar_expect_tx(&ldbar[buf], load_bytes);
      if (has_res) tma_load_3d(slab + buf * GEMM_BUF, &io.res, &ldbar[buf], col, mrow, b);
      if (has_auxin) tma_load_3d(slab + buf * GEMM_BUF + GEMM_SLOT_A, &io.aux_in, &ldbar[buf], col, mrow, b);
    };

    // folded LayerNorm: the first six (mean, M2) partials of this thread's row and the row's shift, fetched one tile ahead (a
    // dependent global round trip per tile would dominate the small-K tower GEMMs)
    float2 stat_nx[6];
    float shift_nx = 0.f;
    const float2* stats2 = reinterpret_cast<const float2*>(p.ln_stats);
    const long long stats_bs2 = p.ln_stats_bs >> 1;
    auto load_stats = [&](const Coord& c) {
#pragma unroll
      for (int j = 0; j < 6; ++j) stat_nx[j] = make_float2(0.f, 0.f);
      shift_nx = 0.f;
      if (c.tile >= total_tiles) return;
      const int row = row0_of(c) + lane;
      if (row >= p.M) return;
      if (p.ln_shift) shift_nx = __ldg(p.ln_shift + (long long)c.b * p.M + row);
      if (LNX != LN_CONSUME) return;
      const float2* st = stats2 + c.b * stats_bs2 + row;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (j < p.ln_parts) stat_nx[j] = __ldg(st + j * p.M);
    };
    // columns covered by partial j (see GemmArgs::ln_prod_bn)
    auto part_cols = [&](int j) {
      if (p.ln_prod_bn <= 0) return (float)p.ln_c;
      const int nt = j >> 1, h = j & 1;
      const int ncol = min(p.ln_prod_bn, p.ln_c - nt * p.ln_prod_bn);
      const int nch = (ncol + GEMM_EC - 1) / GEMM_EC;
      return (float)(GEMM_EC * ((nch - h + 1) >> 1));
    };
    // Chan's update: fold partial (nb, mb, qb) into the running (na, ma, qa)
    auto chan = [&](float& na, float& ma, float& qa, float nb, float mb, float qb) {
      if (nb <= 0.f) return;
      const float n = na + nb, dm = mb - ma, f = __fdividef(nb, n);
      ma = fmaf(dm, f, ma);
      qa += fmaf(dm * dm, na * f, qb);
      na = n;
    };
    Coord cur{pair, pair % n_tiles, (pair / n_tiles) % pair_rows, pair / (n_tiles * pair_rows)};
    const bool ln_on = LNX == LN_CONSUME && p.ln_stats != nullptr;
    const bool shift_on = ln_on || (LNX == LN_PRODUCE && p.ln_shift != nullptr);
    if (shift_on) load_stats(cur);
    uint32_t it = 0;
    bool pre_issued = false;                                 // the loads of this warp's first chunk of the coming tile are in flight
    if (has_loads && cur.tile < total_tiles && half < nch_of(cur)) {
      if (elect_one()) issue_loads(cur.nt * BN + half * GEMM_EC, row0_of(cur), cur.b, 0);
      __syncwarp();
      pre_issued = true;
    }
    uint32_t ti = 0;
    for (; cur.tile < total_tiles; ++ti) {
      const int tile = cur.tile, b = cur.b, n0 = cur.nt * BN, mrow = row0_of(cur), nch = nch_of(cur);
      Coord nxt = cur;
      advance(nxt);
      const uint32_t acc = ti & 1;
      const float* bias = p.bias ? p.bias + (long long)b * p.bias_bs : nullptr;
      const float* colsum = ln_on ? p.ln_colsum + (long long)b * p.bias_bs : nullptr;
      float ln_a = 1.f, ln_b = 0.f, rs = 0.f, rq = 0.f, piv = 0.f;
      const float shift = shift_nx;
      if (ln_on) {                                 // this thread's row: mean / rstd from the producer's partials
        float na = 0.f, ma = 0.f, qa = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (j < p.ln_parts) chan(na, ma, qa, part_cols(j), stat_nx[j].x, stat_nx[j].y);
        if (p.ln_parts > 6 && nch > 0 && mrow + lane < p.M) {
          const float2* st = stats2 + b * stats_bs2 + mrow + lane;
          for (int q0 = 6; q0 < p.ln_parts; q0 += 6) {                     // six independent loads in flight per round trip
            float2 t2[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) t2[j] = q0 + j < p.ln_parts ? __ldg(st + (q0 + j) * p.M) : make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 6; ++j)
              if (q0 + j < p.ln_parts) chan(na, ma, qa, part_cols(q0 + j), t2[j].x, t2[j].y);
          }
        }
        load_stats(nxt);                           // the next tile's partials travel while this tile is processed
        const float var = qa * p.ln_inv_c;
        const float mres = ma - shift;             // what is left of the mean in the centred 16-bit copy
        ln_a = rsqrtf(var + p.ln_eps);
        ln_b = ln_a * mres;
        if (p.ln_health && cur.nt == 0 && half == 0 && nch > 0 && mrow + lane < p.M) {
          const float m2 = mres * mres;
          if (m2 > 1024.f * (var + p.ln_eps)) atomicAdd(p.ln_health, 1u);            // |mean - shift| > 32 sigma
          if (F16 && var + m2 > 1.0e8f) atomicAdd(p.ln_health + 1, 1u);              // rms of the copy > 1e4: fp16 range at risk
        }
      } else if (shift_on) {
        load_stats(nxt);
      }
      mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
      tc_fence_after();
      if (trc && ew == 0 && lane == 0 && ti < 2) trc[6 + ti] = clock64();
#pragma unroll 1
      for (int c = half; c < nch; c += 2, ++it) {
        const int buf = it & 1;
        const int col = n0 + c * GEMM_EC;
        uint8_t* SA = slab + buf * GEMM_BUF;
        uint8_t* SB = SA + GEMM_SLOT_A;
        // (1) accumulators: 32 columns of this thread's row (asynchronous until tmem_ld_wait)
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * GEMM_EC, r);
        // (2) bias (and, for a folded LayerNorm, the column sums) of the 32 columns: warp-uniform addresses, every load into its
        //     own registers and all of them in flight together with the TMEM load
        float bv[32];
        float4 svv[LNX == LN_CONSUME ? 8 : 1];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#ifndef VV_EXP_NOBIAS
          if (bias && col + 4 * j < p.N) t = __ldg(reinterpret_cast<const float4*>(bias + col + 4 * j));
#endif
          bv[4 * j] = t.x; bv[4 * j + 1] = t.y; bv[4 * j + 2] = t.z; bv[4 * j + 3] = t.w;
          if (LNX == LN_CONSUME) {
            svv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#ifndef VV_EXP_NOBIAS
            if (colsum && col + 4 * j < p.N) svv[j] = __ldg(reinterpret_cast<const float4*>(colsum + col + 4 * j));
#endif
          }
        }
        // (3) prefetch the residual / pre-activation of this warp's NEXT chunk into the other buffer
        if (has_loads) {
          if (c == half && !pre_issued) {                              // (rare) nobody prefetched this tile's first chunk
            if (elect_one()) issue_loads(col, mrow, b, buf);
          }
          pre_issued = false;
          if (c + 2 < nch) {
            if (elect_one()) issue_loads(col + 2 * GEMM_EC, mrow, b, buf ^ 1);
          } else if (nxt.tile < total_tiles && half < nch_of(nxt)) {
            if (elect_one()) issue_loads(nxt.nt * BN + half * GEMM_EC, row0_of(nxt), nxt.b, buf ^ 1);
            pre_issued = true;
          }
        } else {
          if (elect_one()) tma_store_wait_read1();                     // the store issued two chunks ago has read this buffer
        }
        __syncwarp();
        tmem_ld_wait();
        if (LNX == LN_CONSUME && colsum) {                                             // folded LayerNorm: c_n - rstd mean s_n
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 nb = splat2(-ln_b);
            const float2 t0 = fma2(nb, make_float2(svv[j].x, svv[j].y), make_float2(bv[4 * j], bv[4 * j + 1]));
            const float2 t1 = fma2(nb, make_float2(svv[j].z, svv[j].w), make_float2(bv[4 * j + 2], bv[4 * j + 3]));
            bv[4 * j] = t0.x; bv[4 * j + 1] = t0.y; bv[4 * j + 2] = t1.x; bv[4 * j + 3] = t1.y;
          }
        }
        if (has_loads) {
          mbar_wait(&ldbar[buf], (ld_phase >> buf) & 1);
          ld_phase ^= 1u << buf;
        }
        // (4) fused math, results staged in place
        epilogue_chunk<EPI, F16, LNX>(p, r, bv, SA, SB, lane, has_res, has_auxin, has_auxout, ln_a, rs, rq, piv, c == half, shift);
        // (5) bulk stores
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
          if (p.out_f32) tma_store_3d(&io.f32, SA, col, mrow, b);
          if (p.out_bf16) {
            if (p.split_n > 0) tma_store_3d(&io.bf16, SB, col % p.split_n, mrow, col / p.split_n);
            else tma_store_3d(&io.bf16, SB, col, mrow, b);
          }
          if (has_auxout) tma_store_3d(&io.aux, SA, col, mrow, b);
          tma_store_commit();
        }
      }
      if (LNX == LN_PRODUCE && p.stats_out && mrow + lane < p.M) {      // partial (mean, M2) of this warp's columns of the row; zeros if it had none
        float2* so = reinterpret_cast<float2*>(p.stats_out + (long long)b * p.stats_out_bs);
        const int ncols = nch > half ? GEMM_EC * ((nch - half + 1) >> 1) : 0;
        float2 st = make_float2(0.f, 0.f);
        if (ncols > 0) {
          const float inv = 1.0f / (float)ncols;
          st = make_float2(fmaf(rs, inv, piv), fmaxf(fmaf(-rs * inv, rs, rq), 0.f));
        }
        so[(long long)(2 * cur.nt + half) * p.M + mrow + lane] = st;
      }
      // this warp no longer needs accumulator stage `acc`: tell the leader's MMA thread
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
      if (trc && ew == 0 && lane == 0 && ti < 2) trc[8 + ti] = clock64();
      (void)tile;
      cur = nxt;
    }
    if (elect_one()) tma_store_wait_read0();          // the staging slabs must outlive the bulk stores that read them
    __syncwarp();
  }
  if (trc && threadIdx.x == 0) { trc[1] = clock64(); trc[13] = globaltimer_ns(); }
  tc_fence_before();
  cluster_sync_all();                                 // neither CTA may retire while its peer can still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
  if (trc && threadIdx.x == 0) trc[14] = globaltimer_ns();                                       // about to exit
}

}  // namespace vv
