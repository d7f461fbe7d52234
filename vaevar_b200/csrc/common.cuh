// Shared device helpers for the vaevar_b200 kernels (sm_100a only).
// Thin inline-PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define VV_DEVINL __device__ __forceinline__

namespace vv {

VV_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

VV_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------
VV_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
VV_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
VV_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

VV_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
VV_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
VV_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a descriptor / phase bug shows up as a trap (launch error) instead of a hung GPU.
VV_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { __trap(); }
  }
}

// ---- TMA ------------------------------------------------------------------------------------
VV_DEVINL void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 3-D tiled load global -> shared, completion counted in bytes on `bar`.
VV_DEVINL void tma_load_3d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 3-D tiled store shared -> global (bulk async-group completion); rows / columns beyond the tensor extent are clipped.
VV_DEVINL void tma_store_3d(const void* desc, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
VV_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
VV_DEVINL void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
VV_DEVINL void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// Asynchronous prefetch of `bytes` (multiple of 16, 16-byte aligned address) of global memory into L2.
VV_DEVINL void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// ---- programmatic dependent launch ------------------------------------------------------------
// Every kernel of the engine is launched with programmatic stream serialization: its prologue may overlap the tail of the
// kernel before it; pdl_wait() blocks until that kernel has completed and its writes are visible.  Both are no-ops when the
// kernel is launched without the attribute.
VV_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
VV_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

VV_DEVINL unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- clusters / CTA pairs -------------------------------------------------------------------
VV_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
VV_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
VV_DEVINL uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// 2-CTA TMA load: data lands in THIS CTA's shared memory, the byte count is credited to the mbarrier at cluster
// address `mbar_cluster_addr` (the pair leader's "full" barrier).
VV_DEVINL void tma_load_3d_2cta(void* smem_dst, const void* desc, uint32_t mbar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// Arrive on the mbarrier at the same offset in CTA `rank` of the cluster.  Default (CTA-scope release) semantics: what the
// arrival orders is TMEM traffic, which the tcgen05 fences around it cover; a cluster-scope release would cost a MEMBAR.GPU.
VV_DEVINL void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  const uint32_t addr = mapa_shared(smem_u32(bar), rank);
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
VV_DEVINL void tmem_alloc_2cta(uint32_t* smem_out, uint32_t ncols) {   // one warp in EACH CTA of the pair, collectively
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols) : "memory");
}
VV_DEVINL void tmem_relinquish_2cta() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
VV_DEVINL void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows from each CTA) * B^T (N/2 rows from each CTA); issued by ONE thread of the leader CTA.
VV_DEVINL void umma_16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once the issued MMAs retire) on the mbarrier at this offset in every CTA of `cta_mask`.
VV_DEVINL void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------
VV_DEVINL void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {   // whole warp, .sync.aligned
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols) : "memory");
}
VV_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
VV_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
VV_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
VV_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread.
VV_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on `bar` once every previously issued tcgen05.mma of this thread has completed.
VV_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane+i), columns [col, col+32).
VV_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
VV_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The store twin of tmem_ld32: thread i of the warp writes lane (base_lane+i), columns [col, col+32).
VV_DEVINL void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
VV_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (rows of 64 bf16 = 128 B,
// 8-row swizzle atoms 1024 B apart).  Bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor.
VV_DEVINL uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                            // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                    // SBO = 1024 B   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                            // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: A,B both fp16 (format 0) or both bf16 (format 1), K-major, D fp32, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_16(int M, int N, bool f16) {
  return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// ---- math -----------------------------------------------------------------------------------
// Exact-erf GELU (swinblock.py:14 nn.GELU) with erfc from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the
// 16-bit rounding of the result).  With e = exp(-x^2/2), t = 1/(1 + p|x|/sqrt2) and q = t*poly(t)*e = erfc(|x|/sqrt2):
//   Phi(x) = 1 - q/2 (x >= 0),  q/2 (x < 0)   =>   gelu(x) = x Phi(x) = max(x, 0) - |x| q / 2
//   gelu'(x) = Phi(x) + x e / sqrt(2 pi)
// One MUFU.EX2, one MUFU.RCP and ~12 FMA-pipe instructions per element, branch-free.
VV_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
VV_DEVINL float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erfc(|x|/sqrt2); *e_out = exp(-x^2/2)
VV_DEVINL float gelu_erfc_abs(float x, float* e_out) {
  const float e = ex2_approx(x * (x * -0.72134752044448170f));          // -0.5 * log2(e)
  const float t = rcp_approx(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752f, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  *e_out = e;
  return poly * (t * e);
}
VV_DEVINL float gelu_erf(float x) {
  float e;
  const float q = gelu_erfc_abs(x, &e);
  return fmaf(fabsf(x) * -0.5f, q, fmaxf(x, 0.0f));
}
VV_DEVINL float gelu_erf_grad(float x) {
  float e;
  const float q = gelu_erfc_abs(x, &e);
  const float cdf = 0.5f + copysignf(fmaf(-0.5f, q, 0.5f), x);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}
// gelu(x) and gelu'(x) together (they share the exponential and the erfc polynomial): the forward GEMM saves gelu'(u) instead
// of the pre-activation u, so the gradient GEMM's epilogue is a plain multiply.
VV_DEVINL void gelu_erf_both(float x, float* y, float* dy) {
  float e;
  const float q = gelu_erfc_abs(x, &e);
  const float sh = copysignf(0.5f, x);                   // Phi(x) = 1/2 + sign(x) (1/2 - q/2)
  const float cdf = fmaf(-sh, q, 0.5f + sh);
  *y = x * cdf;
  *dy = fmaf(x * 0.39894228040143268f, e, cdf);
}
// ---- packed fp32 (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per instruction, bit-identical to the scalar
// ones).  The CUDA-core epilogues of the GEMMs are bound by the FP32 pipe's issue rate; pairing halves their FMA-pipe
// instruction count at unchanged precision.
VV_DEVINL float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
VV_DEVINL float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
VV_DEVINL float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
VV_DEVINL float2 splat2(float v) { return make_float2(v, v); }
// gelu / gelu' of two elements: the arithmetic of gelu_erf_both, paired (MUFU and the sign / abs bit operations stay scalar)
VV_DEVINL void gelu_erf_both2(float2 x, float2* y, float2* dy) {
  const float2 xs = mul2(x, splat2(-0.72134752044448170f));
  const float2 ea = mul2(x, xs);                                      // -x^2 log2(e) / 2
  const float2 e = make_float2(ex2_approx(ea.x), ex2_approx(ea.y));
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = fma2(ax, splat2(0.3275911f * 0.70710678118654752f), splat2(1.0f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 poly = fma2(splat2(1.061405429f), t, splat2(-1.453152027f));
  poly = fma2(poly, t, splat2(1.421413741f));
  poly = fma2(poly, t, splat2(-0.284496736f));
  poly = fma2(poly, t, splat2(0.254829592f));
  const float2 q = mul2(poly, mul2(t, e));                            // erfc(|x| / sqrt2)
  const float2 sh = make_float2(copysignf(0.5f, x.x), copysignf(0.5f, x.y));
  const float2 nsh = make_float2(-sh.x, -sh.y);
  const float2 cdf = fma2(nsh, q, add2(sh, splat2(0.5f)));            // Phi(x) = 1/2 + sign(x) (1/2 - q/2)
  *y = mul2(x, cdf);
  *dy = fma2(mul2(x, splat2(0.39894228040143268f)), e, cdf);
}
VV_DEVINL float2 gelu_erf2(float2 x) {
  float2 y, dy;
  gelu_erf_both2(x, &y, &dy);
  return y;
}
VV_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
VV_DEVINL uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
VV_DEVINL float2 unpack_bf16(uint32_t w) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&w);
  return __bfloat1622float2(t);
}

}  // namespace vv
