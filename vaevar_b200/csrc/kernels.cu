// Memory-bound kernels of one U-shaped Swin network application and its hand-derived adjoint:
// LayerNorm fwd/bwd (with the PatchMerging gather / PatchExpand pixel-shuffle folded into the addressing),
// 4x4-window attention fwd/bwd (roll + window partition/reverse folded into the addressing, P recomputed in
// the backward), the 2x2/stride-2 patch operators, and the per-channel affine seams.
#include <stdlib.h>

#include "ops.h"

namespace vv {

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VV_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// =============================================================================================
// LayerNorm
// =============================================================================================
template <int MAP>
VV_DEVINL long long ln_elem_off(int r, int c, int C, int gw, long long ld) {
  if (MAP == MAP_PLAIN) {
    return (long long)r * ld + c;
  } else if (MAP == MAP_MERGE) {
    const int D = C >> 2, chunk = c / D, cc = c - chunk * D;
    const int hw = gw >> 1, i = r / hw, j = r - i * hw;
    const int tok = (2 * i + (chunk & 1)) * gw + 2 * j + (chunk >> 1);
    return (long long)tok * ld + cc;
  } else {
    const int I = r / gw, J = r - I * gw;
    const int tok = (I >> 1) * (gw >> 1) + (J >> 1);
    const int chunk = (I & 1) * 2 + (J & 1);
    return (long long)tok * ld + chunk * C + c;
  }
}

// A row of C floats is owned by LANES lanes (8, 16 or 32: 4, 2 or 1 rows per warp); each lane holds NV float4 vectors,
// vector k of lane l = columns 4*(l + LANES*k) .. +3, so every access of the sub-warp is one contiguous 16*LANES-byte segment.
// Narrow rows (C = 96 / 192 of the towers) thus still move 16 bytes per lane per access instead of 4.
VV_DEVINL uint32_t pack16_rt(float a, float b, bool f16) { return f16 ? pack16<true>(a, b) : pack16<false>(a, b); }
VV_DEVINL void st4_16(bf16* dst, const float4& v, bool f16) {
  uint2 w;
  w.x = pack16_rt(v.x, v.y, f16);
  w.y = pack16_rt(v.z, v.w, f16);
  *reinterpret_cast<uint2*>(dst) = w;
}
template <int LANES>
VV_DEVINL float row_sum(float v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LANES, int NV, int MAP>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const LnArgs a) {
  constexpr int RPW = 32 / LANES;
  const int lane = threadIdx.x & 31, sl = lane % LANES;
  const int r = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LANES;
  const int b = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  const bool live = r < a.rows;                    // dead rows still take part in the shuffles
  const int rr = live ? r : a.rows - 1;
  const float* x = a.x + (long long)b * a.x_bs;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k] = *reinterpret_cast<const float4*>(x + ln_elem_off<MAP>(rr, 4 * (sl + LANES * k), a.C, a.gw, a.ld_x));
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  const float mean = row_sum<LANES>(s) / a.C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
    q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  const float qc = row_sum<LANES>(q);
  if (!live) return;
  if (a.stats_out) {                               // statistics-only mode: (mean, M2) of the row, the row's shift (= its mean) and
    if (sl == 0) {                                 // the CENTRED 16-bit copy x - mean a folded LayerNorm consumes (GemmArgs::ln_stats)
      reinterpret_cast<float2*>(a.stats_out)[(long long)b * a.rows + r] = make_float2(mean, qc);
      if (a.shift_out) a.shift_out[(long long)b * a.rows + r] = mean;
    }
    const float sh = a.shift_out ? mean : 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float4 c4 = make_float4(v[k].x - sh, v[k].y - sh, v[k].z - sh, v[k].w - sh);
      st4_16(a.out_bf16 + (long long)b * a.ob_bs + (long long)r * a.ld_ob + 4 * (sl + LANES * k), c4, a.out_f16 != 0);
    }
    return;
  }
  const float rstd = rsqrtf(qc / a.C + a.eps);
  const float* g = a.gamma + (long long)b * a.gb_bs;
  const float* be = a.beta + (long long)b * a.gb_bs;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (sl + LANES * k);
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c)), bb = __ldg(reinterpret_cast<const float4*>(be + c));
    float4 y;
    y.x = (v[k].x - mean) * rstd * gg.x + bb.x; y.y = (v[k].y - mean) * rstd * gg.y + bb.y;
    y.z = (v[k].z - mean) * rstd * gg.z + bb.z; y.w = (v[k].w - mean) * rstd * gg.w + bb.w;
    if (a.out_bf16) st4_16(a.out_bf16 + (long long)b * a.ob_bs + (long long)r * a.ld_ob + c, y, a.out_f16 != 0);
    if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + (long long)b * a.of_bs + (long long)r * a.ld_of + c) = y;
  }
}

template <int LANES, int NV, int MAP>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdArgs a) {
  constexpr int RPW = 32 / LANES;
  const int lane = threadIdx.x & 31, sl = lane % LANES;
  const int r = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LANES;
  const int b = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  const bool live = r < a.rows;
  const int rr = live ? r : a.rows - 1;
  const float* x = a.x + (long long)b * a.x_bs;
  const float* dy = a.dy + (long long)b * a.dy_bs + (long long)rr * a.ld_dy;
  const bf16* dy16 = a.dy16 ? a.dy16 + (long long)b * a.dy_bs + (long long)rr * a.ld_dy : nullptr;
  const float* g = a.gamma + (long long)b * a.gb_bs;
  float4 v[NV], gd[NV], rs[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (sl + LANES * k);
    v[k] = *reinterpret_cast<const float4*>(x + ln_elem_off<MAP>(rr, c, a.C, a.gw, a.ld_x));
    if (dy16) {
      const uint2 w = *reinterpret_cast<const uint2*>(dy16 + c);
      const float2 lo = unpack_bf16(w.x), hi = unpack_bf16(w.y);
      gd[k] = make_float4(lo.x, lo.y, hi.x, hi.y);
    } else {
      gd[k] = *reinterpret_cast<const float4*>(dy + c);
    }
    if (a.dres) rs[k] = *reinterpret_cast<const float4*>(a.dres + (long long)b * a.dres_bs + ln_elem_off<MAP>(rr, c, a.C, a.gw, a.ld_dres));
    else rs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  const float mean = row_sum<LANES>(s) / a.C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
    q += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
  }
  const float rstd = rsqrtf(row_sum<LANES>(q) / a.C + a.eps);
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + 4 * (sl + LANES * k)));
    v[k].x *= rstd; v[k].y *= rstd; v[k].z *= rstd; v[k].w *= rstd;                  // xhat
    gd[k].x *= gg.x; gd[k].y *= gg.y; gd[k].z *= gg.z; gd[k].w *= gg.w;
    m1 += (gd[k].x + gd[k].y) + (gd[k].z + gd[k].w);
    m2 += gd[k].x * v[k].x + gd[k].y * v[k].y + gd[k].z * v[k].z + gd[k].w * v[k].w;
  }
  m1 = row_sum<LANES>(m1) / a.C;
  m2 = row_sum<LANES>(m2) / a.C;
  if (!live) return;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (sl + LANES * k);
    float4 d;
    d.x = (gd[k].x - m1 - v[k].x * m2) * rstd + rs[k].x; d.y = (gd[k].y - m1 - v[k].y * m2) * rstd + rs[k].y;
    d.z = (gd[k].z - m1 - v[k].z * m2) * rstd + rs[k].z; d.w = (gd[k].w - m1 - v[k].w * m2) * rstd + rs[k].w;
    *reinterpret_cast<float4*>(a.dx + (long long)b * a.dx_bs + ln_elem_off<MAP>(r, c, a.C, a.gw, a.ld_dx)) = d;
    if (a.dx_bf16) st4_16(a.dx_bf16 + (long long)b * a.dxb_bs + ln_elem_off<MAP>(r, c, a.C, a.gw, a.ld_dxb), d, false);
  }
}

// key = map * 1000 + C / 4 (float4 vectors per row)
#define VV_LN_CASE(KERNEL, ARGS, LANES, NV, MAP)                                                          \
  launch_kernel(KERNEL<LANES, NV, MAP>, dim3((ARGS.rows + 8 * (32 / LANES) - 1) / (8 * (32 / LANES)), ARGS.batch), dim3(256), 0, s, ARGS)
#define VV_LN_DISPATCH(KERNEL, ARGS)                                                     \
  {                                                                                      \
    switch (ARGS.map * 1000 + ARGS.C / 4) {                                              \
      case 16: VV_LN_CASE(KERNEL, ARGS, 8, 2, MAP_PLAIN); break;      /* C = 64   */     \
      case 24: VV_LN_CASE(KERNEL, ARGS, 8, 3, MAP_PLAIN); break;      /* C = 96   */     \
      case 32: VV_LN_CASE(KERNEL, ARGS, 16, 2, MAP_PLAIN); break;     /* C = 128  */     \
      case 48: VV_LN_CASE(KERNEL, ARGS, 16, 3, MAP_PLAIN); break;     /* C = 192  */     \
      case 96: VV_LN_CASE(KERNEL, ARGS, 32, 3, MAP_PLAIN); break;     /* C = 384  */     \
      case 288: VV_LN_CASE(KERNEL, ARGS, 32, 9, MAP_PLAIN); break;    /* C = 1152 */     \
      case 1064: VV_LN_CASE(KERNEL, ARGS, 32, 2, MAP_MERGE); break;   /* C = 256  */     \
      case 1096: VV_LN_CASE(KERNEL, ARGS, 32, 3, MAP_MERGE); break;   /* C = 384  */     \
      case 1192: VV_LN_CASE(KERNEL, ARGS, 32, 6, MAP_MERGE); break;   /* C = 768  */     \
      case 2048: VV_LN_CASE(KERNEL, ARGS, 16, 3, MAP_EXPAND); break;  /* C = 192  */     \
      case 2016: VV_LN_CASE(KERNEL, ARGS, 8, 2, MAP_EXPAND); break;   /* C = 64   */     \
      case 2024: VV_LN_CASE(KERNEL, ARGS, 8, 3, MAP_EXPAND); break;   /* C = 96   */     \
      default: break;                                                                    \
    }                                                                                    \
  }

bool ln_supported(int map, int C) {
  if (C % 32) return false;
  switch (map * 1000 + C / 4) {
    case 16: case 24: case 32: case 48: case 96: case 288: case 1064: case 1096: case 1192: case 2016: case 2024: case 2048: return true;
    default: return false;
  }
}

void launch_ln_fwd(const LnArgs& a, cudaStream_t s) { VV_LN_DISPATCH(ln_fwd_kernel, a) }
void launch_ln_bwd(const LnBwdArgs& a, cudaStream_t s) { VV_LN_DISPATCH(ln_bwd_kernel, a) }

// =============================================================================================
// Window attention: one warp per (window, head); N = 16 tokens per window, so Q K^T, P V and the four products of the
// backward are single m16n8k16 tensor-core tiles (mma.sync, bf16 in / fp32 accumulate -- the 16-token problem is far
// below a tcgen05 tile).  Operands are staged once with 16-byte cp.async into padded shared-memory rows (conflict-free
// for both the 32-bit fragment loads and ldmatrix); softmax, bias, the latitude mask and dS stay in fp32 registers;
// results are staged back through the operand buffers and written with 16-byte coalesced stores.
// =============================================================================================
VV_DEVINL void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
VV_DEVINL void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 16-bit operand format of the forward pass: IEEE fp16 (F16) or bf16
template <bool F16>
VV_DEVINL void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (F16) mma_f16_16816(d, a, b0, b1);
  else mma_bf16_16816(d, a, b0, b1);
}
VV_DEVINL uint32_t half2_to_bf162(uint32_t w) {
  const float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
  return pack_bf16(f.x, f.y);
}
VV_DEVINL void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row_ptr)));
}

// SPLIT warps share one (window, head): warp `sl` owns the head-dimension slice [sl*HD/SPLIT, (sl+1)*HD/SPLIT) -- its third of
// the loads, of the K-reduction of S = Q K^T (and dP = dO V^T), and of the output columns; the 16x16 partial products are
// summed across the SPLIT warps through shared memory (one block barrier).  SPLIT = 3 for the d=1152 trunk (head_dim 192):
// 768 (window, head) items are too few warps to keep 148 SMs busy, and a warp's 18-24 KB of loads and ~50 MMAs is a long
// serial chain; SPLIT = 1 (four independent items per CTA) for the head_dim-32 towers.
template <int HD, bool BWD, int SPLIT>
struct AttnSmem {
  static constexpr int WARPS = SPLIT == 1 ? 4 : SPLIT;
  static constexpr int ITEMS = WARPS / SPLIT;    // (window, head) items per CTA
  static constexpr int W2 = HD / 2;             // 32-bit words (16-bit pairs) per token row
  static constexpr int RS = W2 + 4;             // padded row stride in words: 16-byte aligned rows, 4g+t / ldmatrix conflict-free
  static constexpr int MAT = 16 * RS;           // one 16 x HD operand
  static constexpr int PW = 12;                 // words per row of the 16 x 16 bf16 P / dS tiles (24 bf16 = 48 B)
  static constexpr int ITEM_WORDS = (BWD ? 4 : 3) * MAT;                      // Q, K, V (, dO) of one item
  static constexpr int WARP_WORDS = BWD ? 2 * 16 * PW : 0;                    // per-warp P / dS scratch
  static constexpr int RED_WORDS = SPLIT == 1 ? 0 : SPLIT * 32 * (BWD ? 16 : 8);   // partial S (and dP) fragments
  static constexpr int WORDS = ITEMS * ITEM_WORDS + WARPS * WARP_WORDS + RED_WORDS;
  static constexpr int BYTES = WORDS * 4;
};

// F16: qkv (and the forward output) are fp16.  The backward pass recomputes P from the fp16 Q, K exactly as the forward did,
// then converts Q, K, V to bf16 in shared memory: its products pair them with bf16 gradients whose range fp16 cannot hold.
template <int HD, bool BWD, bool F16, int SPLIT>
__global__ void __launch_bounds__(AttnSmem<HD, BWD, SPLIT>::WARPS * 32) attn_kernel(const AttnArgs a) {
  using L = AttnSmem<HD, BWD, SPLIT>;
  constexpr int RS = L::RS, PW = L::PW;
  constexpr int HDS = HD / SPLIT;                // head-dimension columns this warp owns
  constexpr int CHS = HDS / 8;                   // 16-byte chunks per row of the slice
  static_assert(HDS % 16 == 0, "slice must be a multiple of the MMA k step");
  extern __shared__ __align__(16) uint32_t attn_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nww = a.gw >> 2, nwh = a.gh >> 2;
  const int li = warp / SPLIT, sl = warp - li * SPLIT;           // item within the CTA, slice within the item
  const int item = blockIdx.x * L::ITEMS + li;
  pdl_launch_dependents();
  pdl_wait();
  if (item >= nww * nwh * a.heads) return;      // uniform over the SPLIT warps of an item (SPLIT > 1: over the CTA)
  const int b = blockIdx.y;
  const int win = item / a.heads, h = item - win * a.heads;
  const int wi = win / nww, wj = win - wi * nww;
  const int d = a.heads * HD;
  const int c_lo = sl * HDS;                     // first column of the slice

  uint32_t* Qs = attn_sm + li * L::ITEM_WORDS;
  uint32_t* Ks = Qs + L::MAT;
  uint32_t* Vs = Ks + L::MAT;
  uint32_t* dOs = Vs + L::MAT;                                   // BWD only
  uint32_t* Ps = attn_sm + L::ITEMS * L::ITEM_WORDS + warp * L::WARP_WORDS;   // BWD only: this warp's bf16 P
  uint32_t* dSs = Ps + 16 * PW;                                  // BWD only: this warp's bf16 dS
  float* red = reinterpret_cast<float*>(attn_sm + L::ITEMS * L::ITEM_WORDS + L::WARPS * L::WARP_WORDS);   // SPLIT > 1

  // original-grid token index of window-local token tk (roll by -shift folded in; swinblock.py:275, 297)
  auto tok_of = [&](int tk) {
    int row = 4 * wi + (tk >> 2) + a.shift; if (row >= a.gh) row -= a.gh;
    int col = 4 * wj + (tk & 3) + a.shift;  if (col >= a.gw) col -= a.gw;
    return row * a.gw + col;
  };

  const bf16* qkv = a.qkv + (long long)b * a.qkv_bs;
  {
    constexpr int NMAT = BWD ? 4 : 3;
    for (int idx = lane; idx < 16 * 4 * CHS; idx += 32) {
      const int ch = idx % CHS, rowm = idx / CHS, m = rowm & 3, tk = rowm >> 2;   // m: 0 Q, 1 K, 2 V, 3 dO
      if (m >= NMAT) continue;
      const bf16* src = m < 3 ? qkv + (long long)tok_of(tk) * a.ld_qkv + m * d + h * HD
                              : a.dout + (long long)b * a.o_bs + (long long)tok_of(tk) * a.ld_o + h * HD;
      uint32_t* dst = Qs + m * L::MAT + tk * RS + (c_lo >> 1) + 4 * ch;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src + c_lo + 8 * ch) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncwarp();

  // ---- S = Q K^T : two m16n8 tiles (key tokens 0-7 and 8-15), k over this warp's slice of the head dimension ----
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ks = 0; ks < HDS / 16; ++ks) {
    const int w = (c_lo >> 1) + ks * 8 + t;
    const uint32_t af[4] = {Qs[g * RS + w], Qs[(g + 8) * RS + w], Qs[g * RS + w + 4], Qs[(g + 8) * RS + w + 4]};
    mma_16816<F16>(s0, af, Ks[g * RS + w], Ks[g * RS + w + 4]);
    mma_16816<F16>(s1, af, Ks[(g + 8) * RS + w], Ks[(g + 8) * RS + w + 4]);
  }
  if (BWD && F16) {                                              // own slice of Q, K, V -> bf16 in place (the gradient products are bf16)
    __syncwarp();
    for (int idx = lane; idx < 3 * 16 * CHS; idx += 32) {
      const int ch = idx % CHS, rowm = idx / CHS;                // rowm = matrix * 16 + token
      uint4* ptr = reinterpret_cast<uint4*>(Qs + (rowm >> 4) * L::MAT + (rowm & 15) * RS + (c_lo >> 1) + 4 * ch);
      uint4 w = *ptr;
      w.x = half2_to_bf162(w.x); w.y = half2_to_bf162(w.y); w.z = half2_to_bf162(w.z); w.w = half2_to_bf162(w.w);
      *ptr = w;
    }
    __syncwarp();
  }
  // BWD: dP = dO V^T over the same slice (its K-reduction is shared with S's through one exchange)
  float dp0[4] = {0.f, 0.f, 0.f, 0.f}, dp1[4] = {0.f, 0.f, 0.f, 0.f};
  if (BWD) {
#pragma unroll
    for (int ks = 0; ks < HDS / 16; ++ks) {
      const int w = (c_lo >> 1) + ks * 8 + t;
      const uint32_t af[4] = {dOs[g * RS + w], dOs[(g + 8) * RS + w], dOs[g * RS + w + 4], dOs[(g + 8) * RS + w + 4]};
      mma_bf16_16816(dp0, af, Vs[g * RS + w], Vs[g * RS + w + 4]);
      mma_bf16_16816(dp1, af, Vs[(g + 8) * RS + w], Vs[(g + 8) * RS + w + 4]);
    }
  }
  if (SPLIT > 1) {                                               // sum the partial fragments over the SPLIT warps
    constexpr int FR = BWD ? 16 : 8;
    float* mine = red + (sl * 32 + lane) * FR;
    *reinterpret_cast<float4*>(mine) = make_float4(s0[0], s0[1], s0[2], s0[3]);
    *reinterpret_cast<float4*>(mine + 4) = make_float4(s1[0], s1[1], s1[2], s1[3]);
    if (BWD) {
      *reinterpret_cast<float4*>(mine + 8) = make_float4(dp0[0], dp0[1], dp0[2], dp0[3]);
      *reinterpret_cast<float4*>(mine + 12) = make_float4(dp1[0], dp1[1], dp1[2], dp1[3]);
    }
    __syncthreads();
#pragma unroll
    for (int o = 1; o < SPLIT; ++o) {
      const float* oth = red + (((sl + o) % SPLIT) * 32 + lane) * FR;
      const float4 x0 = *reinterpret_cast<const float4*>(oth), x1 = *reinterpret_cast<const float4*>(oth + 4);
      s0[0] += x0.x; s0[1] += x0.y; s0[2] += x0.z; s0[3] += x0.w;
      s1[0] += x1.x; s1[1] += x1.y; s1[2] += x1.z; s1[3] += x1.w;
      if (BWD) {
        const float4 y0 = *reinterpret_cast<const float4*>(oth + 8), y1 = *reinterpret_cast<const float4*>(oth + 12);
        dp0[0] += y0.x; dp0[1] += y0.y; dp0[2] += y0.z; dp0[3] += y0.w;
        dp1[0] += y1.x; dp1[1] += y1.y; dp1[2] += y1.z; dp1[3] += y1.w;
      }
    }
  }
  // ---- P = softmax(scale S + bias + mask); this thread owns rows g and g+8, columns {2t,2t+1} and {8+2t,9+2t} ----
  const float scale = rsqrtf((float)HD);
  const bool masked_win = a.shift > 0 && wi == nwh - 1;          // swinblock.py:236-260: latitude bands only
  const float* rb = a.relbias + (long long)b * a.relbias_bs + h * 256;
  float p0[4], p1[4];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {                               // hh = 0: row g, hh = 1: row g + 8
    const int i = g + 8 * hh;
    const float2 b0 = *reinterpret_cast<const float2*>(rb + i * 16 + 2 * t);
    const float2 b1 = *reinterpret_cast<const float2*>(rb + i * 16 + 8 + 2 * t);
    float v0 = s0[2 * hh] * scale + b0.x, v1 = s0[2 * hh + 1] * scale + b0.y;
    float v2 = s1[2 * hh] * scale + b1.x, v3 = s1[2 * hh + 1] * scale + b1.y;
    if (masked_win) {
      // window rows {0,1} and {2,3} lie in different latitude bands: key columns 0..7 (this thread's v0,v1) are rows 0,1,
      // key columns 8..15 (v2,v3) are rows 2,3; query row i is in the upper band iff (i >> 2) < 2.
      if ((i >> 2) < 2) { v2 += -100.0f; v3 += -100.0f; }
      else { v0 += -100.0f; v1 += -100.0f; }
    }
    float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    v0 = expf(v0 - mx); v1 = expf(v1 - mx); v2 = expf(v2 - mx); v3 = expf(v3 - mx);
    float sum = v0 + v1 + v2 + v3;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    p0[2 * hh] = v0 * inv; p0[2 * hh + 1] = v1 * inv;
    p1[2 * hh] = v2 * inv; p1[2 * hh + 1] = v3 * inv;
  }

  // lane -> row address of the ldmatrix.x4.trans that yields the B fragments (k = token, n = channel) of two
  // adjacent 8-channel tiles of a token-major 16 x HD operand
  const int lm = lane >> 3, lr = lane & 7;
  auto bfrag_ptr = [&](const uint32_t* M, int c0) { return M + ((lm & 1) * 8 + lr) * RS + (c0 >> 1) + (lm >> 1) * 4; };
  // this warp's slice of a staged 16 x HD result (16-bit pairs) -> global rows tok_of(r), 16-byte coalesced
  auto copy_out = [&](const uint32_t* M, bf16* gbase, long long ld, int coloff) {
    for (int idx = lane; idx < 16 * CHS; idx += 32) {
      const int r = idx / CHS, ch = idx - r * CHS;
      *reinterpret_cast<uint4*>(gbase + (long long)tok_of(r) * ld + coloff + c_lo + 8 * ch) =
          *reinterpret_cast<const uint4*>(M + r * RS + (c_lo >> 1) + 4 * ch);
    }
  };

  if (!BWD) {
    // ---- O = P V (this warp's columns) ----
    const uint32_t pa[4] = {pack16<F16>(p0[0], p0[1]), pack16<F16>(p0[2], p0[3]), pack16<F16>(p1[0], p1[1]), pack16<F16>(p1[2], p1[3])};
    __syncwarp();                                                 // every lane is done reading its Q slice before it is reused for O
#pragma unroll
    for (int c0 = c_lo; c0 < c_lo + HDS; c0 += 16) {
      uint32_t vb[4];
      ldmatrix_x4_trans(vb, bfrag_ptr(Vs, c0));
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_16816<F16>(o0, pa, vb[0], vb[1]);
      mma_16816<F16>(o1, pa, vb[2], vb[3]);
      Qs[g * RS + (c0 >> 1) + t] = pack16<F16>(o0[0], o0[1]);
      Qs[(g + 8) * RS + (c0 >> 1) + t] = pack16<F16>(o0[2], o0[3]);
      Qs[g * RS + (c0 >> 1) + 4 + t] = pack16<F16>(o1[0], o1[1]);
      Qs[(g + 8) * RS + (c0 >> 1) + 4 + t] = pack16<F16>(o1[2], o1[3]);
    }
    __syncwarp();
    copy_out(Qs, a.out + (long long)b * a.o_bs, a.ld_o, h * HD);
  } else {
    // ---- dS = P o (dP - rowsum(dP o P)) ----
    float ds0[4], ds1[4];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float rs = dp0[2 * hh] * p0[2 * hh] + dp0[2 * hh + 1] * p0[2 * hh + 1] + dp1[2 * hh] * p1[2 * hh] + dp1[2 * hh + 1] * p1[2 * hh + 1];
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      ds0[2 * hh] = p0[2 * hh] * (dp0[2 * hh] - rs); ds0[2 * hh + 1] = p0[2 * hh + 1] * (dp0[2 * hh + 1] - rs);
      ds1[2 * hh] = p1[2 * hh] * (dp1[2 * hh] - rs); ds1[2 * hh + 1] = p1[2 * hh + 1] * (dp1[2 * hh + 1] - rs);
    }
    // bf16 copies of P and dS for the transposed (ldmatrix.trans) A fragments
    Ps[g * PW + t] = pack_bf16(p0[0], p0[1]);        Ps[(g + 8) * PW + t] = pack_bf16(p0[2], p0[3]);
    Ps[g * PW + 4 + t] = pack_bf16(p1[0], p1[1]);    Ps[(g + 8) * PW + 4 + t] = pack_bf16(p1[2], p1[3]);
    dSs[g * PW + t] = pack_bf16(ds0[0], ds0[1]);     dSs[(g + 8) * PW + t] = pack_bf16(ds0[2], ds0[3]);
    dSs[g * PW + 4 + t] = pack_bf16(ds1[0], ds1[1]); dSs[(g + 8) * PW + 4 + t] = pack_bf16(ds1[2], ds1[3]);
    __syncwarp();
    // A fragments of P^T and dS^T: matrix mi of the x4 load = block (rows (mi>>1)*8.., cols (mi&1)*8..) of the stored tile
    uint32_t pT[4], dsT[4];
    ldmatrix_x4_trans(pT, Ps + ((lm >> 1) * 8 + lr) * PW + (lm & 1) * 4);
    ldmatrix_x4_trans(dsT, dSs + ((lm >> 1) * 8 + lr) * PW + (lm & 1) * 4);
    const uint32_t dsA[4] = {pack_bf16(ds0[0], ds0[1]), pack_bf16(ds0[2], ds0[3]), pack_bf16(ds1[0], ds1[1]), pack_bf16(ds1[2], ds1[3])};

    // ---- dV = P^T dO  -> staged in the V buffer (this warp's V columns are dead after dP) ----
#pragma unroll
    for (int c0 = c_lo; c0 < c_lo + HDS; c0 += 16) {
      uint32_t bb[4];
      ldmatrix_x4_trans(bb, bfrag_ptr(dOs, c0));
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(o0, pT, bb[0], bb[1]);
      mma_bf16_16816(o1, pT, bb[2], bb[3]);
      Vs[g * RS + (c0 >> 1) + t] = pack_bf16(o0[0], o0[1]);
      Vs[(g + 8) * RS + (c0 >> 1) + t] = pack_bf16(o0[2], o0[3]);
      Vs[g * RS + (c0 >> 1) + 4 + t] = pack_bf16(o1[0], o1[1]);
      Vs[(g + 8) * RS + (c0 >> 1) + 4 + t] = pack_bf16(o1[2], o1[3]);
    }
    __syncwarp();                                                 // dO slice fully consumed -> it takes dQ
    // ---- dQ = scale dS K -> staged in the dO buffer ----
#pragma unroll
    for (int c0 = c_lo; c0 < c_lo + HDS; c0 += 16) {
      uint32_t bb[4];
      ldmatrix_x4_trans(bb, bfrag_ptr(Ks, c0));
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(o0, dsA, bb[0], bb[1]);
      mma_bf16_16816(o1, dsA, bb[2], bb[3]);
      dOs[g * RS + (c0 >> 1) + t] = pack_bf16(o0[0] * scale, o0[1] * scale);
      dOs[(g + 8) * RS + (c0 >> 1) + t] = pack_bf16(o0[2] * scale, o0[3] * scale);
      dOs[g * RS + (c0 >> 1) + 4 + t] = pack_bf16(o1[0] * scale, o1[1] * scale);
      dOs[(g + 8) * RS + (c0 >> 1) + 4 + t] = pack_bf16(o1[2] * scale, o1[3] * scale);
    }
    __syncwarp();                                                 // K slice fully consumed -> it takes dK
    // ---- dK = scale dS^T Q -> staged in the K buffer ----
#pragma unroll
    for (int c0 = c_lo; c0 < c_lo + HDS; c0 += 16) {
      uint32_t bb[4];
      ldmatrix_x4_trans(bb, bfrag_ptr(Qs, c0));
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(o0, dsT, bb[0], bb[1]);
      mma_bf16_16816(o1, dsT, bb[2], bb[3]);
      Ks[g * RS + (c0 >> 1) + t] = pack_bf16(o0[0] * scale, o0[1] * scale);
      Ks[(g + 8) * RS + (c0 >> 1) + t] = pack_bf16(o0[2] * scale, o0[3] * scale);
      Ks[g * RS + (c0 >> 1) + 4 + t] = pack_bf16(o1[0] * scale, o1[1] * scale);
      Ks[(g + 8) * RS + (c0 >> 1) + 4 + t] = pack_bf16(o1[2] * scale, o1[3] * scale);
    }
    __syncwarp();
    bf16* dqkv = a.dqkv + (long long)b * a.qkv_bs;
    copy_out(dOs, dqkv, a.ld_qkv, h * HD);              // dQ
    copy_out(Ks, dqkv, a.ld_qkv, d + h * HD);           // dK
    copy_out(Vs, dqkv, a.ld_qkv, 2 * d + h * HD);       // dV
  }
}

template <int HD, bool BWD, bool F16, int SPLIT>
static void launch_attn_t(const AttnArgs& a, cudaStream_t s) {
  using L = AttnSmem<HD, BWD, SPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_kernel<HD, BWD, F16, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
    attr_set = true;
  }
  const int items = (a.gh / 4) * (a.gw / 4) * a.heads;
  dim3 grid((items + L::ITEMS - 1) / L::ITEMS, a.batch);
  launch_kernel(attn_kernel<HD, BWD, F16, SPLIT>, grid, dim3(L::WARPS * 32), L::BYTES, s, a);
}
template <bool BWD>
static void launch_attn_d(const AttnArgs& a, cudaStream_t s) {
  if (a.hd == 32) { if (a.f16) launch_attn_t<32, BWD, true, 1>(a, s); else launch_attn_t<32, BWD, false, 1>(a, s); }
  else if (a.hd == 192) { if (a.f16) launch_attn_t<192, BWD, true, 3>(a, s); else launch_attn_t<192, BWD, false, 3>(a, s); }
}
void launch_attn_fwd(const AttnArgs& a, cudaStream_t s) { launch_attn_d<false>(a, s); }
void launch_attn_bwd(const AttnArgs& a, cudaStream_t s) { launch_attn_d<true>(a, s); }

// =============================================================================================
// Patch operators
// =============================================================================================
// The 2x2 / stride-2 patch operators are small GEMMs (K = 4 C_g <= 56 for P2T, K = D = 96 for T2P).  On the CUDA cores they were
// bound by instruction issue (245 M FMA per launch at ~0.55 IPC per scheduler, ncu); they run on the tensor cores as
// mma.sync.m16n8k8 TF32 tiles instead: fp32 operands rounded to TF32 (10-bit mantissa, like the fp16 operands of every other
// Linear of the network), fp32 accumulation.
VV_DEVINL uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
VV_DEVINL void mma_tf32_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int P2T_TOK = 64;         // tokens per CTA (4 warps x one m16 tile); the tile may span token rows

// tok_out[g][tok][c] = bias[g][c] + sum_k patch[tok][k] Wp[k][c] (+ APE): per warp 16 tokens x D channels, K padded to 8.
// NT = D / 8 n-tiles.
template <int NT>
__global__ void __launch_bounds__(128) p2t_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) float p2t_sm[];
  constexpr int D = NT * 8;
  constexpr int PS = P2T_TOK + 8;                      // patch row stride: (8 t + g) % 32 distinct -> conflict-free A fragments
  constexpr int WS = D + 8;                            // weight row stride: (8 t + g) % 32 distinct -> conflict-free B fragments
  const int g = blockIdx.y;
  const int W0 = a.W >> 1;
  const int L0 = (a.H >> 1) * W0;
  const int t0 = blockIdx.x * P2T_TOK;
  pdl_launch_dependents();
  pdl_wait();
  const int cnt = a.kcnt[g];
  const int K = cnt * 4, Kp = (K + 7) & ~7;
  const int cb = a.cbase[g];
  uint32_t* patch = reinterpret_cast<uint32_t*>(p2t_sm);          // [Kp][PS] tf32: k = ci*4 + rowpar*2 + colpar
  uint32_t* Ws = patch + Kp * PS;                                 // [Kp][WS] tf32
  const long long HW = (long long)a.H * a.W;
  __shared__ int s_chan[32];
  if (threadIdx.x < cnt) s_chan[threadIdx.x] = a.chan[cb + threadIdx.x];
  __syncthreads();
  // Both fills issue a batch of independent loads before the first store (a load -> convert -> store loop is one L2 round
  // trip per iteration: 14 + 11 serialized round trips were two thirds of this kernel's time).
  constexpr int GU = 7;
  for (int base = threadIdx.x; base < Kp * 32; base += GU * 128) {   // k-row x 32 token pairs
    float v0[GU], v1[GU];
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      const int idx = base + u * 128;
      const int e2 = idx & 31, kk = idx >> 5;                       // kk = ci*4 + rp*2 + colpar, token pair (2 e2, 2 e2 + 1)
      v0[u] = v1[u] = 0.f;
      if (idx < Kp * 32 && kk < K) {
        const int ci = kk >> 2, rp = (kk >> 1) & 1, cp = kk & 1;
        const float* img = a.img_in + s_chan[ci] * HW;
        const int tk = t0 + 2 * e2, i0 = tk / W0, j0 = tk - i0 * W0;   // W0 is even: both tokens of the pair lie in one row
        v0[u] = img[(long long)(2 * i0 + rp) * a.W + 2 * j0 + cp];
        v1[u] = img[(long long)(2 * i0 + rp) * a.W + 2 * j0 + 2 + cp];
      }
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      const int idx = base + u * 128;
      if (idx < Kp * 32) *reinterpret_cast<uint2*>(patch + (idx >> 5) * PS + 2 * (idx & 31)) = make_uint2(to_tf32(v0[u]), to_tf32(v1[u]));
    }
  }
  const float* wsrc = a.Wp + (long long)cb * 4 * D;
  constexpr int WU = 6;
  for (int base = threadIdx.x; base < Kp * (D / 4); base += WU * 128) {
    float4 w[WU];
#pragma unroll
    for (int u = 0; u < WU; ++u) {
      const int idx = base + u * 128;
      const int kk = idx / (D / 4), c4 = idx - kk * (D / 4);
      w[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < Kp * (D / 4) && kk < K) w[u] = __ldg(reinterpret_cast<const float4*>(wsrc + (long long)kk * D) + c4);
    }
#pragma unroll
    for (int u = 0; u < WU; ++u) {
      const int idx = base + u * 128;
      const int kk = idx / (D / 4), c4 = idx - kk * (D / 4);
      if (idx < Kp * (D / 4))
        *reinterpret_cast<uint4*>(Ws + kk * WS + 4 * c4) = make_uint4(to_tf32(w[u].x), to_tf32(w[u].y), to_tf32(w[u].z), to_tf32(w[u].w));
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  const int m0 = warp * 16;
  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const float b0 = a.bias ? a.bias[g * D + n * 8 + 2 * tq] : 0.f, b1 = a.bias ? a.bias[g * D + n * 8 + 2 * tq + 1] : 0.f;
    acc[n][0] = b0; acc[n][1] = b1; acc[n][2] = b0; acc[n][3] = b1;
  }
  for (int k0 = 0; k0 < Kp; k0 += 8) {
    const uint32_t a0 = patch[(k0 + tq) * PS + m0 + gq], a1 = patch[(k0 + tq) * PS + m0 + gq + 8];
    const uint32_t a2 = patch[(k0 + tq + 4) * PS + m0 + gq], a3 = patch[(k0 + tq + 4) * PS + m0 + gq + 8];
#pragma unroll
    for (int n = 0; n < NT; ++n)
      mma_tf32_1688(acc[n], a0, a1, a2, a3, Ws[(k0 + tq) * WS + n * 8 + gq], Ws[(k0 + tq + 4) * WS + n * 8 + gq]);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long o = ((long long)g * L0 + t0 + m0 + gq + 8 * h) * D + 2 * tq;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      float2 v = make_float2(acc[n][2 * h], acc[n][2 * h + 1]);
      if (a.ape) { const float2 p2 = __ldg(reinterpret_cast<const float2*>(a.ape + o + n * 8)); v.x += p2.x; v.y += p2.y; }
      *reinterpret_cast<float2*>(a.tok_out + o + n * 8) = v;
    }
  }
}

template <int NT>
static void launch_p2t_t(const PatchArgs& a, cudaStream_t s) {
  const int max_cnt = a.max_cnt > 0 ? a.max_cnt : 32;
  const int Kp = (max_cnt * 4 + 7) & ~7;
  const size_t smem = (size_t)Kp * ((P2T_TOK + 8) + (NT * 8 + 8)) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(p2t_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  const int L0 = (a.H / 2) * (a.W / 2);
  launch_kernel(p2t_kernel<NT>, dim3(L0 / P2T_TOK, a.G), dim3(128), smem, s, a);
}
bool p2t_supported(int D) { return D == 32 || D == 64 || D == 96 || D == 128; }
void launch_p2t(const PatchArgs& a, cudaStream_t s) {
  switch (a.D / 8) {
    case 4: launch_p2t_t<4>(a, s); break;
    case 8: launch_p2t_t<8>(a, s); break;
    case 12: launch_p2t_t<12>(a, s); break;
    case 16: launch_p2t_t<16>(a, s); break;
    default: break;
  }
}

constexpr int T2P_TOK = 128;        // tokens per CTA (8 warps x one m16 tile); the tile may span token rows
constexpr int T2P_NT = 7;           // n-tiles of 8 (= 2 output channels x 4 pixel positions) per pass

// img_out[chan[slot]][2i+p1][2j+p2] = bias[slot] + sum_c X[tok][c] Wp[slot*4 + p1*2 + p2][c]: per warp 16 tokens x (4 cnt) outputs,
// K = D.  The accumulator fragment (row = token, columns 2t, 2t+1) is a horizontally adjacent pixel pair: float2 stores.
__global__ void __launch_bounds__(256) t2p_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) float t2p_sm[];
  const int g = blockIdx.y;
  const int W0 = a.W >> 1;
  const int L0 = (a.H >> 1) * W0;
  const int t0 = blockIdx.x * T2P_TOK;
  const int D = a.D, RS = D + 4, D4 = D >> 2;         // (4 g + t) % 32 distinct -> conflict-free A and B fragments
  uint32_t* Xs = reinterpret_cast<uint32_t*>(t2p_sm);             // [128][RS] tf32
  uint32_t* Ws = Xs + T2P_TOK * RS;                               // [Np][RS] tf32, row n = slot*4 + p1*2 + p2
  pdl_launch_dependents();
  pdl_wait();
  const int cb = a.cbase[g], cnt = a.kcnt[g];
  const int N = cnt * 4, Np = (N + 7) & ~7;
  __shared__ int s_chan[32];
  __shared__ float s_bias[32];
  if (threadIdx.x < cnt) {
    s_chan[threadIdx.x] = a.chan[cb + threadIdx.x];
    s_bias[threadIdx.x] = a.bias ? a.bias[cb + threadIdx.x] : 0.f;
  }
  const float4* src = reinterpret_cast<const float4*>(a.tok_in + ((long long)g * L0 + t0) * D);
  constexpr int FU = 6;                                // batches of independent loads, then the converts / stores
  for (int base = threadIdx.x; base < T2P_TOK * D4; base += FU * 256) {
    float4 v[FU];
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const int idx = base + u * 256;
      v[u] = idx < T2P_TOK * D4 ? __ldg(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const int idx = base + u * 256;
      const int jj = idx / D4, c4 = idx - jj * D4;
      if (idx < T2P_TOK * D4)
        *reinterpret_cast<uint4*>(Xs + jj * RS + 4 * c4) = make_uint4(to_tf32(v[u].x), to_tf32(v[u].y), to_tf32(v[u].z), to_tf32(v[u].w));
    }
  }
  const float4* wsrc = reinterpret_cast<const float4*>(a.Wp + (long long)cb * 4 * D);
  for (int base = threadIdx.x; base < Np * D4; base += FU * 256) {
    float4 w[FU];
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const int idx = base + u * 256;
      w[u] = (idx < Np * D4 && idx / D4 < N) ? __ldg(wsrc + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const int idx = base + u * 256;
      const int row = idx / D4, c4 = idx - row * D4;
      if (idx < Np * D4)
        *reinterpret_cast<uint4*>(Ws + row * RS + 4 * c4) = make_uint4(to_tf32(w[u].x), to_tf32(w[u].y), to_tf32(w[u].z), to_tf32(w[u].w));
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  const int m0 = warp * 16;
  const long long HW = (long long)a.H * a.W;
  long long pix[2];                                    // pixel (2i + p1, 2j) of this thread's two tokens, p1 = (2 tq) >> 1 = tq >> 0 ...
  const int p1 = (2 * tq) >> 1 & 1;                    // columns 2tq, 2tq+1 of an n-tile: pixel position pp = (2 tq) & 3 -> p1 = pp >> 1, p2 = 0 / 1
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int tok = t0 + m0 + gq + 8 * h, i = tok / W0, j = tok - i * W0;
    pix[h] = (long long)(2 * i + p1) * a.W + 2 * j;
  }
  for (int n0 = 0; n0 < Np; n0 += 8 * T2P_NT) {
    float acc[T2P_NT][4];
#pragma unroll
    for (int n = 0; n < T2P_NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    for (int k0 = 0; k0 < D; k0 += 8) {
      const uint32_t a0 = Xs[(m0 + gq) * RS + k0 + tq], a1 = Xs[(m0 + gq + 8) * RS + k0 + tq];
      const uint32_t a2 = Xs[(m0 + gq) * RS + k0 + tq + 4], a3 = Xs[(m0 + gq + 8) * RS + k0 + tq + 4];
#pragma unroll
      for (int n = 0; n < T2P_NT; ++n) {
        if (n0 + 8 * n < Np)                           // block-uniform
          mma_tf32_1688(acc[n], a0, a1, a2, a3, Ws[(n0 + 8 * n + gq) * RS + k0 + tq], Ws[(n0 + 8 * n + gq) * RS + k0 + tq + 4]);
      }
    }
#pragma unroll
    for (int n = 0; n < T2P_NT; ++n) {
      const int col = n0 + 8 * n + 2 * tq;             // output column = slot*4 + pp
      if (col < N) {
        const int slot = col >> 2;
        const float bb = s_bias[slot];
        float* dst = a.img_out + s_chan[slot] * HW;
        *reinterpret_cast<float2*>(dst + pix[0]) = make_float2(acc[n][0] + bb, acc[n][1] + bb);
        *reinterpret_cast<float2*>(dst + pix[1]) = make_float2(acc[n][2] + bb, acc[n][3] + bb);
      }
    }
  }
}

void launch_t2p(const PatchArgs& a, cudaStream_t s) {
  const int max_cnt = a.max_cnt > 0 ? a.max_cnt : 32;
  const int Np = (max_cnt * 4 + 7) & ~7;
  const size_t smem = (size_t)(T2P_TOK + Np) * (a.D + 4) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(t2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  dim3 grid((a.H / 2) * (a.W / 2) / T2P_TOK, a.G);
  launch_kernel(t2p_kernel, grid, dim3(256), smem, s, a);
}

// =============================================================================================
// Per-channel affine
// =============================================================================================
__global__ void chan_affine_kernel(float* out, const float* a, const float* sa, const float* b, const float* sb, const float* t,
                                   int C, long long HW) {
  const int c = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  const float fa = sa ? sa[c] : 1.f, fb = sb ? sb[c] : 1.f, ft = t ? t[c] : 0.f;
  const long long base = (long long)c * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    float v = a[base + p] * fa + ft;
    if (b) v += b[base + p] * fb;
    out[base + p] = v;
  }
}
void launch_chan_affine(float* out, const float* a, const float* sa, const float* b, const float* sb, const float* t, int C,
                        long long HW, cudaStream_t s) {
  dim3 grid((unsigned)((HW + 1023) / 1024 < 64 ? (HW + 1023) / 1024 : 64), C);
  launch_kernel(chan_affine_kernel, grid, dim3(256), 0, s, out, a, sa, b, sb, t, C, HW);
}

}  // namespace vv
