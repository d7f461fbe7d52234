// Memory-bound kernels of one U-shaped Swin network application and its hand-derived adjoint:
// LayerNorm fwd/bwd (with the PatchMerging gather / PatchExpand pixel-shuffle folded into the addressing),
// 4x4-window attention fwd/bwd (roll + window partition/reverse folded into the addressing, P recomputed in
// the backward), the 2x2/stride-2 patch operators, and the per-channel affine seams.
#include "ops.h"

namespace vv {

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

// Each lane owns NPL = C/32 elements of the row as NPL/VEC vectors of VEC consecutive floats:
// element (k, v) = column VEC*(lane + 32*k) + v, so every warp-wide access is one contiguous 128*VEC-byte segment.
template <int VEC> struct VecT;
template <> struct VecT<1> { typedef float T; };
template <> struct VecT<2> { typedef float2 T; };
template <> struct VecT<4> { typedef float4 T; };

template <int VEC>
VV_DEVINL void ldv(float* dst, const float* src) {
  typename VecT<VEC>::T t = *reinterpret_cast<const typename VecT<VEC>::T*>(src);
  const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
  for (int v = 0; v < VEC; ++v) dst[v] = f[v];
}
template <int VEC>
VV_DEVINL void stv(float* dst, const float* src) {
  typename VecT<VEC>::T t;
  float* f = reinterpret_cast<float*>(&t);
#pragma unroll
  for (int v = 0; v < VEC; ++v) f[v] = src[v];
  *reinterpret_cast<typename VecT<VEC>::T*>(dst) = t;
}
template <int VEC>
VV_DEVINL void stv_bf16(bf16* dst, const float* src) {
  if (VEC == 1) {
    dst[0] = __float2bfloat16(src[0]);
  } else if (VEC == 2) {
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16(src[0], src[1]);
  } else {
    uint2 w;
    w.x = pack_bf16(src[0], src[1]);
    w.y = pack_bf16(src[2], src[3]);
    *reinterpret_cast<uint2*>(dst) = w;
  }
}

template <int NPL, int VEC, int MAP>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const LnArgs a) {
  constexpr int NV = NPL / VEC;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (r >= a.rows) return;
  const float* x = a.x + (long long)b * a.x_bs;
  float v[NPL];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    ldv<VEC>(v + k * VEC, x + ln_elem_off<MAP>(r, VEC * (lane + 32 * k), a.C, a.gw, a.ld_x));
#pragma unroll
    for (int i = 0; i < VEC; ++i) s += v[k * VEC + i];
  }
  const float mean = warp_sum(s) / a.C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NPL; ++k) {
    const float d = v[k] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / a.C + a.eps);
  const float* g = a.gamma + (long long)b * a.gb_bs;
  const float* be = a.beta + (long long)b * a.gb_bs;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = VEC * (lane + 32 * k);
    float gg[VEC], bb[VEC], y[VEC];
    ldv<VEC>(gg, g + c);
    ldv<VEC>(bb, be + c);
#pragma unroll
    for (int i = 0; i < VEC; ++i) y[i] = (v[k * VEC + i] - mean) * rstd * gg[i] + bb[i];
    if (a.out_bf16) stv_bf16<VEC>(a.out_bf16 + (long long)b * a.ob_bs + (long long)r * a.ld_ob + c, y);
    if (a.out_f32) stv<VEC>(a.out_f32 + (long long)b * a.of_bs + (long long)r * a.ld_of + c, y);
  }
}

template <int NPL, int VEC, int MAP>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdArgs a) {
  constexpr int NV = NPL / VEC;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (r >= a.rows) return;
  const float* x = a.x + (long long)b * a.x_bs;
  const float* dy = a.dy + (long long)b * a.dy_bs + (long long)r * a.ld_dy;
  const float* g = a.gamma + (long long)b * a.gb_bs;
  float v[NPL], gd[NPL];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    ldv<VEC>(v + k * VEC, x + ln_elem_off<MAP>(r, VEC * (lane + 32 * k), a.C, a.gw, a.ld_x));
    ldv<VEC>(gd + k * VEC, dy + VEC * (lane + 32 * k));
#pragma unroll
    for (int i = 0; i < VEC; ++i) s += v[k * VEC + i];
  }
  const float mean = warp_sum(s) / a.C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NPL; ++k) {
    v[k] -= mean;
    q += v[k] * v[k];
  }
  const float rstd = rsqrtf(warp_sum(q) / a.C + a.eps);
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float gg[VEC];
    ldv<VEC>(gg, g + VEC * (lane + 32 * k));
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int e = k * VEC + i;
      v[e] *= rstd;                     // xhat
      gd[e] *= gg[i];
      m1 += gd[e];
      m2 += gd[e] * v[e];
    }
  }
  m1 = warp_sum(m1) / a.C;
  m2 = warp_sum(m2) / a.C;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = VEC * (lane + 32 * k);
    float d[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) d[i] = (gd[k * VEC + i] - m1 - v[k * VEC + i] * m2) * rstd;
    if (a.dres) {
      float rr[VEC];
      ldv<VEC>(rr, a.dres + (long long)b * a.dres_bs + ln_elem_off<MAP>(r, c, a.C, a.gw, a.ld_dres));
#pragma unroll
      for (int i = 0; i < VEC; ++i) d[i] += rr[i];
    }
    stv<VEC>(a.dx + (long long)b * a.dx_bs + ln_elem_off<MAP>(r, c, a.C, a.gw, a.ld_dx), d);
    if (a.dx_bf16) stv_bf16<VEC>(a.dx_bf16 + (long long)b * a.dxb_bs + ln_elem_off<MAP>(r, c, a.C, a.gw, a.ld_dxb), d);
  }
}

#define VV_LN_DISPATCH(KERNEL, ARGS)                                                     \
  {                                                                                      \
    const int npl = ARGS.C / 32;                                                         \
    dim3 grid((ARGS.rows + 7) / 8, ARGS.batch);                                          \
    switch (ARGS.map * 100 + npl) {                                                      \
      case 2: KERNEL<2, 2, MAP_PLAIN><<<grid, 256, 0, s>>>(ARGS); break;                 \
      case 3: KERNEL<3, 1, MAP_PLAIN><<<grid, 256, 0, s>>>(ARGS); break;                 \
      case 4: KERNEL<4, 4, MAP_PLAIN><<<grid, 256, 0, s>>>(ARGS); break;                 \
      case 6: KERNEL<6, 2, MAP_PLAIN><<<grid, 256, 0, s>>>(ARGS); break;                 \
      case 12: KERNEL<12, 4, MAP_PLAIN><<<grid, 256, 0, s>>>(ARGS); break;               \
      case 36: KERNEL<36, 4, MAP_PLAIN><<<grid, 256, 0, s>>>(ARGS); break;               \
      case 108: KERNEL<8, 4, MAP_MERGE><<<grid, 256, 0, s>>>(ARGS); break;               \
      case 112: KERNEL<12, 4, MAP_MERGE><<<grid, 256, 0, s>>>(ARGS); break;              \
      case 202: KERNEL<2, 2, MAP_EXPAND><<<grid, 256, 0, s>>>(ARGS); break;              \
      case 203: KERNEL<3, 1, MAP_EXPAND><<<grid, 256, 0, s>>>(ARGS); break;              \
      default: break;                                                                    \
    }                                                                                    \
  }

bool ln_supported(int map, int C) {
  if (C % 32) return false;
  const int key = map * 100 + C / 32;
  switch (key) {
    case 2: case 3: case 4: case 6: case 12: case 36: case 108: case 112: case 202: case 203: return true;
    default: return false;
  }
}

void launch_ln_fwd(const LnArgs& a, cudaStream_t s) { VV_LN_DISPATCH(ln_fwd_kernel, a) }
void launch_ln_bwd(const LnBwdArgs& a, cudaStream_t s) { VV_LN_DISPATCH(ln_bwd_kernel, a) }

// =============================================================================================
// Window attention: one warp per (window, head); N = 16 tokens per window.
// =============================================================================================
template <int HD, bool BWD>
struct AttnSmem {
  static constexpr int W2 = HD / 2;             // 32-bit words (bf16 pairs) per token row
  static constexpr int RS = W2 + 2;             // padded row stride (even: 8-byte cp.async rows; 2i+w banks: conflict-free)
  static constexpr int MAT = 16 * RS;           // one 16 x HD operand
  static constexpr int WORDS = (BWD ? 4 : 3) * MAT + (BWD ? 2 : 1) * 16 * 17;
  static constexpr int BYTES = 4 * WORDS * 4;   // 4 warps per CTA
};

template <int HD, bool BWD>
__global__ void __launch_bounds__(128) attn_kernel(const AttnArgs a) {
  using L = AttnSmem<HD, BWD>;
  constexpr int W2 = L::W2, RS = L::RS;
  extern __shared__ uint32_t attn_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nww = a.gw >> 2, nwh = a.gh >> 2;
  const int item = blockIdx.x * 4 + warp;
  if (item >= nww * nwh * a.heads) return;      // warp-uniform; only __syncwarp below
  const int b = blockIdx.y;
  const int win = item / a.heads, h = item - win * a.heads;
  const int wi = win / nww, wj = win - wi * nww;
  const int d = a.heads * HD;

  uint32_t* Qs = attn_sm + warp * L::WORDS;
  uint32_t* Ks = Qs + L::MAT;
  uint32_t* Vs = Ks + L::MAT;
  uint32_t* dOs = Vs + L::MAT;                                   // BWD only
  float* P = reinterpret_cast<float*>(Qs + (BWD ? 4 : 3) * L::MAT);
  float* dS = P + 16 * 17;                                       // BWD only

  // original-grid token index of window-local token t (roll by -shift folded in; swinblock.py:275, 297)
  auto tok_of = [&](int t) {
    int row = 4 * wi + (t >> 2) + a.shift; if (row >= a.gh) row -= a.gh;
    int col = 4 * wj + (t & 3) + a.shift;  if (col >= a.gw) col -= a.gw;
    return row * a.gw + col;
  };

  // Stage Q, K, V (and dO) with 8-byte cp.async: every copy of the warp is in flight before the single wait.
  const bf16* qkv = a.qkv + (long long)b * a.qkv_bs;
  {
    constexpr int CH = W2 / 2;                                   // 8-byte chunks per operand row
    constexpr int NMAT = BWD ? 4 : 3;
    for (int idx = lane; idx < 16 * 4 * CH; idx += 32) {
      const int ch = idx % CH, rowm = idx / CH, m = rowm & 3, t = rowm >> 2;     // m: 0 Q, 1 K, 2 V, 3 dO
      if (m >= NMAT) continue;
      const bf16* src = m < 3 ? qkv + (long long)tok_of(t) * a.ld_qkv + m * d + h * HD
                              : a.dout + (long long)b * a.o_bs + (long long)tok_of(t) * a.ld_o + h * HD;
      uint32_t* dst = Qs + m * L::MAT + t * RS + 2 * ch;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src + 4 * ch) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncwarp();

  // ---- S = scale * Q K^T + bias + mask ; P = softmax(S) : lane -> row i, 8 columns ----
  const int i = lane >> 1, jh = lane & 1;
  const float scale = rsqrtf((float)HD);
  float sc[8];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) sc[jj] = 0.f;
#pragma unroll 4
  for (int w = 0; w < W2; ++w) {
    const float2 q2 = unpack_bf16(Qs[i * RS + w]);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float2 k2 = unpack_bf16(Ks[(jh * 8 + jj) * RS + w]);
      sc[jj] = fmaf(q2.x, k2.x, fmaf(q2.y, k2.y, sc[jj]));
    }
  }
  const float* rb = a.relbias + (long long)b * a.relbias_bs + (h * 16 + i) * 16 + jh * 8;
  const bool masked_win = a.shift > 0 && wi == nwh - 1;          // swinblock.py:236-260: latitude bands only
  float mx = -3.0e38f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = jh * 8 + jj;
    float sv = sc[jj] * scale + rb[jj];
    if (masked_win && ((i >> 2) < 2) != ((j >> 2) < 2)) sv += -100.0f;
    sc[jj] = sv;
    mx = fmaxf(mx, sv);
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  float sum = 0.f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    sc[jj] = expf(sc[jj] - mx);
    sum += sc[jj];
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    sc[jj] *= inv;
    P[i * 17 + jh * 8 + jj] = sc[jj];
  }

  if (!BWD) {
    __syncwarp();
    // ---- O = P V : lane -> channel pair, all 16 rows ----
    bf16* out = a.out + (long long)b * a.o_bs;
    for (int w = lane; w < W2; w += 32) {
      float2 o[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) o[r] = make_float2(0.f, 0.f);
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
        const float2 v2 = unpack_bf16(Vs[j * RS + w]);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float p = P[r * 17 + j];
          o[r].x = fmaf(p, v2.x, o[r].x);
          o[r].y = fmaf(p, v2.y, o[r].y);
        }
      }
#pragma unroll
      for (int r = 0; r < 16; ++r)
        reinterpret_cast<uint32_t*>(out + (long long)tok_of(r) * a.ld_o + h * HD)[w] = pack_bf16(o[r].x, o[r].y);
    }
  } else {
    // ---- dP = dO V^T ; dS = P o (dP - rowsum(dP o P)) ----
    float dp[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) dp[jj] = 0.f;
#pragma unroll 4
    for (int w = 0; w < W2; ++w) {
      const float2 o2 = unpack_bf16(dOs[i * RS + w]);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float2 v2 = unpack_bf16(Vs[(jh * 8 + jj) * RS + w]);
        dp[jj] = fmaf(o2.x, v2.x, fmaf(o2.y, v2.y, dp[jj]));
      }
    }
    float rs = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) rs = fmaf(dp[jj], sc[jj], rs);
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) dS[i * 17 + jh * 8 + jj] = sc[jj] * (dp[jj] - rs);
    __syncwarp();

    bf16* dqkv = a.dqkv + (long long)b * a.qkv_bs;
    for (int w = lane; w < W2; w += 32) {
      float2 acc[16];
      // dV[j] = sum_i P[i][j] dO[i]
#pragma unroll
      for (int r = 0; r < 16; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll 4
      for (int ii = 0; ii < 16; ++ii) {
        const float2 o2 = unpack_bf16(dOs[ii * RS + w]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p = P[ii * 17 + j];
          acc[j].x = fmaf(p, o2.x, acc[j].x);
          acc[j].y = fmaf(p, o2.y, acc[j].y);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        reinterpret_cast<uint32_t*>(dqkv + (long long)tok_of(j) * a.ld_qkv + 2 * d + h * HD)[w] = pack_bf16(acc[j].x, acc[j].y);
      // dQ[i] = scale * sum_j dS[i][j] K[j]
#pragma unroll
      for (int r = 0; r < 16; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
        const float2 k2 = unpack_bf16(Ks[j * RS + w]);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float g = dS[r * 17 + j];
          acc[r].x = fmaf(g, k2.x, acc[r].x);
          acc[r].y = fmaf(g, k2.y, acc[r].y);
        }
      }
#pragma unroll
      for (int r = 0; r < 16; ++r)
        reinterpret_cast<uint32_t*>(dqkv + (long long)tok_of(r) * a.ld_qkv + h * HD)[w] = pack_bf16(acc[r].x * scale, acc[r].y * scale);
      // dK[j] = scale * sum_i dS[i][j] Q[i]
#pragma unroll
      for (int r = 0; r < 16; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll 4
      for (int ii = 0; ii < 16; ++ii) {
        const float2 q2 = unpack_bf16(Qs[ii * RS + w]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float g = dS[ii * 17 + j];
          acc[j].x = fmaf(g, q2.x, acc[j].x);
          acc[j].y = fmaf(g, q2.y, acc[j].y);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        reinterpret_cast<uint32_t*>(dqkv + (long long)tok_of(j) * a.ld_qkv + d + h * HD)[w] = pack_bf16(acc[j].x * scale, acc[j].y * scale);
    }
  }
}

template <int HD, bool BWD>
static void launch_attn_t(const AttnArgs& a, cudaStream_t s) {
  using L = AttnSmem<HD, BWD>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_kernel<HD, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
    attr_set = true;
  }
  const int items = (a.gh / 4) * (a.gw / 4) * a.heads;
  dim3 grid((items + 3) / 4, a.batch);
  attn_kernel<HD, BWD><<<grid, 128, L::BYTES, s>>>(a);
}
void launch_attn_fwd(const AttnArgs& a, cudaStream_t s) {
  if (a.hd == 32) launch_attn_t<32, false>(a, s);
  else if (a.hd == 192) launch_attn_t<192, false>(a, s);
}
void launch_attn_bwd(const AttnArgs& a, cudaStream_t s) {
  if (a.hd == 32) launch_attn_t<32, true>(a, s);
  else if (a.hd == 192) launch_attn_t<192, true>(a, s);
}

// =============================================================================================
// Patch operators
// =============================================================================================
constexpr int P2T_TOK = 16;
constexpr int P2T_KMAX = 128;

__global__ void __launch_bounds__(128) p2t_kernel(const PatchArgs a) {
  __shared__ float patch[P2T_TOK][P2T_KMAX];
  const int g = blockIdx.y;
  const int W0 = a.W >> 1;
  const int L0 = (a.H >> 1) * W0;
  const int t0 = blockIdx.x * P2T_TOK;
  const int K = a.kcnt[g] * 4;
  const int cb = a.cbase[g];
  const long long HW = (long long)a.H * a.W;
  for (int idx = threadIdx.x; idx < P2T_TOK * K; idx += blockDim.x) {
    const int tt = idx / K, k = idx - tt * K;
    const int tok = t0 + tt;
    const int i = tok / W0, j = tok - i * W0;
    const int ch = a.chan[cb + (k >> 2)];
    patch[tt][k] = a.img_in[ch * HW + (long long)(2 * i + ((k >> 1) & 1)) * a.W + 2 * j + (k & 1)];
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= a.D) return;
  float acc[P2T_TOK];
  const float b0 = a.bias ? a.bias[g * a.D + c] : 0.f;
#pragma unroll
  for (int tt = 0; tt < P2T_TOK; ++tt) acc[tt] = b0;
  const float* w = a.Wp + (long long)cb * 4 * a.D + c;
  for (int k = 0; k < K; ++k) {
    const float wv = w[(long long)k * a.D];
#pragma unroll
    for (int tt = 0; tt < P2T_TOK; ++tt) acc[tt] = fmaf(patch[tt][k], wv, acc[tt]);
  }
#pragma unroll
  for (int tt = 0; tt < P2T_TOK; ++tt) {
    const long long o = ((long long)g * L0 + t0 + tt) * a.D + c;
    a.tok_out[o] = acc[tt] + (a.ape ? a.ape[o] : 0.f);
  }
}

void launch_p2t(const PatchArgs& a, cudaStream_t s) {
  const int L0 = (a.H / 2) * (a.W / 2);
  dim3 grid(L0 / P2T_TOK, a.G);
  p2t_kernel<<<grid, ((a.D + 31) / 32) * 32, 0, s>>>(a);
}

constexpr int T2P_TOK = 32;

__global__ void __launch_bounds__(128) t2p_kernel(const PatchArgs a) {
  extern __shared__ float Xs[];                       // [32][D+1]
  const int g = blockIdx.z;
  const int W0 = a.W >> 1;
  const int L0 = (a.H >> 1) * W0;
  const int i = blockIdx.y;
  const int j0 = blockIdx.x * T2P_TOK;
  const int D = a.D, RS = D + 1;
  const float* src = a.tok_in + ((long long)g * L0 + (long long)i * W0 + j0) * D;
  for (int idx = threadIdx.x; idx < T2P_TOK * D; idx += blockDim.x) {
    const int jj = idx / D, c = idx - jj * D;
    Xs[jj * RS + c] = src[idx];
  }
  __syncthreads();
  const int p1 = threadIdx.x >> 6, xx = threadIdx.x & 63, jj = xx >> 1, p2 = xx & 1;
  const int cb = a.cbase[g], cnt = a.kcnt[g];
  const long long HW = (long long)a.H * a.W;
  const float* xr = Xs + jj * RS;
  for (int slot = 0; slot < cnt; ++slot) {
    const float* w = a.Wp + ((long long)(cb + slot) * 4 + p1 * 2 + p2) * D;
    float acc = a.bias ? a.bias[cb + slot] : 0.f;
#pragma unroll 8
    for (int c = 0; c < D; ++c) acc = fmaf(xr[c], __ldg(w + c), acc);
    const int ch = a.chan[cb + slot];
    a.img_out[ch * HW + (long long)(2 * i + p1) * a.W + 2 * j0 + xx] = acc;
  }
}

void launch_t2p(const PatchArgs& a, cudaStream_t s) {
  dim3 grid((a.W / 2) / T2P_TOK, a.H / 2, a.G);
  t2p_kernel<<<grid, 128, T2P_TOK * (a.D + 1) * sizeof(float), s>>>(a);
}

// =============================================================================================
// Per-channel affine
// =============================================================================================
__global__ void chan_affine_kernel(float* out, const float* a, const float* sa, const float* b, const float* sb, const float* t,
                                   int C, long long HW) {
  const int c = blockIdx.y;
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
  chan_affine_kernel<<<grid, 256, 0, s>>>(out, a, sa, b, sb, t, C, HW);
}

}  // namespace vv
