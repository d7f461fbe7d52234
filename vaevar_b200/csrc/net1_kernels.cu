// Kernels of the forecast network LGUnet_all_1 (networks/LGUnet_all.py:743-777) beyond those it shares with LGUnet_all:
// 2-D rotary embedding, window / whole-grid attention with online softmax, the 3 x 2 / stride-2 patch embedding and the
// overlap-adding ConvTranspose2d head.  Forward only: the DA cycle never differentiates the forecast (da_4dvar.py:1329, 666-681).
#include <stdlib.h>

#include "ops.h"

namespace vv {
namespace {

VV_DEVINL void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
VV_DEVINL void ldsm_x4_trans(uint32_t (&r)[4], const void* smem_row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row_ptr)));
}
VV_DEVINL uint32_t pack_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

}  // namespace

// =============================================================================================
// rope2, in place on q and k
// =============================================================================================
__global__ void __launch_bounds__(256) rope_kernel(const RopeArgs a) {
  const int half = a.hd >> 1;
  const long long per_tok = (long long)a.heads * half;
  const long long n = (long long)a.gh * a.gw * per_tok;
  const int b = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  __half* base = reinterpret_cast<__half*>(a.qkv) + (long long)b * a.qkv_bs;
  const int d = a.heads * a.hd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tok = (int)(i / per_tok), rem = (int)(i - (long long)tok * per_tok);
    const int h = rem / half, j = rem - h * half;
    const int row = tok / a.gw, col = tok - row * a.gw;
    int rr = row - a.sh; if (rr < 0) rr += a.gh;                 // coordinates in the rolled frame (Attention.py:560)
    int cc = col - a.sw; if (cc < 0) cc += a.gw;
    const int pos = (rr % a.wh) * a.ww + (cc % a.ww);
    const float2 cs = __ldg(a.table + (long long)pos * half + j);
    __half* p = base + (long long)tok * a.ld_qkv + h * a.hd + j;
#pragma unroll
    for (int m = 0; m < 2; ++m) {                                // q, then k
      const float x1 = __half2float(p[m * d]), x2 = __half2float(p[m * d + half]);
      p[m * d] = __float2half_rn(x1 * cs.x - x2 * cs.y);
      p[m * d + half] = __float2half_rn(x2 * cs.x + x1 * cs.y);
    }
  }
}
void launch_rope(const RopeArgs& a, cudaStream_t s) {
  const long long n = (long long)a.gh * a.gw * a.heads * (a.hd / 2);
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 148 * 16);
  launch_kernel(rope_kernel, dim3(blocks, a.batch), dim3(256), 0, s, a);
}

// =============================================================================================
// Attention over windows of any size: one CTA = one (window, head, query tile of 16 x WARPS rows); keys / values stream through
// shared memory in tiles of KT tokens; S = Q K^T and O += P V are m16n8k16 tensor-core tiles (fp16 operands, fp32 accumulation),
// the softmax is the online (running max / running sum) form, so a window may be the whole 90 x 180 grid.
// =============================================================================================
template <int HD, int WARPS, int KT>
struct Attn1Smem {
  static constexpr int QR = 16 * WARPS;
  static constexpr int RS = HD / 2 + 4;           // padded row stride in 32-bit words (16-byte aligned, conflict-free fragments)
  static constexpr int WORDS = (QR + 2 * KT) * RS;
  static constexpr int BYTES = WORDS * 4;
};

template <int HD, int WARPS, int KT>
__global__ void __launch_bounds__(WARPS * 32) attn1_kernel(const Attn1Args a) {
  using L = Attn1Smem<HD, WARPS, KT>;
  constexpr int RS = L::RS, QR = L::QR, CH = HD / 8;            // CH: 16-byte chunks per row
  extern __shared__ __align__(16) uint32_t at1_sm[];
  uint32_t* Qs = at1_sm;
  uint32_t* Ks = Qs + QR * RS;
  uint32_t* Vs = Ks + KT * RS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int N = a.wh * a.ww;
  const int nq = (N + QR - 1) / QR;
  const int nww = a.gw / a.ww;
  int item = blockIdx.x;
  const int qt = item % nq; item /= nq;
  const int h = item % a.heads; const int win = item / a.heads;
  const int wi = win / nww, wj = win - wi * nww;
  const int b = blockIdx.y;
  const int d = a.heads * HD;
  pdl_launch_dependents();
  pdl_wait();
  const __half* qkv = reinterpret_cast<const __half*>(a.qkv) + (long long)b * a.qkv_bs;
  auto tok_of = [&](int n) {
    const int r = n / a.ww, c = n - r * a.ww;
    int row = wi * a.wh + r + a.sh; if (row >= a.gh) row -= a.gh;
    int col = wj * a.ww + c + a.sw; if (col >= a.gw) col -= a.gw;
    return row * a.gw + col;
  };
  // rows [n0, n0 + rows) of matrix m (0 q, 1 k, 2 v) -> dst; rows beyond N are zero-filled
  auto stage = [&](uint32_t* dst, int m, int n0, int rows) {
    for (int idx = threadIdx.x; idx < rows * CH; idx += WARPS * 32) {
      const int r = idx / CH, ch = idx - r * CH;
      uint32_t* sp = dst + r * RS + 4 * ch;
      if (n0 + r < N) {
        const __half* src = qkv + (long long)tok_of(n0 + r) * a.ld_qkv + m * d + h * HD + 8 * ch;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sp)), "l"(src) : "memory");
      } else {
        *reinterpret_cast<uint4*>(sp) = make_uint4(0, 0, 0, 0);
      }
    }
  };
  const int q0 = qt * QR;
  stage(Qs, 0, q0, QR);
  const int m0 = warp * 16;
  // latitude bands of a masked window (the last window row of a shifted frame): local rows < wh - sh vs the rest
  const bool masked = a.mask && a.sh > 0 && wi == a.gh / a.wh - 1;
  const int band_row = a.wh - a.sh;
  const int qi0 = q0 + m0 + g, qi1 = qi0 + 8;                    // this thread's two query rows (window-local indices)
  const int qb0 = (qi0 / a.ww) < band_row ? 0 : 1, qb1 = (qi1 / a.ww) < band_row ? 0 : 1;

  float o[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float mrun0 = -INFINITY, mrun1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sc = a.scale * 1.4426950408889634f;              // softmax in base 2

  for (int kv0 = 0; kv0 < N; kv0 += KT) {
    __syncthreads();                                            // the previous tile's K / V are no longer read
    stage(Ks, 1, kv0, KT);
    stage(Vs, 2, kv0, KT);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- S = Q K^T ----
    float s[KT / 8][4];
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const int w = ks * 8 + t;
      const uint32_t af[4] = {Qs[(m0 + g) * RS + w], Qs[(m0 + g + 8) * RS + w], Qs[(m0 + g) * RS + w + 4], Qs[(m0 + g + 8) * RS + w + 4]};
#pragma unroll
      for (int j = 0; j < KT / 8; ++j) mma_f16(s[j], af, Ks[(8 * j + g) * RS + w], Ks[(8 * j + g) * RS + w + 4]);
    }
    // ---- scale, mask, online softmax ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kn = kv0 + 8 * j + 2 * t + e;
        bool dead = kn >= N;
        bool dead0 = dead, dead1 = dead;
        if (masked && !dead) {
          const int kb = (kn / a.ww) < band_row ? 0 : 1;
          dead0 = kb != qb0; dead1 = kb != qb1;
        }
        s[j][e] = dead0 ? -INFINITY : s[j][e] * sc;
        s[j][2 + e] = dead1 ? -INFINITY : s[j][2 + e] * sc;
        mx0 = fmaxf(mx0, s[j][e]); mx1 = fmaxf(mx1, s[j][2 + e]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(mrun0, mx0), mn1 = fmaxf(mrun1, mx1);
    const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;     // a row that has seen no live key yet
    const float c0 = ex2_approx(mrun0 - ms0), c1 = ex2_approx(mrun1 - ms1);                 // exp2(-inf) = 0
    mrun0 = mn0; mrun1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      s[j][0] = ex2_approx(s[j][0] - ms0); s[j][1] = ex2_approx(s[j][1] - ms0);
      s[j][2] = ex2_approx(s[j][2] - ms1); s[j][3] = ex2_approx(s[j][3] - ms1);
      ps0 += s[j][0] + s[j][1]; ps1 += s[j][2] + s[j][3];
    }
    l0 = l0 * c0 + ps0; l1 = l1 * c1 + ps1;
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
    // ---- O += P V ----
    const int lm = lane >> 3, lr = lane & 7;
#pragma unroll
    for (int kb = 0; kb < KT / 16; ++kb) {
      const uint32_t pa[4] = {pack_h2(s[2 * kb][0], s[2 * kb][1]), pack_h2(s[2 * kb][2], s[2 * kb][3]),
                              pack_h2(s[2 * kb + 1][0], s[2 * kb + 1][1]), pack_h2(s[2 * kb + 1][2], s[2 * kb + 1][3])};
#pragma unroll
      for (int c16 = 0; c16 < HD / 16; ++c16) {
        uint32_t vb[4];
        ldsm_x4_trans(vb, Vs + (16 * kb + (lm & 1) * 8 + lr) * RS + c16 * 8 + (lm >> 1) * 4);
        mma_f16(o[2 * c16], pa, vb[0], vb[1]);
        mma_f16(o[2 * c16 + 1], pa, vb[2], vb[3]);
      }
    }
  }
  // ---- normalise, stage through this warp's Q rows, store ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
  __syncwarp();
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    Qs[(m0 + g) * RS + n * 4 + t] = pack_h2(o[n][0] * i0, o[n][1] * i0);
    Qs[(m0 + g + 8) * RS + n * 4 + t] = pack_h2(o[n][2] * i1, o[n][3] * i1);
  }
  __syncwarp();
  __half* out = reinterpret_cast<__half*>(a.out) + (long long)b * a.o_bs;
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    const int r = idx / CH, ch = idx - r * CH;
    const int qi = q0 + m0 + r;
    if (qi < N)
      *reinterpret_cast<uint4*>(out + (long long)tok_of(qi) * a.ld_o + h * HD + 8 * ch) = *reinterpret_cast<const uint4*>(Qs + (m0 + r) * RS + 4 * ch);
  }
}

template <int HD, int WARPS, int KT>
static void launch_attn1_t(const Attn1Args& a, cudaStream_t s) {
  using L = Attn1Smem<HD, WARPS, KT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn1_kernel<HD, WARPS, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
    attr_set = true;
  }
  const int N = a.wh * a.ww, nq = (N + L::QR - 1) / L::QR;
  const long long items = (long long)(a.gh / a.wh) * (a.gw / a.ww) * a.heads * nq;
  launch_kernel(attn1_kernel<HD, WARPS, KT>, dim3((unsigned)items, a.batch), dim3(WARPS * 32), L::BYTES, s, a);
}
// =============================================================================================
// Long windows (the whole-grid first LG stage: 16 200 tokens): the same online-softmax attention with the operand traffic of a
// flash kernel -- one CTA = one (window, head, 128 query rows), eight warps of 16 rows; keys / values stream through a
// double-buffered cp.async ring of 64-token tiles (the loads of tile i + 1 fly while tile i is multiplied); Q, K fragments come
// from ldmatrix.x4 (one shared-memory instruction per two MMAs), V fragments from ldmatrix.x4.trans.
// =============================================================================================
VV_DEVINL void ldsm_x4(uint32_t (&r)[4], const void* smem_row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row_ptr)));
}

template <int HD>
struct Attn1LongSmem {
  static constexpr int WARPS = 8, QR = 128, KT = 64;
  static constexpr int RS = HD / 2 + 4;
  static constexpr int WORDS = (QR + 4 * KT) * RS;              // Q, K[2], V[2]
  static constexpr int BYTES = WORDS * 4;
};

template <int HD>
__global__ void __launch_bounds__(256, 1) attn1_long_kernel(const Attn1Args a) {
  using L = Attn1LongSmem<HD>;
  constexpr int RS = L::RS, QR = L::QR, KT = L::KT, CH = HD / 8;
  extern __shared__ __align__(16) uint32_t at1l_sm[];
  uint32_t* Qs = at1l_sm;
  uint32_t* KV = Qs + QR * RS;                                   // stage b: K at KV + b * 2 * KT * RS, V right behind it
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int N = a.wh * a.ww;
  const int nq = (N + QR - 1) / QR;
  const int nww = a.gw / a.ww;
  int item = blockIdx.x;
  const int qt = item % nq; item /= nq;
  const int h = item % a.heads; const int win = item / a.heads;
  const int wi = win / nww, wj = win - wi * nww;
  const int b = blockIdx.y;
  const int d = a.heads * HD;
  pdl_launch_dependents();
  pdl_wait();
  const __half* qkv = reinterpret_cast<const __half*>(a.qkv) + (long long)b * a.qkv_bs;
  auto tok_of = [&](int n) {
    const int r = n / a.ww, c = n - r * a.ww;
    int row = wi * a.wh + r + a.sh; if (row >= a.gh) row -= a.gh;
    int col = wj * a.ww + c + a.sw; if (col >= a.gw) col -= a.gw;
    return row * a.gw + col;
  };
  auto stage = [&](uint32_t* dst, int m, int n0, int rows) {
    for (int idx = threadIdx.x; idx < rows * CH; idx += 256) {
      const int r = idx / CH, ch = idx - r * CH;
      uint32_t* sp = dst + r * RS + 4 * ch;
      if (n0 + r < N) {
        const __half* src = qkv + (long long)tok_of(n0 + r) * a.ld_qkv + m * d + h * HD + 8 * ch;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sp)), "l"(src) : "memory");
      } else {
        *reinterpret_cast<uint4*>(sp) = make_uint4(0, 0, 0, 0);
      }
    }
  };
  const int q0 = qt * QR;
  const int nkt = (N + KT - 1) / KT;
  stage(Qs, 0, q0, QR);
  stage(KV, 1, 0, KT);
  stage(KV + KT * RS, 2, 0, KT);
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int m0 = warp * 16;
  const bool masked = a.mask && a.sh > 0 && wi == a.gh / a.wh - 1;
  const int band_row = a.wh - a.sh;
  const int qi0 = q0 + m0 + g, qi1 = qi0 + 8;
  const int qb0 = (qi0 / a.ww) < band_row ? 0 : 1, qb1 = (qi1 / a.ww) < band_row ? 0 : 1;
  const bool warp_live = q0 + m0 < N;                           // a warp whose 16 rows all lie beyond N only helps with the loads

  float o[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float mrun0 = -INFINITY, mrun1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sc = a.scale * 1.4426950408889634f;
  // ldmatrix lane addressing: A fragments of Q (rows m0.., 16 x 16 block at k-step ks) and B fragments of K (two 8-key n-tiles)
  const uint32_t* q_lane = Qs + (m0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * RS + 4 * (lane >> 4);
  const int k_lane = ((lane & 7) + 8 * (lane >> 4)) * RS + 4 * ((lane >> 3) & 1);
  const int lm = lane >> 3, lr = lane & 7;
  const int v_lane = ((lm & 1) * 8 + lr) * RS + (lm >> 1) * 4;

  for (int it = 0; it < nkt; ++it) {
    const int kv0 = it * KT;
    uint32_t* Ks = KV + (it & 1) * 2 * KT * RS;
    uint32_t* Vs = Ks + KT * RS;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                             // tile `it` has landed for everyone; tile it - 1 is no longer read
    if (it + 1 < nkt) {
      uint32_t* Kn = KV + ((it + 1) & 1) * 2 * KT * RS;
      stage(Kn, 1, kv0 + KT, KT);
      stage(Kn + KT * RS, 2, kv0 + KT, KT);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (!warp_live) continue;
    // ---- S = Q K^T ----
    float s[KT / 8][4];
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      uint32_t af[4];
      ldsm_x4(af, q_lane + ks * 8);
#pragma unroll
      for (int j = 0; j < KT / 8; j += 2) {
        uint32_t kb[4];
        ldsm_x4(kb, Ks + j * 8 * RS + k_lane + ks * 8);
        mma_f16(s[j], af, kb[0], kb[1]);
        mma_f16(s[j + 1], af, kb[2], kb[3]);
      }
    }
    // ---- scale, mask, online softmax ----
    const bool ragged = kv0 + KT > N;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float v0 = s[j][e] * sc, v1 = s[j][2 + e] * sc;
        if (ragged || masked) {
          const int kn = kv0 + 8 * j + 2 * t + e;
          bool dead0 = kn >= N, dead1 = dead0;
          if (masked && !dead0) {
            const int kb = (kn / a.ww) < band_row ? 0 : 1;
            dead0 = kb != qb0; dead1 = kb != qb1;
          }
          if (dead0) v0 = -INFINITY;
          if (dead1) v1 = -INFINITY;
        }
        s[j][e] = v0; s[j][2 + e] = v1;
        mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(mrun0, mx0), mn1 = fmaxf(mrun1, mx1);
    const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;
    const float c0 = ex2_approx(mrun0 - ms0), c1 = ex2_approx(mrun1 - ms1);
    mrun0 = mn0; mrun1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      s[j][0] = ex2_approx(s[j][0] - ms0); s[j][1] = ex2_approx(s[j][1] - ms0);
      s[j][2] = ex2_approx(s[j][2] - ms1); s[j][3] = ex2_approx(s[j][3] - ms1);
      ps0 += s[j][0] + s[j][1]; ps1 += s[j][2] + s[j][3];
    }
    l0 = l0 * c0 + ps0; l1 = l1 * c1 + ps1;
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
    // ---- O += P V ----
#pragma unroll
    for (int kb = 0; kb < KT / 16; ++kb) {
      const uint32_t pa[4] = {pack_h2(s[2 * kb][0], s[2 * kb][1]), pack_h2(s[2 * kb][2], s[2 * kb][3]),
                              pack_h2(s[2 * kb + 1][0], s[2 * kb + 1][1]), pack_h2(s[2 * kb + 1][2], s[2 * kb + 1][3])};
#pragma unroll
      for (int c16 = 0; c16 < HD / 16; ++c16) {
        uint32_t vb[4];
        ldsm_x4_trans(vb, Vs + 16 * kb * RS + v_lane + c16 * 8);
        mma_f16(o[2 * c16], pa, vb[0], vb[1]);
        mma_f16(o[2 * c16 + 1], pa, vb[2], vb[3]);
      }
    }
  }
  // ---- normalise, stage through this warp's Q rows, store ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
  __syncwarp();
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    Qs[(m0 + g) * RS + n * 4 + t] = pack_h2(o[n][0] * i0, o[n][1] * i0);
    Qs[(m0 + g + 8) * RS + n * 4 + t] = pack_h2(o[n][2] * i1, o[n][3] * i1);
  }
  __syncwarp();
  __half* out = reinterpret_cast<__half*>(a.out) + (long long)b * a.o_bs;
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    const int r = idx / CH, ch = idx - r * CH;
    const int qi = q0 + m0 + r;
    if (qi < N)
      *reinterpret_cast<uint4*>(out + (long long)tok_of(qi) * a.ld_o + h * HD + 8 * ch) = *reinterpret_cast<const uint4*>(Qs + (m0 + r) * RS + 4 * ch);
  }
}

template <int HD>
static void launch_attn1_long_t(const Attn1Args& a, cudaStream_t s) {
  using L = Attn1LongSmem<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn1_long_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
    attr_set = true;
  }
  const int N = a.wh * a.ww, nq = (N + L::QR - 1) / L::QR;
  const long long items = (long long)(a.gh / a.wh) * (a.gw / a.ww) * a.heads * nq;
  launch_kernel(attn1_long_kernel<HD>, dim3((unsigned)items, a.batch), dim3(256), L::BYTES, s, a);
}

bool attn1_supported(int hd) { return hd == 32 || hd == 64 || hd == 192; }
void launch_attn1(const Attn1Args& a, cudaStream_t s) {
  const int N = a.wh * a.ww;
  if (N <= 80) {                      // one key tile, five query warps: the 6 x 12 windows (72 tokens)
    if (a.hd == 32) launch_attn1_t<32, 5, 80>(a, s);
    else if (a.hd == 64) launch_attn1_t<64, 5, 80>(a, s);
    else if (a.hd == 192) launch_attn1_t<192, 5, 80>(a, s);
  } else {                            // long windows (the whole-grid first LG stage): 128-row query tiles, 64-key tiles, double-buffered
    if (a.hd == 32) launch_attn1_long_t<32>(a, s);
    else if (a.hd == 64) launch_attn1_long_t<64>(a, s);
    else if (a.hd == 192) launch_attn1_long_t<192>(a, s);
  }
}

// =============================================================================================
// Patch embedding, 3 x 2 kernel / stride 2
// =============================================================================================
constexpr int P32_TOK = 64;
__global__ void __launch_bounds__(256) patch32_kernel(const Patch32Args a) {
  extern __shared__ __align__(16) float p32_sm[];
  const int g = blockIdx.y;
  const int L0 = a.h0 * a.w0;
  const int t0 = blockIdx.x * P32_TOK;
  pdl_launch_dependents();
  pdl_wait();
  const int cnt = a.kcnt[g], cb = a.cbase[g];
  const int K = cnt * 6;
  float* patch = p32_sm;                         // [P32_TOK][K + 1]
  float* Ws = patch + P32_TOK * (K + 1);         // [K][D]
  const long long HW = (long long)a.H * a.W;
  for (int idx = threadIdx.x; idx < P32_TOK * K; idx += 256) {
    const int k = idx / P32_TOK, tk = idx - k * P32_TOK;          // consecutive threads -> consecutive tokens (stride-2 pixels)
    const int ci = k / 6, kr = (k - ci * 6) >> 1, kc = k & 1;
    const int tok = t0 + tk;
    float v = 0.f;
    if (tok < L0) {
      const int i = tok / a.w0, j = tok - i * a.w0;
      v = a.img[(cb + ci) * HW + (long long)(2 * i + kr) * a.W + 2 * j + kc];
    }
    patch[tk * (K + 1) + k] = v;
  }
  const float* wsrc = a.Wp + (long long)cb * 6 * a.D;
  for (int idx = threadIdx.x; idx < K * a.D; idx += 256) Ws[idx] = wsrc[idx];
  __syncthreads();
  for (int idx = threadIdx.x; idx < P32_TOK * a.D; idx += 256) {
    const int tk = idx / a.D, c = idx - tk * a.D;
    const int tok = t0 + tk;
    if (tok >= L0) continue;
    float acc = a.bias[g * a.D + c];
    const float* pr = patch + tk * (K + 1);
    for (int k = 0; k < K; ++k) acc = fmaf(pr[k], Ws[k * a.D + c], acc);
    const long long o = ((long long)g * L0 + tok) * a.D + c;
    a.tok[o] = acc + a.ape[o];
  }
}
void launch_patch32(const Patch32Args& a, cudaStream_t s) {
  const int K = a.max_cnt * 6;
  const size_t smem = (size_t)(P32_TOK * (K + 1) + K * a.D) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(patch32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  const int L0 = a.h0 * a.w0;
  launch_kernel(patch32_kernel, dim3((L0 + P32_TOK - 1) / P32_TOK, a.G), dim3(256), smem, s, a);
}

// =============================================================================================
// ConvTranspose2d head, 3 x 2 kernel / stride 2: output-stationary (no atomics) -- a CTA owns image rows 2 i and 2 i + 1 over 64
// patch columns of one group and reads patch rows i (kernel rows 0, 1) and i - 1 (kernel row 2, the overlap onto row 2 i).
// =============================================================================================
constexpr int CT_TOK = 64;
__global__ void __launch_bounds__(256) convt32_kernel(const ConvT32Args a) {
  extern __shared__ __align__(16) float ct_sm[];
  const int g = blockIdx.z, i = blockIdx.y, j0 = blockIdx.x * CT_TOK;
  const int D = a.D, DS = D + 1;
  pdl_launch_dependents();
  pdl_wait();
  const int cnt = a.kcnt[g], cb = a.cbase[g];
  float* X0 = ct_sm;                              // [CT_TOK][D + 1] patch row i      (zeros if i == h0)
  float* X1 = X0 + CT_TOK * DS;                   // [CT_TOK][D + 1] patch row i - 1  (zeros if i == 0)
  float* Ws = X1 + CT_TOK * DS;                   // [cnt][3][2][D]
  const long long L0 = (long long)a.h0 * a.w0;
  for (int idx = threadIdx.x; idx < CT_TOK * D; idx += 256) {
    const int tk = idx / D, c = idx - tk * D;
    const int j = j0 + tk;
    float v0 = 0.f, v1 = 0.f;
    if (j < a.w0) {
      if (i < a.h0) v0 = a.tok[((long long)g * L0 + (long long)i * a.w0 + j) * D + c];
      if (i > 0) v1 = a.tok[((long long)g * L0 + (long long)(i - 1) * a.w0 + j) * D + c];
    }
    X0[tk * DS + c] = v0; X1[tk * DS + c] = v1;
  }
  const float* wsrc = a.Wt + (long long)cb * 6 * D;
  for (int idx = threadIdx.x; idx < cnt * 6 * D; idx += 256) Ws[idx] = wsrc[idx];
  __syncthreads();
  const long long HW = (long long)a.H * a.W;
  const int npx = 2 * CT_TOK;
  // outputs of this CTA: cnt slots x 2 image rows x 128 pixels
  for (int idx = threadIdx.x; idx < cnt * 2 * npx; idx += 256) {
    const int px = idx % npx, rest = idx / npx, par = rest & 1, slot = rest >> 1;
    const int tk = px >> 1, kc = px & 1;
    const int x = 2 * j0 + px, y = 2 * i + par;
    if (x >= a.W || y >= a.H) continue;
    float acc = a.bias[cb + slot];
    const float* w = Ws + (long long)slot * 6 * D;
    if (par == 0) {                                // even row: kernel row 0 of patch row i, kernel row 2 of patch row i - 1
      const float *wa = w + (0 * 2 + kc) * D, *wb = w + (2 * 2 + kc) * D, *xa = X0 + tk * DS, *xb = X1 + tk * DS;
      for (int c = 0; c < D; ++c) acc = fmaf(xa[c], wa[c], fmaf(xb[c], wb[c], acc));
    } else {                                       // odd row: kernel row 1 of patch row i
      const float *wa = w + (1 * 2 + kc) * D, *xa = X0 + tk * DS;
      for (int c = 0; c < D; ++c) acc = fmaf(xa[c], wa[c], acc);
    }
    a.img[a.chan[cb + slot] * HW + (long long)y * a.W + x] = acc;
  }
}
void launch_convt32(const ConvT32Args& a, cudaStream_t s) {
  const size_t smem = (size_t)(2 * CT_TOK * (a.D + 1) + a.max_cnt * 6 * a.D) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(convt32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  launch_kernel(convt32_kernel, dim3((a.w0 + CT_TOK - 1) / CT_TOK, a.h0 + 1, a.G), dim3(256), smem, s, a);
}

}  // namespace vv
