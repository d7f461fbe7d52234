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
VV_DEVINL void ldsm_x4(uint32_t (&r)[4], const void* smem_row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row_ptr)));
}
VV_DEVINL uint32_t pack_h2_rn(float a, float b) {        // the two roundings of __float2half_rn, packed (a in the low half)
  const __half2 h = __halves2half2(__float2half_rn(a), __float2half_rn(b));
  return *reinterpret_cast<const uint32_t*>(&h);
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
// The 6 x 12 windows (72 tokens <= 80: one key tile, plain softmax) as a PERSISTENT kernel: a CTA of five warps (16 query rows each)
// walks over (window, head) items and stages item i + 1 (cp.async, second buffer) while it multiplies item i -- a window is ~1 us
// of work behind several dependent global-memory latencies, so one-item CTAs were latency-bound at a fifth of the HBM roofline.
// rope2 is applied to the staged q, k rows in shared memory when a.rope is set.
// =============================================================================================
template <int HD>
struct Attn1WinSmem {
  static constexpr int WARPS = 5, ROWS = 80;
  static constexpr int RS = HD / 2 + 4;
  static constexpr int BUF_WORDS = 3 * ROWS * RS;               // Q, K, V of one item
  static constexpr int NBUF = (2 * BUF_WORDS * 4 <= 100 * 1024) ? 2 : 1;      // head width 192: one buffer (96 KB), no prefetch
  static constexpr int BYTES = NBUF * BUF_WORDS * 4;
};

template <int HD>
__global__ void __launch_bounds__(160) attn1_win_kernel(const Attn1Args a) {
  using L = Attn1WinSmem<HD>;
  constexpr int RS = L::RS, ROWS = L::ROWS, CH = HD / 8, NB = L::NBUF;
  extern __shared__ __align__(16) uint32_t at1w_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int N = a.wh * a.ww;
  const int nww = a.gw / a.ww;
  const int items = (a.gh / a.wh) * nww * a.heads;
  const int b = blockIdx.y;
  const int d = a.heads * HD;
  pdl_launch_dependents();
  pdl_wait();
  const __half* qkv = reinterpret_cast<const __half*>(a.qkv) + (long long)b * a.qkv_bs;
  __half* out = reinterpret_cast<__half*>(a.out) + (long long)b * a.o_bs;
  auto tok_of = [&](int wi, int wj, int n) {
    const int r = n / a.ww, c = n - r * a.ww;
    int row = wi * a.wh + r + a.sh; if (row >= a.gh) row -= a.gh;
    int col = wj * a.ww + c + a.sw; if (col >= a.gw) col -= a.gw;
    return row * a.gw + col;
  };
  auto stage_item = [&](uint32_t* buf, int item) {              // Q, K, V rows of one (window, head); rows beyond N are zero-filled
    const int h = item % a.heads, win = item / a.heads, wi = win / nww, wj = win - wi * nww;
    for (int idx = threadIdx.x; idx < ROWS * CH; idx += 160) {
      const int r = idx / CH, ch = idx - r * CH;
      uint32_t* sp = buf + r * RS + 4 * ch;
      if (r < N) {
        const __half* src = qkv + (long long)tok_of(wi, wj, r) * a.ld_qkv + h * HD + 8 * ch;
#pragma unroll
        for (int m = 0; m < 3; ++m)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sp + m * ROWS * RS)), "l"(src + m * d) : "memory");
      } else {
#pragma unroll
        for (int m = 0; m < 3; ++m) *reinterpret_cast<uint4*>(sp + m * ROWS * RS) = make_uint4(0, 0, 0, 0);
      }
    }
  };
  const float sc = a.scale * 1.4426950408889634f;              // softmax in base 2
  const int band_row = a.wh - a.sh;
  const int m0 = warp * 16;
  const int lm = lane >> 3, lr = lane & 7;
  int item = blockIdx.x, cur = 0;
  if (item < items) stage_item(at1w_sm, item);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (; item < items; item += gridDim.x) {
    uint32_t* Qs = at1w_sm + cur * L::BUF_WORDS;
    uint32_t* Ks = Qs + ROWS * RS;
    uint32_t* Vs = Ks + ROWS * RS;
    const int nxt = item + gridDim.x;
    if (NB == 2) {
      if (nxt < items) stage_item(at1w_sm + (cur ^ 1) * L::BUF_WORDS, nxt);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int h = item % a.heads, win = item / a.heads, wi = win / nww, wj = win - wi * nww;
    if (a.rope) {
      // rope2 (positional_encodings.py:255-268) on the staged q and k rows: pair (x[j], x[hd / 2 + j]) of window-local token n turns
      // by table[n][j]; two adjacent pairs per step, fp32 arithmetic, rounded back to fp16 exactly as the stand-alone kernel does
      constexpr int HP = HD / 4;
      for (int idx = threadIdx.x; idx < 2 * N * HP; idx += 160) {
        const int rr = idx / HP, jp = idx - rr * HP;
        const int n = rr < N ? rr : rr - N;
        uint32_t* row = (rr < N ? Qs : Ks) + n * RS;
        const float4 cs = __ldg(reinterpret_cast<const float4*>(a.rope + (long long)n * (HD / 2) + 2 * jp));
        const float2 x1 = __half22float2(*reinterpret_cast<__half2*>(&row[jp]));
        const float2 x2 = __half22float2(*reinterpret_cast<__half2*>(&row[HD / 4 + jp]));
        row[jp] = pack_h2_rn(x1.x * cs.x - x2.x * cs.y, x1.y * cs.z - x2.y * cs.w);
        row[HD / 4 + jp] = pack_h2_rn(x2.x * cs.x + x1.x * cs.y, x2.y * cs.z + x1.y * cs.w);
      }
      __syncthreads();
    }
    // ---- S = Q K^T: this warp's 16 query rows against the 80 key rows ----
    float s[ROWS / 8][4];
#pragma unroll
    for (int j = 0; j < ROWS / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      uint32_t af[4];
      ldsm_x4(af, Qs + (m0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * RS + 4 * (lane >> 4) + ks * 8);
#pragma unroll
      for (int j = 0; j < ROWS / 8; j += 2) {
        uint32_t kb[4];
        ldsm_x4(kb, Ks + (j * 8 + (lane & 7) + 8 * (lane >> 4)) * RS + 4 * ((lane >> 3) & 1) + ks * 8);
        mma_f16(s[j], af, kb[0], kb[1]);
        mma_f16(s[j + 1], af, kb[2], kb[3]);
      }
    }
    // ---- scale, mask (the two latitude bands of the last window row of a shifted frame), softmax ----
    const bool masked = a.mask && a.sh > 0 && wi == a.gh / a.wh - 1;
    const int qi0 = m0 + g, qi1 = qi0 + 8;
    const int qb0 = (qi0 / a.ww) < band_row ? 0 : 1, qb1 = (qi1 / a.ww) < band_row ? 0 : 1;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < ROWS / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kn = 8 * j + 2 * t + e;
        bool dead0 = kn >= N, dead1 = dead0;
        if (masked && !dead0) {
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
    if (mx0 == -INFINITY) mx0 = 0.f;
    if (mx1 == -INFINITY) mx1 = 0.f;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < ROWS / 8; ++j) {
      s[j][0] = ex2_approx(s[j][0] - mx0); s[j][1] = ex2_approx(s[j][1] - mx0);
      s[j][2] = ex2_approx(s[j][2] - mx1); s[j][3] = ex2_approx(s[j][3] - mx1);
      l0 += s[j][0] + s[j][1]; l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
    // ---- O = P V, normalised, staged through this warp's own Q rows (read by no other warp), stored ----
    uint32_t pa[ROWS / 16][4];
#pragma unroll
    for (int kb = 0; kb < ROWS / 16; ++kb) {
      pa[kb][0] = pack_h2(s[2 * kb][0], s[2 * kb][1]); pa[kb][1] = pack_h2(s[2 * kb][2], s[2 * kb][3]);
      pa[kb][2] = pack_h2(s[2 * kb + 1][0], s[2 * kb + 1][1]); pa[kb][3] = pack_h2(s[2 * kb + 1][2], s[2 * kb + 1][3]);
    }
    __syncwarp();
#pragma unroll
    for (int c16 = 0; c16 < HD / 16; ++c16) {
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kb = 0; kb < ROWS / 16; ++kb) {
        uint32_t vb[4];
        ldsm_x4_trans(vb, Vs + (16 * kb + (lm & 1) * 8 + lr) * RS + c16 * 8 + (lm >> 1) * 4);
        mma_f16(o0, pa[kb], vb[0], vb[1]);
        mma_f16(o1, pa[kb], vb[2], vb[3]);
      }
      Qs[(m0 + g) * RS + c16 * 8 + t] = pack_h2(o0[0] * i0, o0[1] * i0);
      Qs[(m0 + g + 8) * RS + c16 * 8 + t] = pack_h2(o0[2] * i1, o0[3] * i1);
      Qs[(m0 + g) * RS + c16 * 8 + 4 + t] = pack_h2(o1[0] * i0, o1[1] * i0);
      Qs[(m0 + g + 8) * RS + c16 * 8 + 4 + t] = pack_h2(o1[2] * i1, o1[3] * i1);
    }
    __syncwarp();
    for (int idx = lane; idx < 16 * CH; idx += 32) {
      const int r = idx / CH, ch = idx - r * CH;
      const int qi = m0 + r;
      if (qi < N)
        *reinterpret_cast<uint4*>(out + (long long)tok_of(wi, wj, qi) * a.ld_o + h * HD + 8 * ch) = *reinterpret_cast<const uint4*>(Qs + (m0 + r) * RS + 4 * ch);
    }
    __syncthreads();                                            // the buffer may be refilled (prefetch of the item after next / the next item)
    if (NB == 2) cur ^= 1;
    else {
      if (nxt < items) stage_item(at1w_sm, nxt);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
}

template <int HD>
static void launch_attn1_win_t(const Attn1Args& a, cudaStream_t s) {
  using L = Attn1WinSmem<HD>;
  static bool attr_set = false;
  static int per_sm = 1;
  if (!attr_set) {
    cudaFuncSetAttribute(attn1_win_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attn1_win_kernel<HD>, 160, L::BYTES) != cudaSuccess || per_sm < 1) per_sm = 1;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)(a.gh / a.wh) * (a.gw / a.ww) * a.heads;
  const long long grid = std::min<long long>(items, std::max(1LL, (long long)sms * per_sm / std::max(1, a.batch)));
  launch_kernel(attn1_win_kernel<HD>, dim3((unsigned)grid, a.batch), dim3(160), L::BYTES, s, a);
}

// =============================================================================================
// Long windows (the whole-grid first LG stage: 16 200 tokens): the same online-softmax attention with the operand traffic of a
// flash kernel -- one CTA = one (window, head, 128 query rows), eight warps of 16 rows; keys / values stream through a
// double-buffered cp.async ring of 64-token tiles (the loads of tile i + 1 fly while tile i is multiplied); Q, K fragments come
// from ldmatrix.x4 (one shared-memory instruction per two MMAs), V fragments from ldmatrix.x4.trans.
// =============================================================================================

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

// =============================================================================================
// The whole-grid stage on the 5th-generation tensor cores (tcgen05 / TMEM / TMA): 16 200 queries x 16 200 keys x 6 heads of width 192
// is 1.2 TFLOP per block -- the dominant cost of one forecast.
//
// One CTA = one head x 256 query rows (two 128-row blocks a, b that ping-pong on the tensor pipe), keys / values stream in 64-token
// tiles.  warp 0: TMA producer (Q once; a two-stage ring of K tiles [64 keys][192] and V^T tiles [192][64 keys]); warp 1: issues
// tcgen05.mma.cta_group::1 (S_x = Q_x K^T: M 128, N 64, K 192; O_x += P_x V: M 128, N 192, K 64) and owns TMEM (O_a, O_b: 2 x 192
// columns, S_a, S_b: 2 x 64); warps 2-5 / 6-9: softmax of block a / b, thread = query row (tcgen05.ld gives a thread its row's 64
// scores: row maximum and sum need no shuffles), P written as fp16 into a 128B-swizzled K-major tile the second MMA reads.
// The running maximum is LAZY: P = exp2(s - m_ref) with m_ref raised (and O, l rescaled through tcgen05.ld / tcgen05.st) only when
// a row's maximum exceeds it by more than 2^8 -- after the first tiles that is rare, so O stays in TMEM untouched.
// V is transposed once per launch into vt[head * 192 + c][key] so that both MMA operands are K-major TMA tiles.
// =============================================================================================
constexpr int TC_HD = 192, TC_QB = 128, TC_KT = 64;
constexpr int TC_Q_BYTES = TC_QB * TC_HD * 2;        // 48 KB: three 64-column swizzle atoms of 128 rows x 128 B
constexpr int TC_K_BYTES = TC_KT * TC_HD * 2;        // 24 KB: three atoms of 64 rows x 128 B
constexpr int TC_V_BYTES = TC_HD * TC_KT * 2;        // 24 KB: one atom column, 192 rows x 128 B
constexpr int TC_P_BYTES = TC_QB * TC_KT * 2;        // 16 KB: 128 rows x 128 B
constexpr int TC_OFF_K = 2 * TC_Q_BYTES, TC_OFF_V = TC_OFF_K + 2 * TC_K_BYTES, TC_OFF_P = TC_OFF_V + 2 * TC_V_BYTES;
constexpr int TC_OFF_BAR = TC_OFF_P + 2 * TC_P_BYTES;
constexpr int TC_SMEM = TC_OFF_BAR + 16 * 8 + 16 + 1024;
static_assert(TC_SMEM <= 232448, "shared memory budget");
constexpr int TC_THREADS = 320;
constexpr float TC_LAZY = 8.0f;                      // rescale when a row maximum exceeds the reference by more than 2^8

struct Attn1TcParams {
  int N, heads, d;                                   // tokens (= keys = queries of the one window), heads, heads * 192
  bf16* out; long long ld_o;
  float sc;                                          // scale * log2(e)
};

__global__ void __launch_bounds__(256) vt_transpose_kernel(const bf16* __restrict__ qkv, long long ld_qkv, int N, int d, bf16* __restrict__ vt, long long vt_ld) {
  __shared__ uint16_t tile[64][66];
  const int n0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  pdl_launch_dependents();
  pdl_wait();
  const uint16_t* src = reinterpret_cast<const uint16_t*>(qkv) + 2 * d + c0;
  for (int idx = threadIdx.x; idx < 64 * 32; idx += 256) {       // 64 tokens x 32 channel pairs
    const int r = idx >> 5, cp = idx & 31;
    uint32_t v = 0;
    if (n0 + r < N) v = *reinterpret_cast<const uint32_t*>(src + (long long)(n0 + r) * ld_qkv + 2 * cp);
    tile[r][2 * cp] = (uint16_t)(v & 0xffff); tile[r][2 * cp + 1] = (uint16_t)(v >> 16);
  }
  __syncthreads();
  uint16_t* dst = reinterpret_cast<uint16_t*>(vt);
  for (int idx = threadIdx.x; idx < 64 * 32; idx += 256) {       // 64 channels x 32 token pairs
    const int c = idx >> 5, np = idx & 31;
    const uint32_t v = (uint32_t)tile[2 * np][c] | ((uint32_t)tile[2 * np + 1][c] << 16);
    *reinterpret_cast<uint32_t*>(dst + (long long)(c0 + c) * vt_ld + n0 + 2 * np) = v;      // vt_ld >= round_up(N, 64): in bounds
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attn1_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const Attn1TcParams p) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);
  uint64_t* q_full = bars;              // [1]
  uint64_t* kv_full = bars + 1;         // [2]
  uint64_t* kv_empty = bars + 3;        // [2]
  uint64_t* s_full = bars + 5;          // [2]  (block a, b)
  uint64_t* s_empty = bars + 7;         // [2]
  uint64_t* p_full = bars + 9;          // [2]
  uint64_t* pv_done = bars + 11;        // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 13);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int nqb = (p.N + 2 * TC_QB - 1) / (2 * TC_QB);
  const int h = blockIdx.x / nqb, qblk = blockIdx.x - h * nqb;
  const int q0 = qblk * 2 * TC_QB;
  const int nkt = (p.N + TC_KT - 1) / TC_KT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * TC_Q_BYTES);
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int at = 0; at < 3; ++at)
          tma_load_3d(smem + x * TC_Q_BYTES + at * (TC_QB * 128), &tmQ, q_full, h * TC_HD + at * 64, q0 + x * TC_QB, 0);
    }
    __syncwarp();
    for (int i = 0; i < nkt; ++i) {
      const int st = i & 1;
      mbar_wait(&kv_empty[st], ((i >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&kv_full[st], TC_K_BYTES + TC_V_BYTES);
#pragma unroll
        for (int at = 0; at < 3; ++at)
          tma_load_3d(smem + TC_OFF_K + st * TC_K_BYTES + at * (TC_KT * 128), &tmK, &kv_full[st], p.d + h * TC_HD + at * 64, i * TC_KT, 0);
        tma_load_3d(smem + TC_OFF_V + st * TC_V_BYTES, &tmV, &kv_full[st], i * TC_KT, h * TC_HD, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc_s = make_idesc_16(TC_QB, TC_KT, true);
    const uint32_t idesc_o = make_idesc_16(TC_QB, TC_HD, true);
    const uint32_t sb = smem_u32(smem);
    const uint32_t tO[2] = {tmem_base, tmem_base + TC_HD};
    const uint32_t tS[2] = {tmem_base + 2 * TC_HD, tmem_base + 2 * TC_HD + TC_KT};
    auto issue_s = [&](int x, int i) {               // S_x(i) = Q_x K(i)^T
      const int st = i & 1;
      mbar_wait(&s_empty[x], (i & 1) ^ 1);             // the softmax warps hold S_x(i - 1) in registers
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int at = 0; at < 3; ++at) {
          const uint64_t da = make_smem_desc_sw128(sb + x * TC_Q_BYTES + at * (TC_QB * 128));
          const uint64_t db = make_smem_desc_sw128(sb + TC_OFF_K + st * TC_K_BYTES + at * (TC_KT * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS[x], da + 2 * k, db + 2 * k, idesc_s, (at | k) ? 1u : 0u);
        }
        umma_commit(&s_full[x]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int x, int i) {              // O_x += P_x(i) V(i)
      const int st = i & 1;
      mbar_wait(&p_full[x], i & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = make_smem_desc_sw128(sb + TC_OFF_P + x * TC_P_BYTES);
        const uint64_t db = make_smem_desc_sw128(sb + TC_OFF_V + st * TC_V_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tO[x], da + 2 * k, db + 2 * k, idesc_o, (i | k) ? 1u : 0u);
        umma_commit(&pv_done[x]);
        if (x == 1) umma_commit(&kv_empty[st]);        // the last reader of this K / V stage
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0, 0);
    issue_s(1, 0);
    for (int i = 0; i < nkt; ++i) {
      const bool more = i + 1 < nkt;
      if (more) { mbar_wait(&kv_full[(i + 1) & 1], ((i + 1) >> 1) & 1); tc_fence_after(); }
      issue_pv(0, i);
      if (more) issue_s(0, i + 1);
      issue_pv(1, i);
      if (more) issue_s(1, i + 1);
    }
  } else {
    // ===== softmax warps: block x, TMEM lane quarter q, thread = query row =====
    const int x = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;                    // row of the 128-row block = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tO = tmem_base + x * TC_HD + lane_base;
    const uint32_t tS = tmem_base + 2 * TC_HD + x * TC_KT + lane_base;
    uint8_t* prow = smem + TC_OFF_P + x * TC_P_BYTES + row * 128;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    float m_ref = -INFINITY, l = 0.f;
    for (int i = 0; i < nkt; ++i) {
      mbar_wait(&s_full[x], i & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(tS, r0);
      tmem_ld32(tS + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(&s_empty[x]);       // S_x may be overwritten by the next tile's MMA
      const int live = p.N - i * TC_KT;                // keys of this tile that exist (TMA zero-filled the others)
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a0 = __uint_as_float(r0[j]) * p.sc, a1 = __uint_as_float(r1[j]) * p.sc;
        if (j >= live) a0 = -INFINITY;
        if (32 + j >= live) a1 = -INFINITY;
        r0[j] = __float_as_uint(a0); r1[j] = __float_as_uint(a1);
        mx = fmaxf(mx, fmaxf(a0, a1));
      }
      // O_x and the P_x buffer are stable once PV_x(i - 1) has retired
      mbar_wait(&pv_done[x], (i & 1) ^ 1);
      tc_fence_after();
      const bool bump = mx > m_ref + TC_LAZY;          // first tile: m_ref = -inf -> true
      if (__any_sync(0xffffffffu, bump)) {
        const float m_new = bump ? mx : m_ref;
        const float f = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new);     // 1 for rows that keep their reference
        if (i > 0) {                                    // rescale the accumulator of every row of this quarter (f = 1: unchanged)
#pragma unroll 1
          for (int c = 0; c < TC_HD / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
            tmem_st32(tO + c * 32, o);
          }
          tmem_st_wait();
        }
        l *= f;
        m_ref = m_new;
      }
      float ps = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {                  // 8 keys = one 16-byte chunk of the P row
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = c8 * 8 + 2 * e;
          const float e0 = ex2_approx(__uint_as_float(j < 32 ? r0[j] : r1[j - 32]) - m_ref);
          const float e1 = ex2_approx(__uint_as_float(j + 1 < 32 ? r0[j + 1] : r1[j + 1 - 32]) - m_ref);
          ps += e0 + e1;
          w[e] = pack_h2(e0, e1);
        }
        *reinterpret_cast<uint4*>(prow + ((static_cast<uint32_t>(c8) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      l += ps;
      fence_proxy_async();                              // P (generic proxy) -> the MMA's async-proxy reads
      tc_fence_before();                                // and the rescaled O (tcgen05.st) before the next accumulate
      __syncwarp();
      if (elect_one()) mbar_arrive(&p_full[x]);
    }
    // ---- O / l -> global ----
    mbar_wait(&pv_done[x], (nkt - 1) & 1);
    tc_fence_after();
    const int qi = q0 + x * TC_QB + row;
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    __half* orow = reinterpret_cast<__half*>(p.out) + (long long)qi * p.ld_o + h * TC_HD;
#pragma unroll 1
    for (int c = 0; c < TC_HD / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + c * 32, o);
      tmem_ld_wait();
      if (qi < p.N) {
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          uint4 w;
          w.x = pack_h2(__uint_as_float(o[8 * g4]) * inv, __uint_as_float(o[8 * g4 + 1]) * inv);
          w.y = pack_h2(__uint_as_float(o[8 * g4 + 2]) * inv, __uint_as_float(o[8 * g4 + 3]) * inv);
          w.z = pack_h2(__uint_as_float(o[8 * g4 + 4]) * inv, __uint_as_float(o[8 * g4 + 5]) * inv);
          w.w = pack_h2(__uint_as_float(o[8 * g4 + 6]) * inv, __uint_as_float(o[8 * g4 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + 8 * g4) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool attn1_tc_eligible(const Attn1Args& a) {
  return a.vt != nullptr && a.hd == TC_HD && a.wh == a.gh && a.ww == a.gw && a.sh == 0 && a.sw == 0 && a.mask == 0 && a.batch == 1 &&
         a.gh * a.gw >= 512 && a.vt_ld >= ((a.gh * a.gw + 63) / 64) * 64 && (a.ld_qkv % 8) == 0 && (a.ld_o % 8) == 0 && (a.vt_ld % 8) == 0;
}
static const char* launch_attn1_tc(const Attn1Args& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    attr_set = true;
  }
  const int N = a.gh * a.gw, d = a.heads * a.hd;
  CUtensorMap tmQ, tmK, tmV;
  const char* e;
  if ((e = encode_tma_2d_16(&tmQ, a.qkv, 3LL * d, N, a.ld_qkv, 64, TC_QB))) return e;
  if ((e = encode_tma_2d_16(&tmK, a.qkv, 3LL * d, N, a.ld_qkv, 64, TC_KT))) return e;
  if ((e = encode_tma_2d_16(&tmV, a.vt, N, (long long)d, a.vt_ld, 64, TC_HD))) return e;
  launch_kernel(vt_transpose_kernel, dim3((N + 63) / 64, d / 64), dim3(256), 0, s, a.qkv, a.ld_qkv, N, d, a.vt, a.vt_ld);
  Attn1TcParams p{N, a.heads, d, a.out, a.ld_o, a.scale * 1.4426950408889634f};
  const int nqb = (N + 2 * TC_QB - 1) / (2 * TC_QB);
  launch_kernel(attn1_tc_kernel, dim3(a.heads * nqb), dim3(TC_THREADS), TC_SMEM, s, tmQ, tmK, tmV, p);
  return nullptr;
}

bool attn1_supported(int hd) { return hd == 32 || hd == 64 || hd == 192; }
bool attn1_fuses_rope(int wh, int ww) { return wh * ww <= 80; }
void launch_attn1(const Attn1Args& a, cudaStream_t s) {
  const int N = a.wh * a.ww;
  if (attn1_tc_eligible(a) && !getenv("VV_NO_TC_ATTN")) {
    if (!launch_attn1_tc(a, s)) return;               // (a descriptor that cannot be encoded falls through to the mma.sync kernel)
  }
  if (N <= 80) {                      // one key tile, five query warps: the 6 x 12 windows (72 tokens)
    if (a.hd == 32) launch_attn1_win_t<32>(a, s);
    else if (a.hd == 64) launch_attn1_win_t<64>(a, s);
    else if (a.hd == 192) launch_attn1_win_t<192>(a, s);
  } else {                            // long windows (the whole-grid first LG stage): 128-row query tiles, 64-key tiles, double-buffered
    if (a.hd == 32) launch_attn1_long_t<32>(a, s);
    else if (a.hd == 64) launch_attn1_long_t<64>(a, s);
    else if (a.hd == 192) launch_attn1_long_t<192>(a, s);
  }
}

// =============================================================================================
// Patch embedding, 3 x 2 kernel / stride 2
// =============================================================================================
// One CTA = 64 tokens x D channels of one group.  The patch tile is staged as [k][token] (image rows are read as coalesced float2
// pixel pairs), the group's weights as [k][D]; a thread accumulates 4 tokens x NC channels in registers (one LDS.128 of patch values
// and NC / 2 LDS.64 of weights per 4 NC FMAs), so the kernel runs on the FMA pipe instead of the shared-memory pipe.
constexpr int P32_TOK = 64;
template <int D>
__global__ void __launch_bounds__(256) patch32_kernel(const Patch32Args a) {
  constexpr int CG = 16, NC = D / CG;                 // 16 channel groups of NC channels x 16 token groups of 4 tokens = 256 threads
  static_assert(D % 32 == 0 && NC % 2 == 0, "D");
  extern __shared__ __align__(16) float p32_sm[];
  const int g = blockIdx.y;
  const int L0 = a.h0 * a.w0;
  const int t0 = blockIdx.x * P32_TOK;
  pdl_launch_dependents();
  pdl_wait();
  const int cnt = a.kcnt[g], cb = a.cbase[g];
  const int K = cnt * 6;
  float* patch = p32_sm;                         // [K][P32_TOK]
  float* Ws = patch + a.max_cnt * 6 * P32_TOK;   // [K][D]
  const long long HW = (long long)a.H * a.W;
  for (int idx = threadIdx.x; idx < cnt * 3 * P32_TOK; idx += 256) {
    const int tk = idx % P32_TOK, rk = idx / P32_TOK, ci = rk / 3, kr = rk - ci * 3;
    const int tok = t0 + tk;
    float2 v = make_float2(0.f, 0.f);
    if (tok < L0) {
      const int i = tok / a.w0, j = tok - i * a.w0;
      v = *reinterpret_cast<const float2*>(a.img + (cb + ci) * HW + (long long)(2 * i + kr) * a.W + 2 * j);     // W even: 8-byte aligned
    }
    patch[(rk * 2) * P32_TOK + tk] = v.x;
    patch[(rk * 2 + 1) * P32_TOK + tk] = v.y;
  }
  const float* wsrc = a.Wp + (long long)cb * 6 * D;
  for (int idx = threadIdx.x; idx < K * D / 4; idx += 256) reinterpret_cast<float4*>(Ws)[idx] = __ldg(reinterpret_cast<const float4*>(wsrc) + idx);
  __syncthreads();
  const int cq = threadIdx.x % CG, tq = threadIdx.x / CG;
  float acc[4][NC];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[t][c] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float4 pv = *reinterpret_cast<const float4*>(patch + k * P32_TOK + 4 * tq);
    float w[NC];
#pragma unroll
    for (int c = 0; c < NC; c += 2) {
      const float2 w2 = *reinterpret_cast<const float2*>(Ws + k * D + NC * cq + c);
      w[c] = w2.x; w[c + 1] = w2.y;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      acc[0][c] = fmaf(pv.x, w[c], acc[0][c]); acc[1][c] = fmaf(pv.y, w[c], acc[1][c]);
      acc[2][c] = fmaf(pv.z, w[c], acc[2][c]); acc[3][c] = fmaf(pv.w, w[c], acc[3][c]);
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int tok = t0 + 4 * tq + t;
    if (tok >= L0) continue;
    const long long o = ((long long)g * L0 + tok) * D + NC * cq;
#pragma unroll
    for (int c = 0; c < NC; c += 2) {
      const float2 b2 = __ldg(reinterpret_cast<const float2*>(a.bias + g * D + NC * cq + c));
      const float2 e2 = __ldg(reinterpret_cast<const float2*>(a.ape + o + c));
      *reinterpret_cast<float2*>(a.tok + o + c) = make_float2(acc[t][c] + b2.x + e2.x, acc[t][c + 1] + b2.y + e2.y);
    }
  }
}
template <int D>
static void launch_patch32_t(const Patch32Args& a, cudaStream_t s) {
  const int K = a.max_cnt * 6;
  const size_t smem = (size_t)(K * P32_TOK + K * D) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(patch32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  const int L0 = a.h0 * a.w0;
  launch_kernel(patch32_kernel<D>, dim3((L0 + P32_TOK - 1) / P32_TOK, a.G), dim3(256), smem, s, a);
}
bool patch32_supported(int D) { return D == 32 || D == 64 || D == 96 || D == 128; }
void launch_patch32(const Patch32Args& a, cudaStream_t s) {
  switch (a.D) {
    case 32: launch_patch32_t<32>(a, s); break;
    case 64: launch_patch32_t<64>(a, s); break;
    case 96: launch_patch32_t<96>(a, s); break;
    case 128: launch_patch32_t<128>(a, s); break;
    default: break;
  }
}

// =============================================================================================
// ConvTranspose2d head, 3 x 2 kernel / stride 2: output-stationary (no atomics) -- a CTA owns image rows 2 i and 2 i + 1 over 64
// patch columns of one group and reads patch rows i (kernel rows 0, 1) and i - 1 (kernel row 2, the overlap onto row 2 i).
// A thread owns two patch columns (four pixels) x two output slots x both image rows: 24 FMAs per channel against two LDS.64 of
// token values and six LDS.64 of weights ([slot][channel][kernel row][kernel column] in shared memory).
// =============================================================================================
constexpr int CT_TOK = 64;
__global__ void __launch_bounds__(256) convt32_kernel(const ConvT32Args a) {
  extern __shared__ __align__(16) float ct_sm[];
  const int nblk = (a.max_cnt + 15) / 16;         // blocks of 16 output slots per group
  const int g = blockIdx.z / nblk, sb = blockIdx.z - g * nblk, i = blockIdx.y, j0 = blockIdx.x * CT_TOK;
  const int D = a.D;
  constexpr int XS = CT_TOK + 2;                  // row stride of the transposed token tiles: 8-byte aligned pairs, 2-way store conflicts
  pdl_launch_dependents();
  pdl_wait();
  const int cnt = a.kcnt[g] - 16 * sb, cb = a.cbase[g] + 16 * sb;      // this CTA's slots: cb .. cb + min(cnt, 16)
  if (cnt <= 0) return;
  float* X0 = ct_sm;                              // [D][XS] patch row i      (zeros if i == h0)
  float* X1 = X0 + D * XS;                        // [D][XS] patch row i - 1  (zeros if i == 0)
  float* Ws = X1 + D * XS;                        // [16 slots][D][3][2] (slots >= cnt: zeros)
  const long long L0 = (long long)a.h0 * a.w0;
  for (int idx = threadIdx.x; idx < CT_TOK * D; idx += 256) {
    const int tk = idx / D, c = idx - tk * D;     // consecutive threads -> consecutive channels of a token (coalesced reads)
    const int j = j0 + tk;
    float v0 = 0.f, v1 = 0.f;
    if (j < a.w0) {
      if (i < a.h0) v0 = a.tok[((long long)g * L0 + (long long)i * a.w0 + j) * D + c];
      if (i > 0) v1 = a.tok[((long long)g * L0 + (long long)(i - 1) * a.w0 + j) * D + c];
    }
    X0[c * XS + tk] = v0; X1[c * XS + tk] = v1;
  }
  const float* wsrc = a.Wt + (long long)cb * 6 * D;            // [slot][3][2][D]
  for (int idx = threadIdx.x; idx < 16 * 6 * D; idx += 256) {
    const int slot = idx / (6 * D), rem = idx - slot * 6 * D, kk = rem / D, c = rem - kk * D;
    Ws[(slot * D + c) * 6 + kk] = slot < cnt ? __ldg(wsrc + idx) : 0.f;
  }
  __syncthreads();
  const int tp = threadIdx.x & 31, sg = threadIdx.x >> 5;      // token pair (2 tp, 2 tp + 1), slot pair (2 sg, 2 sg + 1)
  float ev[2][2][2], od[2][2][2];                              // [slot][token][kernel column]: even / odd image row
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int t = 0; t < 2; ++t) ev[s][t][0] = ev[s][t][1] = od[s][t][0] = od[s][t][1] = 0.f;
  if (2 * sg < cnt) {
    for (int c = 0; c < D; ++c) {
      const float2 x0 = *reinterpret_cast<const float2*>(X0 + c * XS + 2 * tp);
      const float2 x1 = *reinterpret_cast<const float2*>(X1 + c * XS + 2 * tp);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float* w = Ws + ((2 * sg + s) * D + c) * 6;
        const float2 k0 = *reinterpret_cast<const float2*>(w), k1 = *reinterpret_cast<const float2*>(w + 2), k2 = *reinterpret_cast<const float2*>(w + 4);
        ev[s][0][0] = fmaf(x0.x, k0.x, fmaf(x1.x, k2.x, ev[s][0][0])); ev[s][0][1] = fmaf(x0.x, k0.y, fmaf(x1.x, k2.y, ev[s][0][1]));
        ev[s][1][0] = fmaf(x0.y, k0.x, fmaf(x1.y, k2.x, ev[s][1][0])); ev[s][1][1] = fmaf(x0.y, k0.y, fmaf(x1.y, k2.y, ev[s][1][1]));
        od[s][0][0] = fmaf(x0.x, k1.x, od[s][0][0]); od[s][0][1] = fmaf(x0.x, k1.y, od[s][0][1]);
        od[s][1][0] = fmaf(x0.y, k1.x, od[s][1][0]); od[s][1][1] = fmaf(x0.y, k1.y, od[s][1][1]);
      }
    }
  }
  const long long HW = (long long)a.H * a.W;
  const int x = 2 * (j0 + 2 * tp);                             // first of this thread's four pixels
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int slot = 2 * sg + s;
    if (slot >= cnt || x >= a.W) continue;
    const float b = a.bias[cb + slot];
    float* base = a.img + a.chan[cb + slot] * HW;
    const bool four = x + 3 < a.W;                             // W even: a thread has four pixels or two
    if (2 * i < a.H) {
      float* p = base + (long long)(2 * i) * a.W + x;
      *reinterpret_cast<float2*>(p) = make_float2(ev[s][0][0] + b, ev[s][0][1] + b);
      if (four) *reinterpret_cast<float2*>(p + 2) = make_float2(ev[s][1][0] + b, ev[s][1][1] + b);
    }
    if (2 * i + 1 < a.H) {
      float* p = base + (long long)(2 * i + 1) * a.W + x;
      *reinterpret_cast<float2*>(p) = make_float2(od[s][0][0] + b, od[s][0][1] + b);
      if (four) *reinterpret_cast<float2*>(p + 2) = make_float2(od[s][1][0] + b, od[s][1][1] + b);
    }
  }
}
void launch_convt32(const ConvT32Args& a, cudaStream_t s) {
  const size_t smem = (size_t)(2 * (CT_TOK + 2) * a.D + 16 * 6 * a.D) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(convt32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  const int nblk = (a.max_cnt + 15) / 16;
  launch_kernel(convt32_kernel, dim3((a.w0 + CT_TOK - 1) / CT_TOK, a.h0 + 1, a.G * nblk), dim3(256), smem, s, a);
}

}  // namespace vv
