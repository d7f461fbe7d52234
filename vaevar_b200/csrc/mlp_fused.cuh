// The MLP half of a tower Swin block in ONE kernel per direction (swinblock.py:13-29 Mlp, :304-307 x + mlp(norm2(x))), and norm1 + qkv
// of the same blocks (swinblock.py:268-269, 139) as a third mode of the same template:
//
//   MLP_FWD    out = x1 + fc2( gelu( fc1( LN2(x1) ) ) )            saves gelu'(u) for the backward pass
//   MLP_BWD    dx1 = LN2^T( (dy W2 . gelu'(u)) W1 ) + dy            (input-VJP; weights are frozen, da_4dvar.py:590-603)
//   MLP_LIN    qkv = LN1(x) Wqkv^T + b                              (LayerNorm prologue + one Linear, 16-bit output)
//
// The towers (d = 96 / 192, six variable groups batched) are K = 96 / 192 GEMMs: as separate launches their cost was the epilogue
// and the HBM round trips of the 4d-wide hidden activation (27 % of the step for 12 % of the flops).
// Here the hidden activation never leaves the SM: a CTA owns a 128-token tile and walks over the 4d hidden columns in chunks of 64,
//
//   GEMM 1 (chunk c)   acc1[c & 1] = A (128 x d)  *  W1[c]^T (64 x d)          tcgen05.mma.cta_group::1, M 128, N 64
//   epilogue 1         h = gelu(acc1 + b1)  (backward: acc1 * gelu'(u))  -> 16-bit, 128B-swizzled K-major tile in shared memory
//   GEMM 2 (chunk c)   acc2 += h (128 x 64) * W2[:, c]^T (d x 64)               M 128, N d
//
// with acc1 double-buffered in TMEM so that GEMM 1 of chunk c + 1 runs under epilogue 1 of chunk c.  Both directions have this
// shape (backward: A = dy, "W1" = W2^T, "W2" = W1^T), so one kernel template serves both; MLP_LIN stops after epilogue 1 (acc1 + bias
// -> 16 bit -> global).
//   warp 0        TMA producer: a ring of weight chunks (W1[c] | W2[:, c]); backward also the dy tile and the saved gelu'(u) chunks --
//                 up to a whole tile of them in flight (they stream from HBM) -- which land in the very buffers epilogue 1 then
//                 overwrites in place with du
//   warp 1        tcgen05.mma issuer, TMEM owner
//   warps 2..17   sixteen epilogue warps (four per TMEM lane quarter, 16 hidden columns each per chunk; thread = token row).
//                 Forward prologue: LayerNorm of the x1 tile (two-pass fp32; 8 or 16 lanes per row, so every lane is busy and a row
//                 reduction is 3 or 4 shuffles shared by the 4 or 2 rows of a warp pass) straight into the swizzled A operand -- the
//                 LayerNorm is neither a launch nor a folded epilogue here, and the operand is the normalised value.  For d <= 96 the
//                 NEXT tile's prologue runs before the current tile's final epilogue, whose first GEMMs then overlap it.
//                 Final epilogue: acc2 is transposed through shared memory (thread = row -> lanes across channels) so that every
//                 global access of the residual add / LayerNorm statistics / LayerNorm backward is coalesced; the backward's
//                 row-wise phase is deferred into the next tile's chunk loop (the epilogue warps idle there: that loop runs at the
//                 pace of the MMA-issue thread).  gelu'(u) and the MLP_LIN output leave through per-warp 32 x 16 slabs and TMA stores.
// Shared memory: weight ring | A tile | hidden chunks | (forward) per-warp store slabs | (backward, d <= 96) transposition buffer;
// otherwise the transposition buffer aliases hidden | slabs (and A for d >= 128), which are idle by then.
// Measurements behind these choices: DESIGN.md section 6 ("Round 2, second half"); tools/mlp_probe.py, tools/mlp_trace.py, tools/ncu_mlp.sh.
#pragma once
#include "gemm_tcgen05.cuh"

namespace vv {

struct MlpArgs {
  int rows, batch;               // token rows per batch entry (a multiple of 128), batch entries (variable groups)
  int f16;                       // forward: operands and the saved gelu' are IEEE fp16 (1) or bf16 (0); backward: format of the saved gelu'
  int n_out;                     // MLP_LIN only: output columns of the one Linear (3D for qkv)
  float eps;
  const float* x1;               // [batch][rows][D] fp32: the residual stream after the attention half (input of norm2)
  // forward
  const float* b1;               // [batch][4D]: fc1 bias with norm2's beta folded in (c_n = b_n + sum_k beta_k W[n,k]); W1 carries gamma
  const float* b2;               // [batch][D]
  float* out_f32;                // [batch][rows][D] block output
  __nv_bfloat16* out16;          // optional 16-bit copy of (out - shift)
  long long ld16, bs16;
  const float* shift;            // optional [batch][rows]: centre of the 16-bit copy (GemmArgs::ln_shift)
  float* stats_out;              // optional [batch][rows] float2 (mean, M2) of the output rows: the next block's folded norm1
  __nv_bfloat16* u_out;          // [batch][rows][4D] gelu'(u), through the store map
  // backward
  const float* gamma;            // [batch][D] norm2 weight
  const float* dres;             // [batch][rows][D] fp32 gradient of the block output (the residual branch)
  float* dx;                     // [batch][rows][D] gradient of x1
  __nv_bfloat16* dx16;           // bf16 copy of dx
  unsigned long long* trace;     // debug: clock64 stamps of CTA 0 (128 slots, see tools/mlp_trace.py); null in production
};

enum MlpMode : int { MLP_FWD = 0, MLP_BWD = 1, MLP_LIN = 2 };   // MLP_LIN: LayerNorm + ONE Linear (norm1 -> qkv of a tower block), forward

constexpr int MLP_THREADS = 576;
constexpr int MLP_EPI_WARPS = 16;
constexpr int MLP_HC = 64;       // hidden columns per chunk

template <int D, int MODE>
struct MlpSmem {
  static constexpr bool BWD = MODE == MLP_BWD, LIN = MODE == MLP_LIN;
  static constexpr int KSD = (D + 63) / 64;                 // 64-column swizzle slabs of a K = D operand
  static constexpr int NCH = LIN ? (3 * D + MLP_HC - 1) / MLP_HC : 4 * D / MLP_HC;
  static constexpr int W1C = MLP_HC * 128 * KSD;            // W1 chunk: 64 rows x KSD slabs of 128 B
  static constexpr int W2C = LIN ? 0 : D * 128;             // W2 chunk: D rows x one slab
  static constexpr int STAGE = W1C + W2C;
  // PIPE: the transposition buffer of the final epilogue does not alias the A tile, so the next tile's A (forward: its LayerNorm
  // prologue, backward: its TMA load) and first GEMMs run under the final epilogue of the current tile.  MLP_LIN has no final epilogue.
  static constexpr bool PIPE = LIN || D <= 96;
  static constexpr int NST = LIN ? 4 : BWD ? (D <= 64 ? 4 : D <= 128 ? 3 : 2) : (D <= 96 ? 4 : D <= 128 ? 3 : 2);
  // hidden buffers: forward 2 (filled by epilogue 1); backward as many as fit -- they are the prefetch depth of the saved gelu'(u),
  // which streams from HBM (37.7 MB per d = 96 launch) with ~1 us of latency per chunk
  static constexpr int NH = LIN ? 1 : BWD ? (D == 96 ? 5 : 4) : 2;
  static constexpr int A_BYTES = KSD * 16384;
  static constexpr int HID = LIN ? 0 : 16384;
  static constexpr int GP = BWD ? 0 : MLP_EPI_WARPS * 2 * 1024;   // per warp: two slabs of 32 rows x 32 B
  static constexpr int STG_STRIDE = BWD ? 2 * D + 16 : 4 * D + 16;     // transposition buffer: bf16 (backward) / fp32 (forward) rows
  static constexpr int STG_BYTES = LIN ? 0 : 128 * STG_STRIDE;
  static constexpr int OFF_A = NST * STAGE;
  static constexpr int OFF_HID = OFF_A + A_BYTES;
  static constexpr int OFF_GP = OFF_HID + NH * HID;
  static constexpr int OFF_STG_OWN = OFF_GP + GP;           // backward + PIPE: a buffer of its own
  static constexpr bool STG_OWN = BWD && PIPE;
  static constexpr int OFF_STG = STG_OWN ? OFF_STG_OWN : PIPE ? OFF_HID : OFF_A;
  static constexpr int OFF_CONST = OFF_STG_OWN + (STG_OWN ? STG_BYTES : 0);   // forward: two sets of b1 (4D floats) | b2 (D floats)
  static constexpr int CONST_SET = 5 * D * 4;
  static constexpr int CONST_BYTES = BWD ? 0 : 2 * CONST_SET;
  static constexpr int OFF_BAR = (OFF_CONST + CONST_BYTES + 15) / 16 * 16;
  static constexpr int NBAR = 2 * NST + 3 * NH + 10;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
  static_assert(STG_OWN || STG_BYTES <= (PIPE ? 0 : A_BYTES) + NH * HID + GP, "transposition buffer must fit into the buffers it aliases");
  static_assert(STAGE % 1024 == 0 && W1C % 1024 == 0, "swizzle atoms need 1024-byte alignment");
  static_assert(TOTAL <= 232448, "shared memory budget");
};

VV_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
VV_DEVINL void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// named barrier of the sixteen epilogue warps
VV_DEVINL void mlp_epi_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <int D, int MODE, bool F16>
__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmU, const MlpArgs p) {
  constexpr bool BWD = MODE == MLP_BWD, LIN = MODE == MLP_LIN;
  using L = MlpSmem<D, MODE>;
  constexpr int KSD = L::KSD, NCH = L::NCH, NST = L::NST, NH = L::NH, HC = MLP_HC;
  constexpr bool PIPE = L::PIPE;
  constexpr int D4 = D / 4;                       // float4 per row
  constexpr int LPR = D == 192 ? 16 : 8;          // row-wise phases: lanes per row,
  constexpr int V = D4 / LPR;                     //   float4 per lane,
  constexpr int RPW = 32 / LPR;                   //   rows per warp pass,
  constexpr int NP = 8 / RPW;                     //   passes over a warp's 8 rows,
  constexpr int PG = BWD ? (NP * V * 2 <= 12 ? NP : NP / 2) : NP;   // passes whose global loads are in flight together (registers)
  static_assert(D4 % LPR == 0, "row split");
  constexpr int FW = D / 4;                       // acc2 columns per epilogue warp (four warps per lane quarter)
  static_assert(D % 32 == 0 && D >= 64 && D <= 192, "token width");
  constexpr bool OPF16 = BWD ? false : F16;       // MMA operand format (gradients are always bf16)

  extern __shared__ uint8_t mlp_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mlp_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* w_full = bars;                        // [NST]
  uint64_t* w_empty = bars + NST;                 // [NST]
  uint64_t* hid_full = bars + 2 * NST;            // [NH]
  uint64_t* hid_empty = hid_full + NH;            // [NH]
  uint64_t* u_full = hid_full + 2 * NH;           // [NH] backward: the saved gelu' chunk has landed in the hidden buffer
  uint64_t* a_full = hid_full + 3 * NH;
  uint64_t* a_empty = a_full + 1;
  uint64_t* acc1_full = a_full + 2;               // [2]
  uint64_t* acc1_empty = a_full + 4;              // [2]
  uint64_t* acc2_full = a_full + 6;
  uint64_t* acc2_empty = a_full + 7;
  uint64_t* fin_done = a_full + 8;                // backward, not PIPE: the transposition buffer (aliases A | hidden) has been read
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NBAR);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int tpb = p.rows >> 7;                    // 128-row tiles per batch entry
  const int total_tiles = tpb * p.batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmU);
    if (!LIN) tma_prefetch_desc(&tmW2);
    if (BWD) tma_prefetch_desc(&tmA);
    for (int i = 0; i < NST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < NH; ++i) { mbar_init(&hid_full[i], MLP_EPI_WARPS); mbar_init(&hid_empty[i], 1); mbar_init(&u_full[i], 1); }
    mbar_init(a_full, BWD ? 1 : MLP_EPI_WARPS);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], MLP_EPI_WARPS); }
    mbar_init(acc2_full, 1); mbar_init(acc2_empty, MLP_EPI_WARPS);
    mbar_init(fin_done, MLP_EPI_WARPS);
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
    uint32_t st = 0, st_n = 0, hb = 0, hn = 0, ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int g = tile / tpb, row0 = (tile - g * tpb) << 7;
      for (int c = 0; c < NCH; ++c) {
        mbar_wait(&w_empty[st], (st_n & 1) ^ 1);
        if (elect_one()) {
          uint8_t* sp = smem + st * L::STAGE;
          mbar_expect_tx(&w_full[st], L::STAGE);
#pragma unroll
          for (int j = 0; j < KSD; ++j) tma_load_3d(sp + j * (HC * 128), &tmW1, &w_full[st], 64 * j, c * HC, g);
          if (!LIN) tma_load_3d(sp + L::W1C, &tmW2, &w_full[st], c * HC, 0, g);
        }
        __syncwarp();
        if (++st == NST) { st = 0; ++st_n; }
        if (BWD) {
          if (c == 0) {
            if (!PIPE && ti > 0) mbar_wait(fin_done, (ti - 1) & 1);   // A | hidden double as the previous tile's transposition buffer
            mbar_wait(a_empty, (ti & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(a_full, L::A_BYTES);
#pragma unroll
              for (int j = 0; j < KSD; ++j) tma_load_3d(smem + L::OFF_A + j * 16384, &tmA, a_full, 64 * j, row0, g);
            }
            __syncwarp();
          }
          mbar_wait(&hid_empty[hb], (hn & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&u_full[hb], L::HID);
            tma_load_3d(smem + L::OFF_HID + hb * L::HID, &tmU, &u_full[hb], c * HC, row0, g);
          }
          __syncwarp();
          if (++hb == NH) { hb = 0; ++hn; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc1 = make_idesc_16(128, HC, OPF16);
    const uint32_t idesc2 = make_idesc_16(128, D, OPF16);
    const uint32_t sb = smem_u32(smem);
    const uint32_t tacc2 = tmem_base + 2 * HC;
    uint32_t gc = 0, ti = 0;                                           // gc: chunks whose GEMM 1 has been issued
    uint32_t s1 = 0, s1n = 0;                                          // ring stage (and its use count) of the next GEMM 1
    uint32_t s2 = 0, h2 = 0, h2n = 0;                                  // ring stage / hidden buffer (use count) of the next GEMM 2
    unsigned long long* trc = (p.trace && blockIdx.x == 0 && lane == 0) ? p.trace + 64 : nullptr;           // MMA warp: slots 64..127
    int tslot = 0;
    auto stamp = [&]() { if (trc && tslot < 64) trc[tslot++] = clock64(); };
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      stamp();
      mbar_wait(a_full, ti & 1);
      stamp();                                                         // A ready
      tc_fence_after();
      for (int c = 0; c <= NCH; ++c) {
        if (c < NCH) {                                               // GEMM 1 of chunk c
          const uint32_t b = gc & 1, n = gc >> 1;
          mbar_wait(&w_full[s1], s1n & 1);
          stamp();                                                     // GEMM 1: weights there
          mbar_wait(&acc1_empty[b], (n & 1) ^ 1);
          stamp();                                                     // GEMM 1: accumulator free
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
              const uint64_t da = make_smem_desc_sw128(sb + L::OFF_A + (k >> 2) * 16384) + 2 * (k & 3);
              const uint64_t db = make_smem_desc_sw128(sb + s1 * L::STAGE + (k >> 2) * (HC * 128)) + 2 * (k & 3);
              umma_bf16(tmem_base + b * HC, da, db, idesc1, k ? 1u : 0u);
            }
            umma_commit(&acc1_full[b]);
            if (LIN) umma_commit(&w_empty[s1]);                      // no GEMM 2: the weight chunk is free once GEMM 1 has read it
            if (c == NCH - 1) umma_commit(a_empty);                  // every GEMM 1 of the tile has read A
          }
          __syncwarp();
          ++gc;
          if (++s1 == NST) { s1 = 0; ++s1n; }
        }
        if (!LIN && c > 0) {                                         // GEMM 2 of chunk c - 1
          if (c == 1) mbar_wait(acc2_empty, (ti & 1) ^ 1);           // the previous tile's acc2 has been drained
          mbar_wait(&hid_full[h2], h2n & 1);
          stamp();                                                     // GEMM 2: hidden chunk written
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = make_smem_desc_sw128(sb + L::OFF_HID + h2 * L::HID);
            const uint64_t db = make_smem_desc_sw128(sb + s2 * L::STAGE + L::W1C);
#pragma unroll
            for (int k = 0; k < HC / 16; ++k) umma_bf16(tacc2, da + 2 * k, db + 2 * k, idesc2, (c > 1 || k) ? 1u : 0u);
            umma_commit(&hid_empty[h2]);
            umma_commit(&w_empty[s2]);
            if (c == NCH) umma_commit(acc2_full);
          }
          __syncwarp();
          if (++s2 == NST) s2 = 0;
          if (++h2 == NH) { h2 = 0; ++h2n; }
        }
      }
    }
  } else {
    // ===== epilogue warps =====
    const int ew = warp - 2;
    const int q = warp & 3;                                          // TMEM lane quarter
    const int cp = ew >> 2;                                          // which 16 of a chunk's 64 columns / which quarter of acc2's columns
    const int row = q * 32 + lane;                                   // thread = token row of the tile
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    uint8_t* stg = smem + L::OFF_STG;                                // transposition buffer of the final epilogue
    uint8_t* gslab = smem + L::OFF_GP + ew * 2048;
    // Row-wise phases (LayerNorm prologue, residual add + statistics, LayerNorm backward): a warp owns rows ew * 8 .. + 8 of the tile
    // and walks them RPW at a time, LPR lanes per row, V float4 per lane -- every lane busy, a row reduction is log2(LPR) shuffles
    // shared by RPW rows, and a load / store instruction covers RPW full 128-byte (LPR = 8) or 256-byte segments.
    const int sub = lane % LPR, rsel = lane / LPR;
    int g_cur = -1;
    float4 gam[V];                                                   // backward: norm2 weight of this lane's channels
#pragma unroll
    for (int v = 0; v < V; ++v) gam[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto row_sum = [&](float x) {
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      return x;
    };

    // forward prologue of tile `tile` (the ti-th of this CTA): biases of its group into constant set ti & 1, LayerNorm of this warp's
    // rows into the swizzled A operand
    auto prologue = [&](int tile, uint32_t ti) {
      const int g = tile / tpb, row0 = (tile - g * tpb) << 7;
      float* cs = reinterpret_cast<float*>(smem + L::OFF_CONST + (PIPE ? (ti & 1) : 0) * L::CONST_SET);
      if (LIN) {
        for (int i = threadIdx.x - 64; i < p.n_out; i += 512) cs[i] = __ldg(p.b1 + (long long)g * p.n_out + i);
      } else {
        for (int i = threadIdx.x - 64; i < 5 * D; i += 512)
          cs[i] = i < 4 * D ? __ldg(p.b1 + (long long)g * 4 * D + i) : __ldg(p.b2 + (long long)g * D + (i - 4 * D));
      }
      const float4* xr = reinterpret_cast<const float4*>(p.x1 + ((long long)g * p.rows + row0 + ew * 8) * D);
      float4 xv[NP][V];
#pragma unroll
      for (int pz = 0; pz < NP; ++pz)
#pragma unroll
        for (int v = 0; v < V; ++v) xv[pz][v] = __ldg(xr + (pz * RPW + rsel) * D4 + sub + LPR * v);
      float rstd[NP];
#pragma unroll
      for (int pz = 0; pz < NP; ++pz) {
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) s += (xv[pz][v].x + xv[pz][v].y) + (xv[pz][v].z + xv[pz][v].w);
        const float mean = row_sum(s) * (1.0f / D);
        float sq = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float4& x = xv[pz][v];
          x.x -= mean; x.y -= mean; x.z -= mean; x.w -= mean;
          sq += fmaf(x.x, x.x, x.y * x.y) + fmaf(x.z, x.z, x.w * x.w);
        }
        rstd[pz] = rsqrtf(row_sum(sq) * (1.0f / D) + p.eps);
      }
      mbar_wait(a_empty, (ti & 1) ^ 1);                              // the previous tile's GEMM 1s have read A
#pragma unroll
      for (int pz = 0; pz < NP; ++pz) {
        const int r = ew * 8 + pz * RPW + rsel;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int k = 4 * (sub + LPR * v), kk = k & 63;
          uint8_t* dst = smem + L::OFF_A + (k >> 6) * 16384 + r * 128 + ((static_cast<uint32_t>(kk >> 3) ^ static_cast<uint32_t>(r & 7)) << 4) + ((kk >> 2) & 1) * 8;
          uint2 w;
          w.x = pack16<OPF16>(xv[pz][v].x * rstd[pz], xv[pz][v].y * rstd[pz]);
          w.y = pack16<OPF16>(xv[pz][v].z * rstd[pz], xv[pz][v].w * rstd[pz]);
          *reinterpret_cast<uint2*>(dst) = w;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) mbar_arrive(a_full);
    };
    // pull the rows this warp will read at the start of a later phase (x1 of the next tile; backward also dres) towards L2
    auto prefetch_rows = [&](int tile) {
      if (tile >= total_tiles) return;
      const int g = tile / tpb, row0 = (tile - g * tpb) << 7;
      const long long off = ((long long)g * p.rows + row0 + ew * 8) * D;
      constexpr int LINES = 8 * D * 4 / 128;                         // 128-byte lines of 8 fp32 rows
      for (int l = lane; l < LINES; l += 32) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x1 + off + l * 32));
        if (BWD) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dres + off + l * 32));
      }
    };

    // backward: LayerNorm backward (norm2) of row r of the tile (global row grow) + the residual branch, from the transposition buffer:
    //   dx = rstd (g - mean(g) - xhat mean(g xhat)) + dres,  g = dh gamma          x, dr: this lane's float4s of x1 / dres
    auto ln_bwd_row = [&](int r, long long grow, float4 (&x4)[V], const float4 (&dr4)[V], const float4 (&gm)[V]) {
      float4 gg[V];
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const uint2 w = *reinterpret_cast<const uint2*>(stg + r * L::STG_STRIDE + 8 * (sub + LPR * v));
        const float2 d0 = unpack_bf16(w.x), d1 = unpack_bf16(w.y);
        gg[v] = make_float4(d0.x * gm[v].x, d0.y * gm[v].y, d1.x * gm[v].z, d1.y * gm[v].w);
        s += (x4[v].x + x4[v].y) + (x4[v].z + x4[v].w);
      }
      const float mean = row_sum(s) * (1.0f / D);
      float sa = 0.f, sb2 = 0.f, sc = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float4& x = x4[v];
        const float4 gq = gg[v];
        x.x -= mean; x.y -= mean; x.z -= mean; x.w -= mean;
        sa += fmaf(x.x, x.x, x.y * x.y) + fmaf(x.z, x.z, x.w * x.w);
        sb2 += (gq.x + gq.y) + (gq.z + gq.w);
        sc += fmaf(gq.x, x.x, gq.y * x.y) + fmaf(gq.z, x.z, gq.w * x.w);
      }
      sa = row_sum(sa); sb2 = row_sum(sb2); sc = row_sum(sc);
      const float rstd = rsqrtf(sa * (1.0f / D) + p.eps);
      const float mb = sb2 * (1.0f / D), mc = sc * (1.0f / D) * rstd * rstd;
      float4* orow = reinterpret_cast<float4*>(p.dx + grow * D);
      __nv_bfloat16* o16 = p.dx16 + grow * D;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 x = x4[v], gq = gg[v], dr = dr4[v];
        float4 o;
        o.x = fmaf(rstd, gq.x - mb - x.x * mc, dr.x); o.y = fmaf(rstd, gq.y - mb - x.y * mc, dr.y);
        o.z = fmaf(rstd, gq.z - mb - x.z * mc, dr.z); o.w = fmaf(rstd, gq.w - mb - x.w * mc, dr.w);
        orow[sub + LPR * v] = o;
        *reinterpret_cast<uint2*>(o16 + 4 * (sub + LPR * v)) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    };
    // backward, PIPE: the row-wise phase of a tile is DEFERRED into the chunk loop of the next one (the epilogue warps idle there:
    // the loop runs at the pace of the MMA issue thread) -- pass pz at the top of chunk pz + 1, its loads one chunk earlier.  The
    // transposition buffer has its own shared memory then; it is rewritten only after every warp has arrived for the next tile's last
    // chunk, i.e. after its deferred passes.
#ifdef VV_MLP_NODEFER
    constexpr bool DEFER = false;
#else
    constexpr bool DEFER = BWD && PIPE && NP + 1 < NCH;
#endif
    bool pending = false;
    long long rbase_prev = 0;
    float4 xd[V], dd[V], gamd[V];                                  // gamd: norm2 weight of the deferred tile's group
    auto deferred_load = [&](int pz) {
      const float4* xq = reinterpret_cast<const float4*>(p.x1 + (rbase_prev + ew * 8 + pz * RPW + rsel) * D);
      const float4* dq = reinterpret_cast<const float4*>(p.dres + (rbase_prev + ew * 8 + pz * RPW + rsel) * D);
#pragma unroll
      for (int v = 0; v < V; ++v) { xd[v] = __ldg(xq + sub + LPR * v); dd[v] = __ldg(dq + sub + LPR * v); }
    };
    auto deferred_step = [&](int c) {                                // called at the top of chunk c of the tile that follows
      if (c >= 1 && c - 1 < NP) {
        const int r = ew * 8 + (c - 1) * RPW + rsel;
        ln_bwd_row(r, rbase_prev + r, xd, dd, gamd);
      }
      if (c < NP) deferred_load(c);
      if (c == NP) pending = false;
    };

    unsigned long long* trc = (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0) ? p.trace : nullptr;     // epilogue warp 0: slots 0..63
    int tslot = 0;
    auto stamp = [&]() { if (trc && tslot < 64) trc[tslot++] = clock64(); };
    if (!BWD && PIPE) {
      if ((int)blockIdx.x < total_tiles) prologue(blockIdx.x, 0);
      mlp_epi_sync();                                                // the first tile's biases are in place for every warp
    }
    uint32_t gc = 0, ti = 0, it = 0, hb = 0, hn = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      stamp();                                                       // tile start
      const int g = tile / tpb, row0 = (tile - g * tpb) << 7;
      const long long rbase = (long long)g * p.rows + row0;          // first token row of the tile in [batch * rows]
      const float* consts = reinterpret_cast<const float*>(smem + L::OFF_CONST + (PIPE ? (ti & 1) : 0) * L::CONST_SET);
      if (!BWD && !PIPE) {
        mlp_epi_sync();                                              // everybody has left the previous tile's buffers
        prologue(tile, ti);
        mlp_epi_sync();                                              // the biases are in place for every warp
      }
      if (BWD && g != g_cur) {
#pragma unroll
        for (int v = 0; v < V; ++v) gam[v] = __ldg(reinterpret_cast<const float4*>(p.gamma + (long long)g * D) + sub + LPR * v);
        g_cur = g;
      }
#ifndef VV_MLP_NOPF
      prefetch_rows(tile + gridDim.x);
      if (BWD) prefetch_rows(tile);
#endif

      // ---- epilogue 1, chunk by chunk ----
#pragma unroll 1
      for (int c = 0; c < NCH; ++c, ++gc, ++it) {
        const uint32_t b = gc & 1, n = gc >> 1;
        if (DEFER && pending) deferred_step(c);
        mbar_wait(&acc1_full[b], n & 1);
        stamp();                                                     // chunk: accumulator seen
        tc_fence_after();
        uint32_t r[16];
        tmem_ld16(tmem_base + lane_base + b * HC + cp * 16, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (elect_one()) mbar_arrive(&acc1_empty[b]);
        uint8_t* hrow = smem + L::OFF_HID + hb * L::HID + row * 128;
        const uint32_t o0 = (static_cast<uint32_t>(2 * cp) ^ sw) << 4, o1 = (static_cast<uint32_t>(2 * cp + 1) ^ sw) << 4;
        if (LIN) {
          // the one Linear: acc + bias -> 16-bit -> this warp's staging slab -> TMA store (columns beyond n_out do not exist)
          if (c * HC + cp * 16 < p.n_out) {
            const float4* bs = reinterpret_cast<const float4*>(consts + c * HC + cp * 16);
            uint32_t w[8];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 bb = bs[i4];
              const float2 a0 = add2(make_float2(__uint_as_float(r[4 * i4]), __uint_as_float(r[4 * i4 + 1])), make_float2(bb.x, bb.y));
              const float2 a1 = add2(make_float2(__uint_as_float(r[4 * i4 + 2]), __uint_as_float(r[4 * i4 + 3])), make_float2(bb.z, bb.w));
              w[2 * i4] = pack16<F16>(a0.x, a0.y); w[2 * i4 + 1] = pack16<F16>(a1.x, a1.y);
            }
            uint8_t* slab = gslab + (it & 1) * 1024;
            if (elect_one()) tma_store_wait_read1();                 // the store issued two chunks ago has read this slab
            __syncwarp();
            *reinterpret_cast<uint4*>(slab + lane * 32) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(slab + lane * 32 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
            fence_proxy_async();
            __syncwarp();
            if (elect_one()) {
              tma_store_3d(&tmU, slab, c * HC + cp * 16, row0 + q * 32, g);
              tma_store_commit();
            }
          } else {
            --it;                                                    // no store issued: keep the slab parity in step with the store groups
          }
        } else if (!BWD) {
          const float4* bs = reinterpret_cast<const float4*>(consts + c * HC + cp * 16);
          float y[16], d[16];
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 bb = bs[i4];
            float2 y2, d2;
            gelu_erf_both2(add2(make_float2(__uint_as_float(r[4 * i4]), __uint_as_float(r[4 * i4 + 1])), make_float2(bb.x, bb.y)), &y2, &d2);
            y[4 * i4] = y2.x; y[4 * i4 + 1] = y2.y; d[4 * i4] = d2.x; d[4 * i4 + 1] = d2.y;
            gelu_erf_both2(add2(make_float2(__uint_as_float(r[4 * i4 + 2]), __uint_as_float(r[4 * i4 + 3])), make_float2(bb.z, bb.w)), &y2, &d2);
            y[4 * i4 + 2] = y2.x; y[4 * i4 + 3] = y2.y; d[4 * i4 + 2] = d2.x; d[4 * i4 + 3] = d2.y;
          }
          // gelu'(u): this warp's 32 x 16 piece -> its own staging slab -> TMA store
          uint8_t* slab = gslab + (it & 1) * 1024;
          if (elect_one()) tma_store_wait_read1();                   // the store issued two chunks ago has read this slab
          __syncwarp();
          *reinterpret_cast<uint4*>(slab + lane * 32) =
              make_uint4(pack16<F16>(d[0], d[1]), pack16<F16>(d[2], d[3]), pack16<F16>(d[4], d[5]), pack16<F16>(d[6], d[7]));
          *reinterpret_cast<uint4*>(slab + lane * 32 + 16) =
              make_uint4(pack16<F16>(d[8], d[9]), pack16<F16>(d[10], d[11]), pack16<F16>(d[12], d[13]), pack16<F16>(d[14], d[15]));
          stamp();                                                   // chunk: math done, gelu' staged
          mbar_wait(&hid_empty[hb], (hn & 1) ^ 1);                    // GEMM 2 of the chunk that used this hidden buffer last has read it
          stamp();                                                   // chunk: hidden buffer free
          *reinterpret_cast<uint4*>(hrow + o0) =
              make_uint4(pack16<F16>(y[0], y[1]), pack16<F16>(y[2], y[3]), pack16<F16>(y[4], y[5]), pack16<F16>(y[6], y[7]));
          *reinterpret_cast<uint4*>(hrow + o1) =
              make_uint4(pack16<F16>(y[8], y[9]), pack16<F16>(y[10], y[11]), pack16<F16>(y[12], y[13]), pack16<F16>(y[14], y[15]));
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) {
            tma_store_3d(&tmU, slab, c * HC + cp * 16, row0 + q * 32, g);
            tma_store_commit();
            mbar_arrive(&hid_full[hb]);
          }
        } else {
          stamp();
          mbar_wait(&u_full[hb], hn & 1);                             // gelu'(u) of this chunk sits where du goes
          stamp();
          const bool uf = p.f16 != 0;
          const uint4 u0 = *reinterpret_cast<const uint4*>(hrow + o0), u1 = *reinterpret_cast<const uint4*>(hrow + o1);
          const uint32_t uw[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 m = mul2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), unpack16(uw[i], uf));
            w[i] = pack_bf16(m.x, m.y);
          }
          *reinterpret_cast<uint4*>(hrow + o0) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(hrow + o1) = make_uint4(w[4], w[5], w[6], w[7]);
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) mbar_arrive(&hid_full[hb]);
        }
        if (++hb == NH) { hb = 0; ++hn; }
      }

      // forward, PIPE: the next tile's LayerNorm prologue comes BEFORE this tile's final epilogue -- its GEMM 1s then run under it
      if (!BWD && PIPE && tile + (int)gridDim.x < total_tiles) prologue(tile + gridDim.x, ti + 1);
      if (LIN) {                                                     // no second GEMM, no final epilogue
        mlp_epi_sync();                                              // the next tile's biases are in place for every warp
        continue;
      }

      // ---- final epilogue: acc2 (thread = row) -> transposition buffer -> row-wise phase with coalesced global traffic ----
      // what the row-wise phase needs from global memory is requested first: PG passes per group (register budget)
      stamp();                                                       // chunks (and the next prologue) done
      const float4* xr = reinterpret_cast<const float4*>(p.x1 + (rbase + ew * 8) * D);
      const float4* rr = reinterpret_cast<const float4*>((BWD ? p.dres : p.x1) + (rbase + ew * 8) * D);
      float4 xv[PG][V], dv[BWD ? PG : 1][V];
      auto load_rows = [&](int pz0) {
#pragma unroll
        for (int pz = 0; pz < PG; ++pz)
#pragma unroll
          for (int v = 0; v < V; ++v) {
            xv[pz][v] = __ldg(xr + ((pz0 + pz) * RPW + rsel) * D4 + sub + LPR * v);
            if (BWD) dv[pz][v] = __ldg(rr + ((pz0 + pz) * RPW + rsel) * D4 + sub + LPR * v);
          }
      };
      if (!DEFER) load_rows(0);
      mbar_wait(acc2_full, ti & 1);
      stamp();                                                       // acc2 complete
      tc_fence_after();
      if (!BWD) {
        if (elect_one()) tma_store_wait_read0();                     // the staging slabs are part of the transposition buffer
        __syncwarp();
        mlp_epi_sync();
      }
      {
        uint32_t a[FW];
#pragma unroll
        for (int i = 0; i < FW / 8; ++i) tmem_ld8(tmem_base + lane_base + 2 * HC + cp * FW + 8 * i, a + 8 * i);
        tmem_ld_wait();
        tc_fence_before();
        uint8_t* srow = stg + row * L::STG_STRIDE;
        if (!BWD) {
          const float4* b2s = reinterpret_cast<const float4*>(consts + 4 * D + cp * FW);
#pragma unroll
          for (int i = 0; i < FW / 4; ++i) {
            const float4 bb = b2s[i];
            *reinterpret_cast<float4*>(srow + (cp * FW + 4 * i) * 4) =
                make_float4(__uint_as_float(a[4 * i]) + bb.x, __uint_as_float(a[4 * i + 1]) + bb.y, __uint_as_float(a[4 * i + 2]) + bb.z,
                            __uint_as_float(a[4 * i + 3]) + bb.w);
          }
        } else {
#pragma unroll
          for (int i = 0; i < FW / 8; ++i)
            *reinterpret_cast<uint4*>(srow + (cp * FW + 8 * i) * 2) =
                make_uint4(pack_bf16(__uint_as_float(a[8 * i]), __uint_as_float(a[8 * i + 1])), pack_bf16(__uint_as_float(a[8 * i + 2]), __uint_as_float(a[8 * i + 3])),
                           pack_bf16(__uint_as_float(a[8 * i + 4]), __uint_as_float(a[8 * i + 5])), pack_bf16(__uint_as_float(a[8 * i + 6]), __uint_as_float(a[8 * i + 7])));
        }
      }
      __syncwarp();
      if (elect_one()) mbar_arrive(acc2_empty);
      mlp_epi_sync();
      stamp();                                                       // transposed
      if (DEFER) {
        pending = true; rbase_prev = rbase;
#pragma unroll
        for (int v = 0; v < V; ++v) gamd[v] = gam[v];
      }
#pragma unroll 1
      for (int pz0 = 0; pz0 < (DEFER ? 0 : NP); pz0 += PG) {
        if (pz0 > 0) load_rows(pz0);
#pragma unroll
        for (int pz = 0; pz < PG; ++pz) {
          const int r = ew * 8 + (pz0 + pz) * RPW + rsel;            // row of the tile this lane works on
          const long long grow = rbase + r;
          if (!BWD) {
            float4* orow = reinterpret_cast<float4*>(p.out_f32 + grow * D);
            float s = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const int j = sub + LPR * v;
              const float4 y = *reinterpret_cast<const float4*>(stg + r * L::STG_STRIDE + 16 * j);
              float4& x = xv[pz][v];
              x = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
              orow[j] = x;
              s += (x.x + x.y) + (x.z + x.w);
            }
            if (p.stats_out) {
              const float mean = row_sum(s) * (1.0f / D);
              float sq = 0.f;
#pragma unroll
              for (int v = 0; v < V; ++v) {
                const float4 x = xv[pz][v];
                const float dx = x.x - mean, dy = x.y - mean, dz = x.z - mean, dw = x.w - mean;
                sq += fmaf(dx, dx, dy * dy) + fmaf(dz, dz, dw * dw);
              }
              sq = row_sum(sq);
              if (sub == 0) reinterpret_cast<float2*>(p.stats_out)[grow] = make_float2(mean, sq);
            }
            if (p.out16) {
              const float sh = p.shift ? __ldg(p.shift + grow) : 0.f;
              __nv_bfloat16* crow = p.out16 + (long long)g * p.bs16 + (long long)(row0 + r) * p.ld16;
#pragma unroll
              for (int v = 0; v < V; ++v) {
                const float4 x = xv[pz][v];
                *reinterpret_cast<uint2*>(crow + 4 * (sub + LPR * v)) = make_uint2(pack16<F16>(x.x - sh, x.y - sh), pack16<F16>(x.z - sh, x.w - sh));
              }
            }
          } else {
            if (BWD) ln_bwd_row(r, grow, xv[pz], dv[BWD ? pz : 0], gam);
          }
        }
      }
      stamp();                                                       // tile done
      if (!BWD) {
        if (PIPE) mlp_epi_sync();                                    // the transposition buffer aliases hidden | staging: everybody has read it
      } else if (!PIPE) {
        __syncwarp();
        if (elect_one()) mbar_arrive(fin_done);
      }
    }
    if (DEFER && pending) {                                          // the last tile's row-wise phase
#pragma unroll 1
      for (int c = 0; c <= NP; ++c) deferred_step(c);
    }
    if (!BWD) {
      if (elect_one()) tma_store_wait_read0();
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vv
