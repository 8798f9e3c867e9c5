// GEMM for the scoring stages: C[M,N] = act(A[M,K] * W[N,K]^T + bias[N] + rowbias[idx(row)][N]) (+ C), fp32 in HBM.
//
// Every dense contraction of BIG-C (models/model_0v10.py:446-458 per-frame MLPs and the k=3 conv as 3 taps,
// :103-117 encoder, :178-225 decoder, :478-507 head) and of the grounding net (models/grd_model_v5.py:331-373)
// is a "rows x shared weight" product, so rows of ALL tracks / queries / videos of a shard are stacked
// into one M and run through this kernel.
//
// Precision modes (SURVEY.md section 7 "discrete decisions under reduced precision"):
//   VSG_GEMM_SIMT   (0)  plain fp32 FFMA tiles -- the on-device comparator used by the tests;
//   VSG_GEMM_TF32   (1)  tcgen05.mma kind::tf32, operands read as fp32 straight from TMA-staged smem;
//   VSG_GEMM_3XTF32 (2)  fp32-faithful: A = Ah + Al split on the fly in smem by 4 transform warps, W = Wh + Wl
//                        pre-split in HBM; D += Al*Wh + Ah*Wl + Ah*Wh (error ~2^-21, like an fp32 GEMM).
//   VSG_GEMM_TF32_BF16X2 (3)  fp32-class at 2/3 of the tensor time of (2): the two CORRECTION products only need ~8 significant
//                        bits (they are 2^-11 of the result), so they run as kind::f16 bf16 MMAs (K=16 per instruction, i.e. half
//                        the issue slots of tf32): D += bf16(Al)*bf16(W) + bf16(A)*bf16(Wl) + tf32(A)*tf32(W).  The split warps
//                        write the two bf16 A tiles (SWIZZLE_32B rows of 16) next to the fp32 A tile; bf16(W), bf16(Wl) are
//                        pre-computed in HBM (vsg_split_bf16).  Per-product error ~2^-18 worst case.
//
//   VSG_GEMM_BF16   (4)  reduced precision, ONE kind::f16 pass: A and W are bf16 in HBM (A written as bf16 by the producing kernel /
//                        epilogue or by vsg_cast_bf16), fp32 accumulate, C as fp32 and / or bf16.  Runs as the single-pass kernel
//                        (MODE 1 geometry: 128-byte stage rows = 64 bf16) with B16 = true; CTA pairs use cta_group::2 MMAs.
//
//   VSG_GEMM_FP16X3 (5)  fp32-class at 3/4 of the tensor slots of (3): BOTH operands are split into fp16 pairs, x = hi + lo with
//                        hi = fp16_rn(x) (11 significant bits) and lo = fp16_rn(x - hi) (11 more), and all three products run as kind::f16
//                        MMAs (1 issue slot per MAC each): D += A_lo * W_hi + A_hi * W_lo + A_hi * W_hi.  fp16 has 5 exponent bits, so range
//                        is handled explicitly: W is pre-scaled by an exact power of two (max |W 2^s| in [2^13, 2^14): its low part stays
//                        normal for |w| >= 2^-16 max |W|), A optionally by `a_scale` (a power of two, default 1); both are undone by `alpha`
//                        in the epilogue.  A's low part is exact to 2^-25 ABSOLUTE for |a a_scale| < 2^-2 (fp16 subnormals) and to 2^-23
//                        relative above, i.e. the product is fp32-class as long as the bulk of |A a_scale| is >= ~1e-2; |A a_scale| must
//                        stay below 65504 (else inf).  The split warps build the two fp16 A tiles IN PLACE over the fp32 tile the TMA
//                        delivered (8 KB -> 4 + 4 KB; reads and writes of a group separated by a named barrier), so a stage holds loaded
//                        bytes only: 16 KB per CTA of a pair -> 12 stages in flight (the fp32-class kernels are bound by the latency of
//                        the load -> split -> MMA -> free loop divided by the stage count, not by the tensor pipe).  W operands come as
//                        pre-swizzled tile images only (vsg_build_weight_image_fp16).
//
// tcgen05 kernel anatomy (persistent, one CTA per SM; 128 x BN output tiles per CTA, BN = 256 where N allows, else 128):
//   warp 0       TMA producer: A tiles by cp.async.bulk.tensor.2d (mbarrier complete_tx), W tiles by tensor loads or -- mode 3 -- by
//                contiguous cp.async.bulk copies of pre-swizzled weight-tile images (vsg_build_weight_image)
//   warp 1       MMA issuer: the whole warp walks the uniform loop, one elected lane issues tcgen05.mma + tcgen05.commit
//   warp 2       TMEM allocator (2 x BN columns: the fp32 accumulator is double buffered, epilogue(i) overlaps mainloop(i+1))
//   warps 4-7    epilogue: tcgen05.ld 32x32b.x32 -> bias / row-bias / accumulate / ReLU / residual -> 128B-swizzled staging ->
//                cp.async.bulk.tensor store (edge slabs: 16-byte stores)
//   warps 8-15   (split modes) two groups of 4 warps that own alternate stages and build the low-order A operands in shared memory
//                (mode 2: A_lo fp32; mode 3: bf16(A_lo), bf16(A); CONV variant: also the depthwise conv of the raw X tile)
// Cluster variants (template CL): 1 = single CTA; 2 = CTA pairs sharing an N tile, W half-tiles TMA-multicast; 3 = CTA-pair MMA
// (tcgen05.mma.cta_group::2, M = 256 per pair): each CTA stages its 128 A rows and half of W -> 32 KB stages, 6 deep.
// Stage depth: 192 KB of stages per CTA (+ 32 KB store staging).  History of how it got here: profiles/README.md.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <string.h>

namespace vsg {

constexpr int BM = 128;                               // tile rows
// BK (fp32 per stage row) is a template parameter: 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B).
// The 3xTF32 kernel with 256-wide tiles uses BK=16 so that 4 pipeline stages fit (2 stages of BK=32 cannot cover the
// TMA + split latency).
// BN (tile columns = MMA N) is a template parameter: 256 where N allows it -- an SS-mode M=128 MMA reads
// (128 + N) * 32 B of shared memory per N/2 cycles, i.e. 128 B/clk at N=128 (all of the smem bandwidth) but 96 B/clk at N=256.

struct GemmEpilogue {
  const float* bias;        // [N] or null
  const float* rowbias;     // [P][ld_rb] or null
  const int32_t* rb_index;  // [M] row -> rowbias row, or null (then row % rb_period)
  int rb_period;
  int ld_rb;
  int relu;
  int accumulate;           // C += result (before activation)
  const float* residual;    // [M][ld_res] added AFTER the activation, or null
  int ld_res;
  float* C;                 // may be null when only the bf16 copy is wanted
  int ldc;
  __nv_bfloat16* C16;       // optional bf16 copy of the stored values ([M][ldc16]); the A operand of a following bf16 GEMM
  int ldc16;
  int M, N, K;
  int store_hi;             // 3xTF32: also write the masked high part back (0 = rely on the MMA ignoring the low 13 bits)
  float* C_lo;              // optional: x - trunc_tf32(x) of every stored value (the B-side low part for a following 3xTF32 GEMM)
  int lo_c0, lo_c1;         // C_lo is written for columns in [lo_c0, lo_c1) only
  int tma_store;            // 1: full 32x32 output slabs leave through shared memory + cp.async.bulk.tensor (mapC is valid)
  int lo_tma;               // 1: C_lo slabs that lie fully inside the column window leave the same way (mapClo is valid)
  const uint8_t* w_img;     // MODE 3: pre-swizzled shared-memory images of the W tiles (vsg_build_weight_image), or null; MODE 5: required
  float alpha;              // MODE 5: the accumulator is multiplied by alpha (= 1 / (weight image scale * a_scale)) first
  float a_scale;            // MODE 5: power of two applied to A before the fp16 split
  const float* dw_w;        // CONV: depthwise weights [K][dw_k], bias dw_b [K], per-row position in / rows remaining of its sequence
  const float* dw_b;
  const int32_t* seq_pos;
  const int32_t* seq_rem;
  int dw_k;
  // batched problems: tile -> (problem p, m block, n block); p -> (outer = p / batch_inner, inner = p % batch_inner).
  // TMA coordinates and the C pointer are offset per problem; M / N are per-problem extents (batch == 1: plain GEMM).
  int dbg;                  // probe flags (vsg_gemm_debug_flags): 1 skip W loads, 2 skip the A split, 4 skip MMAs, 8 skip A loads
  int batch, batch_inner;
  int a_row_outer, a_row_inner, a_col_outer, a_col_inner;
  int b_row_outer, b_row_inner, b_col_outer, b_col_inner;
  long long c_outer, c_inner;
};

struct TileCoord { int m0, n0, a_row, a_col, b_row, b_col; long long c_off; int c_row, c_col; };
__device__ __forceinline__ TileCoord tile_coord(const GemmEpilogue& ep, int tile, int tiles_m, int tiles_n, int BN_, int m_mul = 1, int m_add = 0) {
  TileCoord t;
  const int per = tiles_m * tiles_n;
  const int p = tile / per, in = tile - p * per;
  const int outer = p / ep.batch_inner, inner = p - outer * ep.batch_inner;
  t.m0 = ((in / tiles_n) * m_mul + m_add) * 128;     // cluster of 2: the pair owns M tiles 2i and 2i+1 of one N tile
  t.n0 = (in % tiles_n) * BN_;
  t.a_row = t.m0 + outer * ep.a_row_outer + inner * ep.a_row_inner;
  t.a_col = outer * ep.a_col_outer + inner * ep.a_col_inner;
  t.b_row = t.n0 + outer * ep.b_row_outer + inner * ep.b_row_inner;
  t.b_col = outer * ep.b_col_outer + inner * ep.b_col_inner;
  t.c_off = outer * ep.c_outer + inner * ep.c_inner;
  t.c_row = (int)(t.c_off / ep.ldc);
  t.c_col = (int)(t.c_off - (long long)t.c_row * ep.ldc);
  return t;
}

// ---------------------------------------------------------------------------------------------------
// PAIR: the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 (M = 256) per k step; each CTA stages its 128 rows of A and only
// HALF of the W tile (TILE_B below is the per-CTA part), so a stage shrinks from 48 to 32 KB and 6 stages fit instead of 4.
// CONV: the A operand is dwconv(X) (depthwise conv over the row axis, DepthWiseSeparableConv1d of grd_model_v5.py:36-56) computed by the
// split warps from a raw X tile with a 3-row halo on each side, so the separate dwconv pass (one HBM round trip) disappears.
template <int MODE, int BN_, bool PAIR = false, bool CONV = false> struct Cfg {
  static constexpr int BK = ((MODE == 2 && BN_ == 256) || MODE == 3 || MODE == 5) ? 16 : 32;
  static constexpr int TILE_A = BM * BK * 4;
  static constexpr int TILE_B = (PAIR ? BN_ / 2 : BN_) * BK * 4;
  static constexpr int RAW_ROWS = BM + 8;                                        // 3-row halo each side, rounded to the 8-row swizzle atom
  static constexpr int RAW_TX = RAW_ROWS * BK * 4;                               // bytes the raw-tile TMA delivers
  static constexpr int RAW_BYTES = CONV ? ((RAW_TX + 1023) / 1024) * 1024 : 0;
  // MODE 5: [f16(A_hi) | f16(A_lo)] built in place over the fp32 A tile (CONV: raw X with halo first, then the two fp16 tiles), then
  // [f16(W_hi) | f16(W_lo)]: only loaded bytes (and for CONV the conv output) live in a stage
  static constexpr int AREG5 = CONV ? RAW_BYTES + TILE_A : TILE_A;
  static constexpr int OFF_RAW = MODE == 5 ? 0 : (MODE >= 2 ? 2 : 1) * (TILE_A + TILE_B);
  static constexpr int STAGE_BYTES = MODE == 5 ? AREG5 + TILE_B : OFF_RAW + RAW_BYTES;   // [A | B_hi] (+ [A_lo | B_lo]; MODE 3: four bf16 tiles) (+ raw X)
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;                      // 6/4 (tf32), 3/4 (3xTF32), 6/4 (tf32+2xbf16), 12/8 (fp16x3)
  static constexpr int DW_BYTES = CONV ? 8 * 1024 : 0;                           // depthwise weights [K][k] + bias [K] (K <= 224 channels, k <= 7)
  static constexpr int SPLIT_GROUPS = 2;                                         // groups of 4 split warps that alternate stages
  static constexpr int THREADS = MODE >= 2 ? 256 + 128 * SPLIT_GROUPS : 256;
  // MODE 2: [A | B_hi | A_lo | B_lo] fp32.   MODE 3: [A f32 | W f32 | bf16(A_lo) | bf16(A) | bf16(W) | bf16(W_lo)]
  static constexpr int OFF_BH = MODE == 5 ? AREG5 : TILE_A;
  static constexpr int OFF_AL = MODE == 5 ? AREG5 - TILE_A / 2 : TILE_A + TILE_B;                    // MODE 5: f16(A_lo) after f16(A_hi)
  static constexpr int OFF_A16 = MODE == 5 ? AREG5 - TILE_A : OFF_AL + TILE_A / 2, OFF_B16 = OFF_AL + TILE_A;
  static constexpr int OFF_BL = MODE == 5 ? OFF_BH + TILE_B / 2 : (MODE == 3 ? OFF_B16 + TILE_B / 2 : 2 * TILE_A + TILE_B);
  static constexpr int W_TX = MODE == 5 ? TILE_B : (MODE >= 2 ? 2 : 1) * TILE_B;   // W bytes a stage receives
  static constexpr int BAR_BYTES = 512;                                          // mbarriers (3 per stage + 4) + the TMEM slot
  static constexpr int TMEM_COLS = 2 * BN_;                                      // double-buffered fp32 accumulator
  static constexpr int STAGING_BYTES = 4 * 2 * 32 * 128;                         // TMA-store staging of the epilogue warps
};

// CL = 2: CTA pairs (cluster of 2) share one N tile; each CTA loads half of the W tile and multicasts it to both, which
// cuts the L2 -> SM operand traffic per SM from A + W to A + W/2 (W is 4/5 of it in the split modes).  MMA / TMEM stay per CTA.
// CL = 3: CTA-pair MMA (cta_group::2, see Cfg<.., PAIR>): rank 0 issues M = 256 instructions over both CTAs' operands, each CTA
// loads only its half of W (no multicast), the peer's split / epilogue warps signal the leader's barriers through DSMEM arrives.
// B16 (MODE 1 only): operands are bf16 (64 elements per 128-byte stage row), one kind::f16 MMA per 16 columns.
template <int MODE, int BN_, bool PROBE, int CL, bool CONV = false, bool B16 = false>
__global__ void __launch_bounds__(Cfg<MODE, BN_>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
               const __grid_constant__ CUtensorMap mapBl, const __grid_constant__ CUtensorMap mapB16,
               const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapClo, const GemmEpilogue ep) {
  constexpr bool PAIR = (CL == 3);
  using CF = Cfg<MODE, BN_, PAIR, CONV>;
  constexpr int STAGES = CF::STAGES;
  constexpr int STAGE_BYTES = CF::STAGE_BYTES;
  constexpr int BN = BN_;
  constexpr int BK = CF::BK;
  constexpr int BKE = B16 ? 2 * CF::BK : CF::BK;          // elements of K per stage (bf16: 64 per 128-byte row)
  constexpr int TILE_A = CF::TILE_A;
  constexpr int ACC_COLS = BN_;
  constexpr int TMEM_COLS = CF::TMEM_COLS;
  constexpr uint32_t IDESC_TF32 = IDesc<BN_>::tf32;
  // stage layout: [A | B_hi] (MODE 1) or [A(hi) | B_hi | A_lo | B_lo] (MODE 2)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + STAGES * STAGE_BYTES;                     // 4 epilogue warps x 2 buffers x (32 rows x 128 B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + CF::STAGING_BYTES);
  uint64_t* full = bars;                   // TMA landed
  uint64_t* empty = bars + STAGES;         // MMAs that read the stage retired
  uint64_t* ready = bars + 2 * STAGES;     // (MODE 2) split done
  uint64_t* acc_full = bars + 3 * STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sdw = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + CF::BAR_BYTES);      // CONV: [K][dw_k] weights, then [K] bias
  if (CONV) {
    for (int i = threadIdx.x; i < ep.K * ep.dw_k; i += blockDim.x) sdw[i] = ep.dw_w[i];
    for (int i = threadIdx.x; i < ep.K; i += blockDim.x) sdw[ep.K * ep.dw_k + i] = ep.dw_b[i];
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CL >= 2 ? (int)cluster_ctarank() : 0;
  const int tiles_m_real = (ep.M + BM - 1) / BM;
  const int tiles_m = CL >= 2 ? (tiles_m_real + 1) / 2 : tiles_m_real;      // clusters: rows of tile PAIRS (an odd tail gets a dummy tile)
  const int tiles_n = (ep.N + BN - 1) / BN;
  const int n_tiles = tiles_m * tiles_n * ep.batch;
  const int kblocks = (ep.K + BKE - 1) / BKE;
  const int tile0 = CL >= 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile_step = CL >= 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapBh);
    if (MODE >= 2) tma_prefetch_desc(&mapBl);
    if (MODE == 3) tma_prefetch_desc(&mapB16);
    if (ep.tma_store) tma_prefetch_desc(&mapC);
    if (ep.lo_tma) tma_prefetch_desc(&mapClo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CL == 2 ? 2 : 1);   // CL == 2: a stage is refilled by BOTH CTAs' multicasts, so both MMAs must have retired it
      // split modes: every thread of the split group that owns the stage (PAIR: + one DSMEM arrive per peer warp);
      // single-pass PAIR: one relay arrive per CTA of the pair (warp 3) once that CTA's TMA data has landed
      mbar_init(&ready[s], MODE >= 2 ? (PAIR ? 132 : 128) : 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], PAIR ? 132 : 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) { tmem_alloc2(tmem_slot, TMEM_COLS); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CL >= 2) cluster_sync_all();        // the peer's barriers are initialised before anything is multicast to / arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_step) {
        const TileCoord tc = tile_coord(ep, tile, tiles_m, tiles_n, BN, CL >= 2 ? 2 : 1, rank);
        for (int kb = 0; kb < kblocks; ++kb) {
          if (PROBE && (ep.dbg & 32)) mbar_spin(&empty[stage], phase ^ 1); else mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          if (PROBE && MODE != 5 && (ep.dbg & 9)) {   // timing probes only (results are garbage)
            const bool la = !(ep.dbg & 8), lw = !(ep.dbg & 1);
            if (!la && !lw) { mbar_arrive(&full[stage]); }
            else {
              mbar_expect_tx(&full[stage], (la ? TILE_A : 0) + (lw ? (MODE >= 2 ? 2 : 1) * CF::TILE_B : 0));
              if (la) tma_load_2d(smem_u32(st), &mapA, &full[stage], kb * BKE + tc.a_col, tc.a_row);
              if (lw) {
                tma_load_2d(smem_u32(st + CF::OFF_BH), &mapBh, &full[stage], kb * BKE + tc.b_col, tc.b_row);
                if (MODE >= 2) tma_load_2d(smem_u32(st + CF::OFF_BL), &mapBl, &full[stage], kb * BKE + tc.b_col, tc.b_row);
                if (MODE == 3) tma_load_2d(smem_u32(st + CF::OFF_B16), &mapB16, &full[stage], kb * BKE + tc.b_col, tc.b_row);
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full[stage], (CONV ? CF::RAW_TX : TILE_A) + CF::W_TX);
          if (CONV) tma_load_2d(smem_u32(st + CF::OFF_RAW), &mapA, &full[stage], kb * BKE + tc.a_col, tc.a_row - 3);   // rows m0-3 .. m0+132 (OOB = 0)
          else tma_load_2d(smem_u32(st), &mapA, &full[stage], kb * BKE + tc.a_col, tc.a_row);
          if (MODE == 5) {
            // two fp16 sub-images [W_hi | W_lo] of BN rows x 32 B per (N tile, k block), adjacent in the image and in the stage
            constexpr int SUB = BN * 32;                   // one sub-image of the whole N tile
            const uint8_t* img = ep.w_img + ((size_t)(tc.n0 / BN) * kblocks + kb) * (size_t)(2 * SUB);
            if (PAIR) {                                    // this CTA's 128 rows of each sub-image
#pragma unroll
              for (int j = 0; j < 2; ++j)
                bulk_load(smem_u32(st + CF::OFF_BH + j * (SUB / 2)), img + j * SUB + rank * (SUB / 2), SUB / 2, &full[stage]);
            } else if (CL == 2) {                          // half of each sub-image, written into both CTAs of the pair
#pragma unroll
              for (int j = 0; j < 2; ++j)
                bulk_load_mc(smem_u32(st + CF::OFF_BH + j * SUB + rank * (SUB / 2)), img + j * SUB + rank * (SUB / 2), SUB / 2, &full[stage], 3);
            } else {
              bulk_load(smem_u32(st + CF::OFF_BH), img, 2 * SUB, &full[stage]);
            }
          } else if (PAIR) {
            // this CTA's half of the W rows only; the pair's MMA reads the other half from the peer's shared memory
            constexpr int TB = CF::TILE_B;                 // per-CTA W f32 bytes
            if (MODE == 3 && ep.w_img != nullptr) {
              const uint8_t* img = ep.w_img + ((size_t)(tc.n0 / BN) * kblocks + kb) * (size_t)(4 * TB);
              bulk_load(smem_u32(st + CF::OFF_BH), img + rank * TB, TB, &full[stage]);
              bulk_load(smem_u32(st + CF::OFF_B16), img + 2 * TB + rank * (TB / 2), TB / 2, &full[stage]);
              bulk_load(smem_u32(st + CF::OFF_BL), img + 3 * TB + rank * (TB / 2), TB / 2, &full[stage]);
            } else {
              const int brow = tc.b_row + rank * (BN / 2);
              tma_load_2d(smem_u32(st + CF::OFF_BH), &mapBh, &full[stage], kb * BKE + tc.b_col, brow);
              if (MODE >= 2) tma_load_2d(smem_u32(st + CF::OFF_BL), &mapBl, &full[stage], kb * BKE + tc.b_col, brow);
              if (MODE == 3) tma_load_2d(smem_u32(st + CF::OFF_B16), &mapB16, &full[stage], kb * BKE + tc.b_col, brow);
            }
          } else if (MODE == 3 && ep.w_img != nullptr) {
            // W tiles as contiguous pre-swizzled images [W f32 | bf16(W) | bf16(W_lo)] per (N tile, k block)
            constexpr int TB = CF::TILE_B;
            const uint8_t* img = ep.w_img + ((size_t)(tc.n0 / BN) * kblocks + kb) * (size_t)(2 * TB);
            if (CL == 2) {
              bulk_load_mc(smem_u32(st + CF::OFF_BH + rank * (TB / 2)), img + rank * (TB / 2), TB / 2, &full[stage], 3);
              bulk_load_mc(smem_u32(st + CF::OFF_B16 + rank * (TB / 4)), img + TB + rank * (TB / 4), TB / 4, &full[stage], 3);
              bulk_load_mc(smem_u32(st + CF::OFF_BL + rank * (TB / 4)), img + TB + TB / 2 + rank * (TB / 4), TB / 4, &full[stage], 3);
            } else {
              bulk_load(smem_u32(st + CF::OFF_BH), img, TB, &full[stage]);
              bulk_load(smem_u32(st + CF::OFF_B16), img + TB, TB, &full[stage]);      // bf16(W) and bf16(W_lo) are adjacent in the stage
            }
          } else if (CL == 2) {
            // this CTA's half of the W rows (maps have BN/2-row boxes), written into both CTAs of the pair
            constexpr int HB = CF::TILE_B / 2;
            const int brow = tc.b_row + rank * (BN / 2);
            tma_load_2d_mc(smem_u32(st + CF::OFF_BH + rank * HB), &mapBh, &full[stage], kb * BKE + tc.b_col, brow, 3);
            if (MODE == 2) tma_load_2d_mc(smem_u32(st + CF::OFF_BL + rank * HB), &mapBl, &full[stage], kb * BKE + tc.b_col, brow, 3);
            if (MODE == 3) {
              tma_load_2d_mc(smem_u32(st + CF::OFF_BL + rank * (HB / 2)), &mapBl, &full[stage], kb * BKE + tc.b_col, brow, 3);
              tma_load_2d_mc(smem_u32(st + CF::OFF_B16 + rank * (HB / 2)), &mapB16, &full[stage], kb * BKE + tc.b_col, brow, 3);
            }
          } else {
            tma_load_2d(smem_u32(st + CF::OFF_BH), &mapBh, &full[stage], kb * BKE + tc.b_col, tc.b_row);
            if (MODE >= 2) tma_load_2d(smem_u32(st + CF::OFF_BL), &mapBl, &full[stage], kb * BKE + tc.b_col, tc.b_row);
            if (MODE == 3) tma_load_2d(smem_u32(st + CF::OFF_B16), &mapB16, &full[stage], kb * BKE + tc.b_col, tc.b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The WHOLE warp walks the (warp-uniform) loop and one elected lane issues: a lane-divergent `if (lane == 0)` loop makes the
    // compiler rebuild every descriptor through ELECT / R2UR sequences (~160 SASS instructions per k-block, measured to take about
    // as long as the MMAs themselves).  Descriptors are one 32-bit add on a per-kernel base.
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t base16 = (smem_u32(smem) & 0x3FFFF) >> 4;                                 // stage 0 start address >> 4
    constexpr uint32_t DESC_HI = (uint32_t)((8 * BK * 4) >> 4) | (1u << 14) | ((BK == 32 ? 2u : 4u) << 29);        // SBO | version | swizzle
    constexpr uint32_t DESC_HI_B16 = (uint32_t)(256 >> 4) | (1u << 14) | (6u << 29);                                // bf16 tiles, SWIZZLE_32B
    constexpr uint32_t LBO = 1u << 16;
    for (int tile = tile0; tile < n_tiles && (!PAIR || rank == 0); tile += tile_step) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      for (int kb = 0; kb < kblocks; ++kb) {
        // (PAIR: the peer's arrives are release.cluster; like CUTLASS's 2-SM pipelines the wait itself is the default try_wait --
        //  an acquire.cluster try_wait per k block was measured to cost ~1 us each)
        if (PROBE && (ep.dbg & 32)) mbar_spin(MODE >= 2 ? &ready[stage] : &full[stage], phase);
        else mbar_wait((MODE >= 2 || PAIR) ? &ready[stage] : &full[stage], phase);
        tc_fence_after();
        if (PAIR) {
          if (elect_one()) {
            constexpr uint32_t ID_TF32 = IDesc<BN_, 256>::tf32, ID_BF16 = IDesc<BN_, 256>::bf16;
            const uint32_t st16 = base16 + (uint32_t)stage * (STAGE_BYTES >> 4);
            auto desc = [&](uint32_t off_bytes, uint32_t hi) -> uint64_t {
              return ((uint64_t)hi << 32) | (uint64_t)((st16 + (off_bytes >> 4)) | LBO);
            };
            const uint64_t a_hi = desc(0, DESC_HI), b_hi = desc(CF::OFF_BH, DESC_HI);
            if (MODE < 2) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                if (B16) umma2_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, ID_BF16, (kb | k) ? 1u : 0u);
                else umma2_tf32(d_tmem, a_hi + 2 * k, b_hi + 2 * k, ID_TF32, (kb | k) ? 1u : 0u);
              }
            } else if (MODE == 5) {
              // corrections first: A_lo W_hi, A_hi W_lo; then A_hi W_hi -- three K = 16 kind::f16 instructions on fp16 tiles
              constexpr uint32_t ID_F16 = IDesc<BN_, 256>::f16;
              umma2_bf16(d_tmem, desc(CF::OFF_AL, DESC_HI_B16), desc(CF::OFF_BH, DESC_HI_B16), ID_F16, kb ? 1u : 0u);
              umma2_bf16(d_tmem, desc(CF::OFF_A16, DESC_HI_B16), desc(CF::OFF_BL, DESC_HI_B16), ID_F16, 1u);
              umma2_bf16(d_tmem, desc(CF::OFF_A16, DESC_HI_B16), desc(CF::OFF_BH, DESC_HI_B16), ID_F16, 1u);
            } else if (MODE == 3) {
              umma2_bf16(d_tmem, desc(CF::OFF_AL, DESC_HI_B16), desc(CF::OFF_B16, DESC_HI_B16), ID_BF16, kb ? 1u : 0u);
              umma2_bf16(d_tmem, desc(CF::OFF_A16, DESC_HI_B16), desc(CF::OFF_BL, DESC_HI_B16), ID_BF16, 1u);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) umma2_tf32(d_tmem, a_hi + 2 * k, b_hi + 2 * k, ID_TF32, 1u);
            } else {
              const uint64_t a_lo = desc(CF::OFF_AL, DESC_HI), b_lo = desc(CF::OFF_BL, DESC_HI);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) umma2_tf32(d_tmem, a_lo + 2 * k, b_hi + 2 * k, ID_TF32, (kb | k) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) umma2_tf32(d_tmem, a_hi + 2 * k, b_lo + 2 * k, ID_TF32, 1u);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) umma2_tf32(d_tmem, a_hi + 2 * k, b_hi + 2 * k, ID_TF32, 1u);
            }
            umma2_commit_mc(&empty[stage], 3);                       // frees the stage in BOTH CTAs
            if (kb == kblocks - 1) umma2_commit_mc(&acc_full[acc], 3);  // both epilogues read their own 128 accumulator rows
          }
        } else if (elect_one()) {
          const uint32_t st16 = base16 + (uint32_t)stage * (STAGE_BYTES >> 4);
          auto desc = [&](uint32_t off_bytes, uint32_t hi) -> uint64_t {
            return ((uint64_t)hi << 32) | (uint64_t)((st16 + (off_bytes >> 4)) | LBO);
          };
          const uint64_t a_hi = desc(0, DESC_HI), b_hi = desc(CF::OFF_BH, DESC_HI);
          if (PROBE && (ep.dbg & 4)) {
          } else if (MODE == 5) {
            umma_bf16(d_tmem, desc(CF::OFF_AL, DESC_HI_B16), desc(CF::OFF_BH, DESC_HI_B16), IDesc<BN_>::f16, kb ? 1u : 0u);
            umma_bf16(d_tmem, desc(CF::OFF_A16, DESC_HI_B16), desc(CF::OFF_BL, DESC_HI_B16), IDesc<BN_>::f16, 1u);
            umma_bf16(d_tmem, desc(CF::OFF_A16, DESC_HI_B16), desc(CF::OFF_BH, DESC_HI_B16), IDesc<BN_>::f16, 1u);
          } else if (MODE == 3) {
            // corrections first (bf16, one K=16 instruction each), then the tf32 main product
            umma_bf16(d_tmem, desc(CF::OFF_AL, DESC_HI_B16), desc(CF::OFF_B16, DESC_HI_B16), IDesc<BN_>::bf16, kb ? 1u : 0u);
            umma_bf16(d_tmem, desc(CF::OFF_A16, DESC_HI_B16), desc(CF::OFF_BL, DESC_HI_B16), IDesc<BN_>::bf16, 1u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_tf32(d_tmem, a_hi + 2 * k, b_hi + 2 * k, IDESC_TF32, 1u);
          } else if (MODE == 2) {
            const uint64_t a_lo = desc(CF::OFF_AL, DESC_HI), b_lo = desc(CF::OFF_BL, DESC_HI);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_tf32(d_tmem, a_lo + 2 * k, b_hi + 2 * k, IDESC_TF32, (kb | k) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_tf32(d_tmem, a_hi + 2 * k, b_lo + 2 * k, IDESC_TF32, 1u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_tf32(d_tmem, a_hi + 2 * k, b_hi + 2 * k, IDESC_TF32, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (B16) umma_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, IDesc<BN_>::bf16, (kb | k) ? 1u : 0u);
              else umma_tf32(d_tmem, a_hi + 2 * k, b_hi + 2 * k, IDESC_TF32, (kb | k) ? 1u : 0u);
            }
          }
          if (PROBE && (ep.dbg & 16)) mbar_arrive(&empty[stage]);
          else if (CL == 2) umma_commit_mc(&empty[stage], 3);
          else umma_commit(&empty[stage]);
          if (kb == kblocks - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp == 3 && PAIR && MODE < 2) {
    // ================= single-pass CTA pairs: stage relay =================
    // the leader's cta_group::2 MMA reads BOTH CTAs' stage; each CTA's relay lane waits for its own TMA data and arrives on the
    // leader's `ready` barrier (count 2) -- the peer through a DSMEM arrive, like the split warps of the fp32-class modes
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_step) {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          if (rank != 0) mbar_arrive_cluster(&ready[stage], 0); else mbar_arrive(&ready[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= epilogue =================
    // Per warp: 32 accumulator rows.  Each 32-column chunk is read from TMEM (one row per lane), gets bias / row-bias /
    // accumulate / ReLU / residual, and leaves either (fast path, full 32x32 slab) through a 128B-swizzled staging buffer and a
    // TMA store -- coalesced, asynchronous, two buffers in flight -- or (edges, unaligned C) through per-row 16-byte stores.
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int row_in_tile = quad * 32 + lane;
    uint8_t* stg = staging + quad * (2 * 4096);
    int acc = 0, sbuf = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = ((ep.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.C) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0);
    const bool bias_vec = (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0;
    const bool res_vec = ((ep.ld_res & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0);
    const bool c16_vec = ((ep.ldc16 & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.C16) & 7) == 0);
    const bool c16_tile = ep.C16 && ((ep.ldc16 & 7) == 0) && ((reinterpret_cast<uintptr_t>(ep.C16) & 15) == 0);
    const bool rb_vec = ((ep.ld_rb & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.rowbias) & 15) == 0);
    for (int tile = tile0; tile < n_tiles; tile += tile_step) {
      const TileCoord tc = tile_coord(ep, tile, tiles_m, tiles_n, BN, CL >= 2 ? 2 : 1, rank);
      const int m0 = tc.m0, n0 = tc.n0;
      const int row = m0 + row_in_tile;
      const bool row_ok = row < ep.M;
      // While this tile's mainloop is still running: pull the lane's residual / accumulate row (one 128-byte line per 32-column chunk)
      // towards L2.  The operands come straight from HBM (written by the previous launch); unprefetched, every chunk of the epilogue
      // waited a full DRAM round trip for them (the residual adds were the top stall of the conv-fused grounding GEMM).
      if (row_ok && (ep.residual || ep.accumulate)) {
        const float* pre = ep.residual ? ep.residual + (size_t)row * ep.ld_res : ep.C + tc.c_off + (size_t)row * ep.ldc;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c)
          if (n0 + c * 32 < ep.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pre + n0 + c * 32));
      }
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * ACC_COLS + ((uint32_t)(quad * 32) << 16);
      const bool slab_rows_ok = m0 + quad * 32 + 32 <= ep.M;                 // warp-uniform
      const float* rb = nullptr;
      if (row_ok && ep.rowbias) {
        const int ri = ep.rb_index ? ep.rb_index[row] : (row % ep.rb_period);
        rb = ep.rowbias + (size_t)ri * ep.ld_rb;
      }
      float* crow = ep.C ? ep.C + tc.c_off + (size_t)(row_ok ? row : 0) * ep.ldc : nullptr;
      __nv_bfloat16* crow16 = (ep.C16 && row_ok) ? ep.C16 + (size_t)row * ep.ldc16 : nullptr;   // plain problems only (no batch offset)
      float* crow_lo = ep.C_lo ? ep.C_lo + tc.c_off + (size_t)(row_ok ? row : 0) * ep.ldc : nullptr;
      const float* res = (row_ok && ep.residual) ? ep.residual + (size_t)row * ep.ld_res : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= ep.N) break;                                              // warp-uniform
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        const bool fast = ep.C && ep.tma_store && vec_ok && slab_rows_ok && col0 + 32 <= ep.N;   // warp-uniform
        // LEAN path (nearly every launch: bias / ReLU / residual only): the generic loop below re-derives four optional operand
        // pointers per 16 bytes and waits for each bias load right before its add -- ~700 SASS instructions and eight exposed load
        // latencies per 32-column chunk, which made the epilogue (not the MMAs, not HBM) the limit of the N = K = 128 grounding GEMMs
        // and of the K = 512 decoder GEMMs (round-2 ncu source view).  Here the chunk's bias comes from ONE coalesced load (lane l holds
        // column col0 + l, broadcast by shuffles) and the residual row from eight independent loads, all issued before the TMEM wait.
        const bool lean = fast && !rb && !ep.accumulate && !crow_lo && !crow16 && (!res || res_vec) && bias_vec;   // warp-uniform
        if (lean) {
          const float bl = ep.bias ? ep.bias[col0 + lane] : 0.f;
          float4 rv[8];
          if (res) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = *reinterpret_cast<const float4*>(res + col0 + 4 * j);
          }
          tmem_ld_wait();
          if (MODE == 5) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * ep.alpha);
          }
          uint8_t* sb = stg + sbuf * 4096;
          if (PROBE && (ep.dbg & 128)) { sbuf ^= 1; continue; }               // probe: TMEM read only
          if (lane == 0 && !(PROBE && (ep.dbg & 256))) bulk_wait_read<1>();   // the store that last read this buffer is done
          __syncwarp();
          const float floor_v = ep.relu ? 0.f : -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            if (ep.bias && !(PROBE && (ep.dbg & 64))) {
              v.x += __shfl_sync(0xffffffffu, bl, j); v.y += __shfl_sync(0xffffffffu, bl, j + 1);
              v.z += __shfl_sync(0xffffffffu, bl, j + 2); v.w += __shfl_sync(0xffffffffu, bl, j + 3);
            }
            v.x = fmaxf(v.x, floor_v); v.y = fmaxf(v.y, floor_v); v.z = fmaxf(v.z, floor_v); v.w = fmaxf(v.w, floor_v);
            if (res) { const float4 t4 = rv[j >> 2]; v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w; }
            *reinterpret_cast<float4*>(sb + lane * 128 + (((j >> 2) ^ (lane & 7)) << 4)) = v;
          }
          if (PROBE && (ep.dbg & 256)) { sbuf ^= 1; continue; }               // probe: no fence / TMA store
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapC, smem_u32(sb), tc.c_col + col0, tc.c_row + m0 + quad * 32);
            bulk_commit();
          }
          sbuf ^= 1;
          continue;
        }
        tmem_ld_wait();
        if (MODE == 5) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * ep.alpha);
        }
        if (fast) {
          uint8_t* sb = stg + sbuf * 4096;
          const bool lo_slab = ep.lo_tma && col0 >= ep.lo_c0 && col0 + 32 <= ep.lo_c1;   // warp-uniform: the low parts leave by TMA too
          if (lane == 0) bulk_wait_read<1>();                                 // the store that last read this buffer is done
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            if (ep.bias) {
              const float4 b = *reinterpret_cast<const float4*>(ep.bias + col0 + j);
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (rb) {
              if (rb_vec) {
                const float4 b = *reinterpret_cast<const float4*>(rb + col0 + j);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
              } else {
                v.x += rb[col0 + j]; v.y += rb[col0 + j + 1]; v.z += rb[col0 + j + 2]; v.w += rb[col0 + j + 3];
              }
            }
            if (ep.accumulate) {
              const float4 o = *reinterpret_cast<const float4*>(crow + col0 + j);
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            if (ep.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if (res) {
              if (res_vec) { const float4 t4 = *reinterpret_cast<const float4*>(res + col0 + j); v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w; }
              else { v.x += res[col0 + j]; v.y += res[col0 + j + 1]; v.z += res[col0 + j + 2]; v.w += res[col0 + j + 3]; }
            }
            // SWIZZLE_128B: 16-byte chunk q of row r sits at chunk q ^ (r & 7) (buffer is 1024-byte aligned)
            *reinterpret_cast<float4*>(sb + lane * 128 + (((j >> 2) ^ (lane & 7)) << 4)) = v;
            if (crow16) {
              const __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
              *reinterpret_cast<uint2*>(crow16 + col0 + j) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
            }
            if (crow_lo && !lo_slab && col0 + j >= ep.lo_c0 && col0 + j < ep.lo_c1) {
              float4 l;
              l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
              l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
              *reinterpret_cast<float4*>(crow_lo + col0 + j) = l;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapC, smem_u32(sb), tc.c_col + col0, tc.c_row + m0 + quad * 32);
            bulk_commit();
          }
          sbuf ^= 1;
          if (lo_slab) {
            // x - trunc_tf32(x) of the slab just staged, built in the other buffer from the staged values
            uint8_t* sb2 = stg + sbuf * 4096;
            if (lane == 0) bulk_wait_read<1>();                               // all but this chunk's own store: sb2's last reader is done
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int o = lane * 128 + (((j >> 2) ^ (lane & 7)) << 4);
              const float4 v = *reinterpret_cast<const float4*>(sb + o);
              float4 l;
              l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
              l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
              *reinterpret_cast<float4*>(sb2 + o) = l;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&mapClo, smem_u32(sb2), tc.c_col + col0, tc.c_row + m0 + quad * 32);
              bulk_commit();
            }
            sbuf ^= 1;
          }
        } else if (!ep.C && c16_tile && slab_rows_ok && col0 + 32 <= ep.N && bias_vec) {       // warp-uniform
          // bf16-only output (the A operand of the next bf16 GEMM): the lane's 32 values are packed to 64 bytes, transposed through the
          // warp's staging buffer and written as 16-byte pieces, four lanes per row -- every store instruction covers 8 full rows
          // (per-lane row stores would touch 32 different rows with 8 bytes each)
          uint8_t* sb = stg;
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[j + i]);
            if (ep.bias) {
              const float4 b0 = *reinterpret_cast<const float4*>(ep.bias + col0 + j), b1 = *reinterpret_cast<const float4*>(ep.bias + col0 + j + 4);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (rb) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += rb[col0 + j + i];
            }
            if (ep.relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (res) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += res[col0 + j + i];
            }
            __nv_bfloat162 p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            // 64-byte rows; 16-byte chunk q of row l sits at chunk q ^ ((l >> 1) & 3) (conflict-free for both access patterns)
            *reinterpret_cast<uint4*>(sb + lane * 64 + ((((j >> 3)) ^ ((lane >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(p);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (lane >> 2), q = lane & 3;
            const uint4 x = *reinterpret_cast<const uint4*>(sb + rr * 64 + ((q ^ ((rr >> 1) & 3)) << 4));
            *reinterpret_cast<uint4*>(ep.C16 + (size_t)(m0 + quad * 32 + rr) * ep.ldc16 + col0 + q * 8) = x;
          }
        } else if (row_ok) {
          if (col0 + 32 <= ep.N && (vec_ok || (!ep.C && bias_vec))) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              if (ep.bias) {
                const float4 b = *reinterpret_cast<const float4*>(ep.bias + col0 + j);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
              }
              if (rb) {
                v.x += rb[col0 + j]; v.y += rb[col0 + j + 1]; v.z += rb[col0 + j + 2]; v.w += rb[col0 + j + 3];
              }
              float4* dst = reinterpret_cast<float4*>(crow + col0 + j);
              if (ep.accumulate) {
                const float4 o = *dst;
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              }
              if (ep.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              if (res) {
              if (res_vec) { const float4 t4 = *reinterpret_cast<const float4*>(res + col0 + j); v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w; }
              else { v.x += res[col0 + j]; v.y += res[col0 + j + 1]; v.z += res[col0 + j + 2]; v.w += res[col0 + j + 3]; }
            }
              if (crow) *dst = v;
              if (crow16 && c16_vec) {
                const __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
                *reinterpret_cast<uint2*>(crow16 + col0 + j) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
              } else if (crow16) {
                crow16[col0 + j] = __float2bfloat16_rn(v.x); crow16[col0 + j + 1] = __float2bfloat16_rn(v.y);
                crow16[col0 + j + 2] = __float2bfloat16_rn(v.z); crow16[col0 + j + 3] = __float2bfloat16_rn(v.w);
              }
              if (crow_lo && col0 + j >= ep.lo_c0 && col0 + j < ep.lo_c1) {
                float4 l;
                l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                *reinterpret_cast<float4*>(crow_lo + col0 + j) = l;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              if (col < ep.N) {
                float v = __uint_as_float(r[j]);
                if (ep.bias) v += ep.bias[col];
                if (rb) v += rb[col];
                if (ep.accumulate) v += crow[col];
                if (ep.relu) v = fmaxf(v, 0.f);
                if (res) v += res[col];
                if (crow) crow[col] = v;
                if (crow16) crow16[col] = __float2bfloat16_rn(v);
                if (crow_lo && col >= ep.lo_c0 && col < ep.lo_c1) crow_lo[col] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
              }
            }
          }
        }
      }
      tc_fence_before();
      if (PAIR && rank != 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&acc_empty[acc], 0);
      } else {
        mbar_arrive(&acc_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_all();                                           // staging reads AND global writes complete before exit
  } else if ((MODE == 3 || MODE == 5) && warp >= 8) {
    // ================= bf16 correction operands of the A tile =================
    // fp32 tile: 128 rows x 16 floats, 64-byte rows, SWIZZLE_64B: 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3).
    // bf16 tiles: 128 rows x 16 bf16, 32-byte rows, SWIZZLE_32B: 16-byte chunk m of row r sits at chunk m ^ ((r >> 2) & 1).
    // two groups of 4 warps alternate stages, so that the split of stage s+1 overlaps the tail (fence + arrive) of stage s.
    // Ownership is by STAGE INDEX (not by iteration): a barrier is then only ever waited on by the group that consumed its previous
    // phase, so no waiter can run two phases ahead of `full[stage]` and alias its parity (with an odd stage count, ownership by
    // iteration let group 0 poll a stage whose previous phase -- owned by group 1 -- had not landed yet)
    const int t = (threadIdx.x - 256) & 127;  // 0..127
    const int grp = (threadIdx.x - 256) >> 7;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_step) {
      // CONV: a thread converts the same rows in every k block -- its rows' sequence positions are loaded once per tile (they were
      // two global loads per 16-byte item and k block: the top stall of the fused kernel in the round-2 ncu capture)
      constexpr int NI = TILE_A / 16 / 128;
      int m0c = 0, seq_p[NI], seq_q[NI];
      if (CONV) {
        m0c = tile_coord(ep, tile, tiles_m, tiles_n, BN, CL >= 2 ? 2 : 1, rank).m0;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int grow = m0c + ((i * 128 + t) >> 2);
          seq_p[i] = grow < ep.M ? ep.seq_pos[grow] : 0;
          seq_q[i] = grow < ep.M ? ep.seq_rem[grow] : 0;
        }
      }
      for (int kb = 0; kb < kblocks; ++kb) {
        if ((stage % CF::SPLIT_GROUPS) != grp) { if (++stage == STAGES) { stage = 0; phase ^= 1; } continue; }
        mbar_wait(&full[stage], phase);
        uint8_t* st = smem + stage * STAGE_BYTES;
        const float4* src = reinterpret_cast<const float4*>(st);
        uint8_t* lo16 = st + CF::OFF_AL;
        uint8_t* a16 = st + CF::OFF_A16;
        // CONV: a thread's items (rows i * 32 + t / 4) all fall on the SAME logical 16-byte chunk, (t & 3) ^ ((t >> 3) & 3): the depthwise
        // weights and bias of its 4 channels are loaded once per k block into registers (they were 4 shared-memory loads per tap and item:
        // the fused kernel ran at the shared-memory pipe's limit)
        float wq[CONV ? 4 : 1][CONV ? 7 : 1];
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (CONV) {
          const int cq = kb * BK + 4 * ((t & 3) ^ ((t >> 3) & 3)), k = ep.dw_k;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int j = 0; j < 7; ++j) wq[ch][j] = j < k ? sdw[(cq + ch) * k + j] : 0.f;
          bq = *reinterpret_cast<const float4*>(sdw + ep.K * k + cq);
        }
        // MODE 5 without CONV: the two fp16 tiles overwrite the fp32 tile they are made from -- every thread of the group holds its
        // items in registers before anybody writes (named barrier 1 + grp over the group's 128 threads)
        float4 xpre[(MODE == 5 && !CONV) ? NI : 1];
        if (MODE == 5 && !CONV) {
#pragma unroll
          for (int i = 0; i < NI; ++i) xpre[i] = src[i * 128 + t];
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        }
#pragma unroll
        for (int i = 0; i < ((PROBE && (ep.dbg & 2)) ? 0 : NI); ++i) {
          const int idx = i * 128 + t;
          const int r = idx >> 2, l = (idx & 3) ^ ((r >> 1) & 3);   // logical chunk l = columns 4l .. 4l+3 of row r
          float4 x;
          if (CONV) {
            // x = dwconv(X)[m0 + r][kb*16 + 4l .. +3] = b + sum_j w[c][j] * X[row + j - k/2][c] inside the row's sequence (same operation
            // order as dwconv_kernel, so the fused product is bit-identical to dwconv followed by the plain GEMM); written to the A tile
            const int grow = m0c + r, k = ep.dw_k;
            x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow < ep.M) {
              const int p = seq_p[i], q = seq_q[i];
              x = bq;
              const uint8_t* raw = st + CF::OFF_RAW;
              const int hk = k >> 1;
              if ((k == 7 || k == 3) && p >= hk && q >= hk) {
                // interior row (all taps inside its sequence: > 98 % of the rows): no per-tap range predicates -- the predicated loop below
                // made the split warps' instruction stream the limit of this kernel (round-2 ncu: IPC 2.2, issue slots 55 % busy, no hot spot)
                if (k == 7) {
#pragma unroll
                  for (int j = 0; j < 7; ++j) {
                    const int rr = r + j;
                    const float4 xin = *reinterpret_cast<const float4*>(raw + rr * 64 + ((l ^ ((rr >> 1) & 3)) << 4));
                    x.x = fmaf(wq[0][j], xin.x, x.x); x.y = fmaf(wq[1][j], xin.y, x.y);
                    x.z = fmaf(wq[2][j], xin.z, x.z); x.w = fmaf(wq[3][j], xin.w, x.w);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 3; ++j) {
                    const int rr = r + 2 + j;
                    const float4 xin = *reinterpret_cast<const float4*>(raw + rr * 64 + ((l ^ ((rr >> 1) & 3)) << 4));
                    x.x = fmaf(wq[0][j], xin.x, x.x); x.y = fmaf(wq[1][j], xin.y, x.y);
                    x.z = fmaf(wq[2][j], xin.z, x.z); x.w = fmaf(wq[3][j], xin.w, x.w);
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                  const int d = j - hk;
                  if (j >= k || d < -p || d > q) continue;
                  const int rr = r + 3 + d;
                  const float4 xin = *reinterpret_cast<const float4*>(raw + rr * 64 + ((l ^ ((rr >> 1) & 3)) << 4));
                  x.x = fmaf(wq[0][j], xin.x, x.x);
                  x.y = fmaf(wq[1][j], xin.y, x.y);
                  x.z = fmaf(wq[2][j], xin.z, x.z);
                  x.w = fmaf(wq[3][j], xin.w, x.w);
                }
              }
            }
            if (MODE == 3) reinterpret_cast<float4*>(st)[idx] = x;       // MODE 5: no MMA reads the fp32 tile
          } else if (MODE == 5) {
            x = xpre[i];
          } else {
            x = src[idx];
          }
          const int dst = r * 32 + ((((l >> 1) ^ ((r >> 2) & 1))) << 4) + (l & 1) * 8;
          if (MODE == 5) {
            // fp16 pair of the (optionally power-of-two scaled) value: hi = fp16_rn(x), lo = fp16_rn(x - hi)
            x.x *= ep.a_scale; x.y *= ep.a_scale; x.z *= ep.a_scale; x.w *= ep.a_scale;
            const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
            const __half2 l0 = __floats2half2_rn(x.x - f0.x, x.y - f0.y);
            const __half2 l1 = __floats2half2_rn(x.z - f1.x, x.w - f1.y);
            *reinterpret_cast<uint2*>(lo16 + dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
            *reinterpret_cast<uint2*>(a16 + dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            continue;
          }
          float4 lo;
          lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
          lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
          lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          const __nv_bfloat162 l0 = __floats2bfloat162_rn(lo.x, lo.y), l1 = __floats2bfloat162_rn(lo.z, lo.w);
          const __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
          *reinterpret_cast<uint2*>(lo16 + dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
          *reinterpret_cast<uint2*>(a16 + dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        }
        fence_proxy_async();
        if (PAIR && rank != 0) {            // peer CTA: one remote (DSMEM) arrive per warp on the leader's barrier
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&ready[stage], 0);
        } else {
          mbar_arrive(&ready[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (MODE == 2 && warp >= 8) {
    // ================= A hi/lo split (element-wise, so it is oblivious to the 128B swizzle) =================
    const int t = (threadIdx.x - 256) & 127;  // 0..127
    const int grp = (threadIdx.x - 256) >> 7;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_step) {
      for (int kb = 0; kb < kblocks; ++kb) {
        if ((stage % CF::SPLIT_GROUPS) != grp) { if (++stage == STAGES) { stage = 0; phase ^= 1; } continue; }
        mbar_wait(&full[stage], phase);
        float4* hi = reinterpret_cast<float4*>(smem + stage * STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + stage * STAGE_BYTES + CF::OFF_AL);
#pragma unroll
        for (int i = 0; i < ((PROBE && (ep.dbg & 2)) ? 0 : TILE_A / 16 / 128); ++i) {
          const int idx = i * 128 + t;
          const float4 x = hi[idx];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
          if (ep.store_hi) hi[idx] = h;
          lo[idx] = l;
        }
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        if (PAIR && rank != 0) {            // peer CTA: one remote (DSMEM) arrive per warp on the leader's barrier
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&ready[stage], 0);
        } else {
          mbar_arrive(&ready[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL >= 2) cluster_sync_all();        // no CTA exits while its peer can still multicast to it / arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// SIMT fp32 reference GEMM (comparator; also the fallback for shapes TMA cannot describe, e.g. K % 4 != 0)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, const GemmEpilogue ep) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float sA[TK][TM + 4];
  __shared__ float sW[TK][TN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < ep.K; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      const int r = i / TK, c = i % TK;
      sA[c][r] = (m0 + r < ep.M && k0 + c < ep.K) ? A[(size_t)(m0 + r) * lda + k0 + c] : 0.f;
      sW[c][r] = (n0 + r < ep.N && k0 + c < ep.K) ? W[(size_t)(n0 + r) * ldw + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; w[i] = sW[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= ep.M) continue;
    const float* rb = nullptr;
    if (ep.rowbias) rb = ep.rowbias + (size_t)(ep.rb_index ? ep.rb_index[row] : row % ep.rb_period) * ep.ld_rb;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= ep.N) continue;
      float v = acc[i][j];
      if (ep.bias) v += ep.bias[col];
      if (rb) v += rb[col];
      float* dst = ep.C + (size_t)row * ep.ldc + col;
      if (ep.accumulate) v += *dst;
      if (ep.relu) v = fmaxf(v, 0.f);
      if (ep.residual) v += ep.residual[(size_t)row * ep.ld_res + col];
      *dst = v;
      if (ep.C_lo) ep.C_lo[(size_t)row * ep.ldc + col] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = w[i];
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = x - h;
  }
}

// bf16 copies of a weight for mode 3: w16 = bf16_rn(w), lo16 = bf16_rn(w - trunc_tf32(w)); rows padded to ld16
__global__ void split_bf16_kernel(const float* __restrict__ w, int ldw, int rows, int cols, __nv_bfloat16* __restrict__ w16,
                                  __nv_bfloat16* __restrict__ lo16, int ld16) {
  const int64_t n = (int64_t)rows * ld16;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld16), c = (int)(i - (int64_t)r * ld16);
    const float x = c < cols ? w[(size_t)r * ldw + c] : 0.f;
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    w16[i] = __float2bfloat16_rn(x);
    lo16[i] = __float2bfloat16_rn(x - h);
  }
}

// fp32 -> bf16 (round to nearest even) copy of a row-major matrix: the A operand of the bf16 mode when its producer wrote fp32.
// One thread per 8 columns (two 16-byte loads, one 16-byte store); the padding columns [cols, ldo) are written as zeros.
__global__ void cast_bf16_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols, __nv_bfloat16* __restrict__ out, int64_t ldo) {
  const int chunks = (int)(ldo >> 3);
  const int64_t total = rows * chunks;
  const bool vec = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / chunks;
    const int c = (int)(i - r * chunks) * 8;
    float v[8];
    const float* src = x + r * ldx + c;
    if (vec && c + 8 <= cols) {
      const float4 a = ldg_stream(reinterpret_cast<const float4*>(src)), b = ldg_stream(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? src[j] : 0.f;
    }
    __nv_bfloat162 p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(out + r * ldo + c) = *reinterpret_cast<const uint4*>(p);
  }
}

// Pre-swizzled shared-memory images of the W tiles of mode 3 (BK = 16): for every (N tile, k block) one contiguous block
//   [ W fp32: bn rows x 64 B, SWIZZLE_64B | bf16(W): bn rows x 32 B, SWIZZLE_32B | bf16(W - trunc_tf32(W)): same ]
// so that the producer fetches a stage's W operands with plain contiguous bulk copies.  One thread per 16-byte fp32 chunk.
__global__ void weight_image_kernel(const float* __restrict__ w, int ldw, int N, int K, int bn, uint8_t* __restrict__ img) {
  const int kblocks = (K + 15) / 16, tiles_n = (N + bn - 1) / bn;
  const int64_t total = (int64_t)tiles_n * kblocks * bn * 4;
  const int tile_b = bn * 64;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & 3);
    const int r = (int)((i >> 2) % bn);
    const int64_t blk = (i >> 2) / bn;                 // nt * kblocks + kb
    const int kb = (int)(blk % kblocks), nt = (int)(blk / kblocks);
    const int row = nt * bn + r, col = kb * 16 + c * 4;
    float x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (row < N && col + j < K) ? w[(size_t)row * ldw + col + j] : 0.f;
    uint8_t* base = img + blk * (int64_t)(2 * tile_b);
    *reinterpret_cast<float4*>(base + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = make_float4(x[0], x[1], x[2], x[3]);
    float lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) lo[j] = x[j] - __uint_as_float(__float_as_uint(x[j]) & 0xFFFFE000u);
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(x[0], x[1]), h1 = __floats2bfloat162_rn(x[2], x[3]);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(lo[0], lo[1]), l1 = __floats2bfloat162_rn(lo[2], lo[3]);
    const int dst = r * 32 + ((((c >> 1) ^ ((r >> 2) & 1))) << 4) + (c & 1) * 8;
    *reinterpret_cast<uint2*>(base + tile_b + dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(base + tile_b + tile_b / 2 + dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  }
}

// Weight-tile images of mode 5 (fp16x3, BK = 16): for every (N tile, k block) one contiguous block of two fp16 sub-images, bn rows x 32 B
// each, SWIZZLE_32B:  [ hi = fp16_rn(w 2^s) | fp16_rn(w 2^s - hi) ].  One thread per 4 elements.
__global__ void weight_image_fp16_kernel(const float* __restrict__ w, int ldw, int N, int K, int bn, float scale, uint8_t* __restrict__ img) {
  const int kblocks = (K + 15) / 16, tiles_n = (N + bn - 1) / bn;
  const int64_t total = (int64_t)tiles_n * kblocks * bn * 4;
  const int sub = bn * 32;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & 3);
    const int r = (int)((i >> 2) % bn);
    const int64_t blk = (i >> 2) / bn;                 // nt * kblocks + kb
    const int kb = (int)(blk % kblocks), nt = (int)(blk / kblocks);
    const int row = nt * bn + r, col = kb * 16 + c * 4;
    __half hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = ((row < N && col + j < K) ? w[(size_t)row * ldw + col + j] : 0.f) * scale;
      hi[j] = __float2half_rn(x);
      lo[j] = __float2half_rn(x - __half2float(hi[j]));
    }
    uint8_t* base = img + blk * (int64_t)(2 * sub);
    const int dst = r * 32 + ((((c >> 1) ^ ((r >> 2) & 1))) << 4) + (c & 1) * 8;
    *reinterpret_cast<uint2*>(base + dst) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(base + sub + dst) = *reinterpret_cast<const uint2*>(lo);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; int rows, cols, ld, box_rows, box_cols, esize;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols &&
           esize == o.esize;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (std::hash<int>()(k.rows) * 31) ^ (std::hash<int>()(k.cols) * 131) ^ (std::hash<int>()(k.ld) * 1031) ^
           (std::hash<int>()(k.box_rows) * 7919) ^ (std::hash<int>()(k.box_cols) * 104729) ^ (size_t)k.esize;
  }
};
static int g_tma_store = 1;
static int g_w_image = 1;
static int g_cluster = 3;
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 2-D row-major [rows][cols] (fp32, or bf16 when esize == 2) with leading dimension ld (elements); box = box_cols x box_rows,
// swizzle span = the box row (128 / 64 / 32 bytes), zero OOB fill
static int get_tensor_map(const void* base, int rows, int cols, int ld, int box_rows, int box_cols, CUtensorMap* out, int esize = 4) {
  MapKey key{base, rows, cols, ld, box_rows, box_cols, esize};
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return VSG_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (driver entry point lookup failed)"); return VSG_E_LAUNCH; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  const int row_bytes = box_cols * esize;
  CUresult r = enc(&m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for [%d x %d] ld %d", (int)r, rows, cols, ld); return VSG_E_LAUNCH; }
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = m;
  }
  *out = m;
  return VSG_OK;
}

// B16 (MODE 1): `A` / `Wh` point to bf16 operands ([rows][lda] / [N][ldw] bf16 elements)
template <int MODE, int BN_, int CL, bool CONV = false, bool B16 = false>
static int launch_tc(const float* A, int lda, int a_rows, int a_cols, const float* Wh, const void* Wl, int ldw, int w_rows, int w_cols,
                     const GemmEpilogue& ep_in, cudaStream_t st, const void* W16 = nullptr, int ldw16 = 0) {
  using CF = Cfg<MODE, BN_, CL == 3, CONV>;
  constexpr int CLN = CL >= 2 ? 2 : 1;            // CTAs per cluster
  CUtensorMap mA, mBh, mBl, mB16;
  constexpr int WBOX = BN_ / CLN;                 // clusters: each CTA of a pair loads (CL == 2: and multicasts) half of the W rows
  constexpr int ES = B16 ? 2 : 4, BKE = B16 ? 2 * CF::BK : CF::BK;
  int rc = get_tensor_map(A, a_rows, a_cols, lda, CONV ? CF::RAW_ROWS : BM, BKE, &mA, ES);      // CONV: raw X tile with its halo rows
  if (rc) return rc;
  rc = get_tensor_map(Wh, w_rows, w_cols, ldw, WBOX, BKE, &mBh, ES);
  if (rc) return rc;
  if (MODE == 2) { rc = get_tensor_map(Wl, w_rows, w_cols, ldw, WBOX, CF::BK, &mBl); if (rc) return rc; } else mBl = mBh;
  mB16 = mBh;
  if (MODE == 3) {
    rc = get_tensor_map(Wl, w_rows, w_cols, ldw16, WBOX, CF::BK, &mBl, 2);
    if (rc) return rc;
    rc = get_tensor_map(W16, w_rows, w_cols, ldw16, WBOX, CF::BK, &mB16, 2);
    if (rc) return rc;
  }
  GemmEpilogue ep = ep_in;
  CUtensorMap mC = mA, mClo = mA;
  ep.tma_store = 0; ep.lo_tma = 0;
  if (g_tma_store && ep.C && (ep.ldc & 3) == 0 && aligned16(ep.C)) {
    // C as a 2-D tensor of 32x32 fp32 boxes (SWIZZLE_128B).  Batched problems address boxes inside the whole C buffer, whose row
    // length is ldc; only slabs that lie fully inside a problem take this path, so the bounds are never relied on for clipping.
    long long rows = ep.M, cols = ep.N;
    if (ep.batch > 1) {
      const long long last = (long long)((ep.batch - 1) / ep.batch_inner) * ep.c_outer + (long long)(ep.batch_inner - 1) * ep.c_inner;
      rows = last / ep.ldc + ep.M + 1;
      cols = ep.ldc;
    }
    if (rows < 0x7fffffffLL && ep.c_outer >= 0 && ep.c_inner >= 0) {
      rc = get_tensor_map(ep.C, (int)rows, (int)cols, ep.ldc, 32, 32, &mC);
      if (rc) return rc;
      ep.tma_store = 1;
      if (ep.C_lo && aligned16(ep.C_lo)) {
        rc = get_tensor_map(ep.C_lo, (int)rows, (int)cols, ep.ldc, 32, 32, &mClo);
        if (rc) return rc;
        ep.lo_tma = 1;
      }
    }
  }
  constexpr int SMEM = CF::STAGES * CF::STAGE_BYTES + CF::STAGING_BYTES + 1024 /*align*/ + CF::BAR_BYTES + CF::DW_BYTES;
  constexpr bool HAS_PROBE = (CL == 1) && !CONV && !B16;   // the timing probes exist for the single-CTA kernel only (compile time)
  static PerDeviceFlag attr_set;
  const int dev_ = current_device();
  if (!attr_set.is_set(dev_)) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<MODE, BN_, false, CL, CONV, B16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(gemm_tc_kernel<MODE, BN_, HAS_PROBE, CL, CONV, B16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(max dynamic smem %d) failed: %s", SMEM, cudaGetErrorString(cudaGetLastError()));
      return VSG_E_LAUNCH;
    }
    attr_set.set(dev_);
  }
  const long long tiles_m = (ep.M + BM - 1) / BM;
  const long long work = (CLN == 2 ? (tiles_m + 1) / 2 : tiles_m) * ((ep.N + BN_ - 1) / BN_) * ep.batch;   // tiles, or tile pairs
  const int slots = sm_count() / CLN;
  const int grid = (int)(work < slots ? work : slots) * CLN;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(CF::THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLN; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = CLN > 1 ? 1 : 0;
  cudaError_t e = (ep.dbg && HAS_PROBE) ? cudaLaunchKernelEx(&cfg, gemm_tc_kernel<MODE, BN_, HAS_PROBE, CL, CONV, B16>, mA, mBh, mBl, mB16, mC, mClo, ep)   // timing probes
                                        : cudaLaunchKernelEx(&cfg, gemm_tc_kernel<MODE, BN_, false, CL, CONV, B16>, mA, mBh, mBl, mB16, mC, mClo, ep);
  if (e != cudaSuccess) { set_error("vsg_gemm(tcgen05): launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return VSG_E_LAUNCH; }
  return check_launch("vsg_gemm(tcgen05)");
}

template <int MODE, int BN_, bool B16 = false>
static int launch_tc_auto(const float* A, int lda, int a_rows, int a_cols, const float* Wh, const void* Wl, int ldw, int w_rows, int w_cols,
                          const GemmEpilogue& ep, cudaStream_t st, const void* W16 = nullptr, int ldw16 = 0) {
  // plain problems with at least one full pair of M tiles run on CTA pairs: one cta_group::2 MMA per pair (split modes and the bf16
  // mode, 256-wide tiles), else two per-CTA MMAs with the W tile multicast
  if (BN_ == 256 && (MODE >= 2 || B16) && g_cluster == 3 && ep.batch == 1 && ep.M > BM && !ep.dbg)
    return launch_tc<MODE, 256, 3, false, B16>(A, lda, a_rows, a_cols, Wh, Wl, ldw, w_rows, w_cols, ep, st, W16, ldw16);
  if (g_cluster >= 2 && ep.batch == 1 && ep.M > BM && !ep.dbg)
    return launch_tc<MODE, BN_, 2, false, B16>(A, lda, a_rows, a_cols, Wh, Wl, ldw, w_rows, w_cols, ep, st, W16, ldw16);
  return launch_tc<MODE, BN_, 1, false, B16>(A, lda, a_rows, a_cols, Wh, Wl, ldw, w_rows, w_cols, ep, st, W16, ldw16);
}

// N=256 tiles whenever they do not add MMA work over N=128 tiles
static inline bool use_bn256(int N) { return 2 * ((N + 255) / 256) <= (N + 127) / 128; }

}  // namespace vsg

using namespace vsg;

static int g_store_hi = 0;
static int g_dbg = 0;
/* timing probes (results become garbage): 1 skip W loads, 2 skip the A split, 4 skip MMAs, 8 skip A loads, 16 free stages by a plain arrive, 32 producer / issuer poll with test_wait;
   lean epilogue: 64 skip the bias shuffles, 128 stop after the TMEM read, 256 skip the fence + TMA store; 0 = normal */
/* validation knob: 0 = the epilogue writes C with per-row 16-byte stores only, 1 (default) = full 32x32 slabs leave through TMA stores */
extern "C" int vsg_gemm_set_tma_store(int on) { int old = vsg::g_tma_store; vsg::g_tma_store = on ? 1 : 0; return old; }
/* validation knob: 1 = one CTA per tile; 2 = CTA pairs that multicast the W tile (per-CTA MMAs); 3 (default) = CTA-pair MMAs
   (tcgen05 cta_group::2, M = 256) where available (split modes, 256-wide tiles), else as 2 */
extern "C" int vsg_gemm_set_cluster(int n) { int old = vsg::g_cluster; vsg::g_cluster = (n >= 1 && n <= 3) ? n : 3; return old; }
/* validation knob: 0 = mode 3 loads its W operands through tensor maps even when a weight image is supplied */
extern "C" int vsg_gemm_set_weight_image(int on) { int old = vsg::g_w_image; vsg::g_w_image = on ? 1 : 0; return old; }
extern "C" int vsg_gemm_debug_flags(int f) { int old = g_dbg; g_dbg = f; return old; }
static int g_force_bn = 0;
/* debug/validation knob: 128 forces the N=128 tile kernel everywhere, 0 = automatic */
extern "C" int vsg_gemm_force_bn(int bn) { int old = g_force_bn; g_force_bn = bn; return old; }
/* debug/validation knob: 1 = the 3xTF32 split also rewrites the A tile with its masked high part */
extern "C" int vsg_gemm_set_store_hi(int on) { int old = g_store_hi; g_store_hi = on ? 1 : 0; return old; }

extern "C" int vsg_split_tf32(const float* w, float* hi, float* lo, int64_t n, void* stream) {
  VSG_REQUIRE(n >= 0, "vsg_split_tf32: n < 0");
  if (n == 0) return VSG_OK;
  VSG_REQUIRE(w && hi && lo, "vsg_split_tf32: null pointer");
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  split_tf32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, hi, lo, n);
  return check_launch("vsg_split_tf32");
}

extern "C" int vsg_split_bf16(const float* w, int ldw, int rows, int cols, void* w16, void* lo16, int ld16, void* stream) {
  VSG_REQUIRE(rows >= 0 && cols >= 0 && ldw >= cols && ld16 >= cols, "vsg_split_bf16: bad extents");
  if (rows == 0 || cols == 0) return VSG_OK;
  VSG_REQUIRE(w && w16 && lo16, "vsg_split_bf16: null pointer");
  int64_t blocks = ((int64_t)rows * ld16 + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  split_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, ldw, rows, cols, (__nv_bfloat16*)w16, (__nv_bfloat16*)lo16, ld16);
  return check_launch("vsg_split_bf16");
}

extern "C" int vsg_cast_bf16(const float* x, int64_t ldx, int64_t rows, int cols, void* out, int64_t ldo, void* stream) {
  VSG_REQUIRE(rows >= 0 && cols >= 0 && ldx >= cols && ldo >= cols && ldo % 8 == 0, "vsg_cast_bf16: bad extents (ldo must be a multiple of 8)");
  if (rows == 0 || cols == 0) return VSG_OK;
  VSG_REQUIRE(x && out && aligned16(out), "vsg_cast_bf16: null or unaligned pointer");
  const int64_t total = rows * (ldo / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 32) blocks = (int64_t)sm_count() * 32;
  cast_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, cols, (__nv_bfloat16*)out, ldo);
  return check_launch("vsg_cast_bf16");
}

extern "C" int vsg_gemm_tile_n(int N) { return (use_bn256(N) && g_force_bn != 128) ? 256 : 128; }

extern "C" int64_t vsg_weight_image_bytes(int N, int K, int bn) {
  if (N <= 0 || K <= 0 || (bn != 128 && bn != 256)) return 0;
  return (int64_t)((N + bn - 1) / bn) * ((K + 15) / 16) * (int64_t)(2 * bn * 64);
}

extern "C" int64_t vsg_weight_image_fp16_bytes(int N, int K, int bn) {
  if (N <= 0 || K <= 0 || (bn != 128 && bn != 256)) return 0;
  return (int64_t)((N + bn - 1) / bn) * ((K + 15) / 16) * (int64_t)(2 * bn * 32);
}

extern "C" int vsg_build_weight_image_fp16(const float* w, int ldw, int N, int K, int bn, float scale, void* img, void* stream) {
  VSG_REQUIRE(N > 0 && K > 0 && ldw >= K && (bn == 128 || bn == 256), "vsg_build_weight_image_fp16: bad extents");
  VSG_REQUIRE(w && img && aligned16(img), "vsg_build_weight_image_fp16: null or unaligned pointer");
  int ex = 0;
  VSG_REQUIRE(scale > 0.f && frexpf(scale, &ex) == 0.5f, "vsg_build_weight_image_fp16: scale must be a power of two");
  const int64_t total = (int64_t)((N + bn - 1) / bn) * ((K + 15) / 16) * bn * 4;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  weight_image_fp16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, ldw, N, K, bn, scale, (uint8_t*)img);
  return check_launch("vsg_build_weight_image_fp16");
}

extern "C" int vsg_build_weight_image(const float* w, int ldw, int N, int K, int bn, void* img, void* stream) {
  VSG_REQUIRE(N > 0 && K > 0 && ldw >= K && (bn == 128 || bn == 256), "vsg_build_weight_image: bad extents");
  VSG_REQUIRE(w && img && aligned16(img), "vsg_build_weight_image: null or unaligned pointer");
  const int64_t total = (int64_t)((N + bn - 1) / bn) * ((K + 15) / 16) * bn * 4;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  weight_image_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, ldw, N, K, bn, (uint8_t*)img);
  return check_launch("vsg_build_weight_image");
}

extern "C" int vsg_gemm_ex(const VsgGemmArgs* a, void* stream) {
  VSG_REQUIRE(a != nullptr, "vsg_gemm_ex: null args");
  const int M = a->M, N = a->N, K = a->K;
  VSG_REQUIRE(M >= 0 && N >= 0 && K >= 0, "vsg_gemm: negative size");
  const int batch = a->batch > 0 ? a->batch : 1;
  if (M == 0 || N == 0) return VSG_OK;
  if (a->mode == 4) {
    // bf16 operands (A16 [M][lda16], W_b16 [N][ldw16]), fp32 accumulate, C as fp32 and / or bf16
    VSG_REQUIRE(batch == 1, "vsg_gemm_ex: mode 4 (bf16) takes plain problems only");
    VSG_REQUIRE(a->A16 && a->W_b16 && (a->C || a->C16), "vsg_gemm_ex: mode 4 needs A16, W_b16 and C and / or C16");
    VSG_REQUIRE(aligned16(a->A16) && aligned16(a->W_b16) && a->lda16 >= K && a->ldw16 >= K && a->lda16 % 8 == 0 && a->ldw16 % 8 == 0,
                "vsg_gemm_ex: mode 4 operands must be 16-byte aligned with leading dimensions that are multiples of 8 and >= K");
    VSG_REQUIRE(!a->C || a->ldc >= N, "vsg_gemm_ex: ldc < N");
    VSG_REQUIRE(!a->C16 || (a->ldc16 >= N && (reinterpret_cast<uintptr_t>(a->C16) & 1) == 0), "vsg_gemm_ex: bad C16");
    VSG_REQUIRE(!a->accumulate || a->C, "vsg_gemm_ex: accumulate needs C");
    VSG_REQUIRE(!a->C_lo && !a->dw_w, "vsg_gemm_ex: mode 4 has no C_lo / fused depthwise conv");
    VSG_REQUIRE(a->rowbias == nullptr || a->rb_index != nullptr || a->rb_period > 0, "vsg_gemm: rowbias needs rb_index or rb_period");
    GemmEpilogue ep;
    memset(&ep, 0, sizeof(ep));
    ep.bias = a->bias; ep.rowbias = a->rowbias; ep.rb_index = a->rb_index; ep.rb_period = a->rb_period; ep.ld_rb = a->ld_rb;
    ep.relu = a->relu; ep.accumulate = a->accumulate; ep.residual = a->residual; ep.ld_res = a->ld_res;
    ep.C = a->C; ep.ldc = a->C ? a->ldc : N; ep.C16 = (__nv_bfloat16*)a->C16; ep.ldc16 = a->ldc16;
    ep.M = M; ep.N = N; ep.K = K; ep.lo_c1 = N; ep.batch = 1; ep.batch_inner = 1;
    const float* A16 = reinterpret_cast<const float*>(a->A16);
    const float* W16f = reinterpret_cast<const float*>(a->W_b16);
    const bool wide4 = use_bn256(N) && g_force_bn != 128;
    return wide4 ? launch_tc_auto<1, 256, true>(A16, a->lda16, M, K, W16f, nullptr, a->ldw16, N, K, ep, (cudaStream_t)stream)
                 : launch_tc_auto<1, 128, true>(A16, a->lda16, M, K, W16f, nullptr, a->ldw16, N, K, ep, (cudaStream_t)stream);
  }
  VSG_REQUIRE(a->A && a->W_hi && a->C, "vsg_gemm: null matrix pointer");
  VSG_REQUIRE(a->lda >= K && a->ldw >= K && a->ldc >= N, "vsg_gemm: leading dimension smaller than extent");
  VSG_REQUIRE(a->rowbias == nullptr || a->rb_index != nullptr || a->rb_period > 0, "vsg_gemm: rowbias needs rb_index or rb_period");
  GemmEpilogue ep;
  ep.bias = a->bias; ep.rowbias = a->rowbias; ep.rb_index = a->rb_index; ep.rb_period = a->rb_period; ep.ld_rb = a->ld_rb;
  ep.store_hi = g_store_hi; ep.dbg = g_dbg;
  ep.relu = a->relu; ep.accumulate = a->accumulate; ep.residual = a->residual; ep.ld_res = a->ld_res; ep.C = a->C; ep.C_lo = a->C_lo;
  ep.C16 = (__nv_bfloat16*)a->C16; ep.ldc16 = a->ldc16;
  ep.lo_c0 = 0; ep.lo_c1 = N; ep.tma_store = 0; ep.lo_tma = 0; ep.w_img = nullptr;
  ep.dw_w = nullptr; ep.dw_b = nullptr; ep.seq_pos = nullptr; ep.seq_rem = nullptr; ep.dw_k = 0;
  VSG_REQUIRE(a->dw_w == nullptr || ((a->mode == 3 || a->mode == 5) && batch == 1), "vsg_gemm_ex: the fused depthwise conv exists for modes 3 and 5, plain problems");
  ep.alpha = 1.f; ep.a_scale = 1.f;
  if (a->lo_col_end > a->lo_col_begin) { ep.lo_c0 = a->lo_col_begin; ep.lo_c1 = a->lo_col_end; }
  ep.ldc = a->ldc; ep.M = M; ep.N = N; ep.K = K;
  ep.batch = batch; ep.batch_inner = a->batch_inner > 0 ? a->batch_inner : 1;
  ep.a_row_outer = a->a_row_outer; ep.a_row_inner = a->a_row_inner; ep.a_col_outer = a->a_col_outer; ep.a_col_inner = a->a_col_inner;
  ep.b_row_outer = a->b_row_outer; ep.b_row_inner = a->b_row_inner; ep.b_col_outer = a->b_col_outer; ep.b_col_inner = a->b_col_inner;
  ep.c_outer = a->c_outer; ep.c_inner = a->c_inner;
  if (batch == 1) {
    ep.a_row_outer = ep.a_row_inner = ep.a_col_outer = ep.a_col_inner = 0;
    ep.b_row_outer = ep.b_row_inner = ep.b_col_outer = ep.b_col_inner = 0;
    ep.c_outer = ep.c_inner = 0;
  }
  const int a_rows = batch > 1 ? a->a_rows : M, a_cols = batch > 1 ? a->a_cols : K;
  const int w_rows = batch > 1 ? a->w_rows : N, w_cols = batch > 1 ? a->w_cols : K;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tma_ok = (a->lda % 4 == 0) && (a->ldw % 4 == 0) && aligned16(a->A) && aligned16(a->W_hi) && K >= 1;
  const int mode = a->mode;
  if (batch > 1) {
    VSG_REQUIRE(mode != 0 && tma_ok, "vsg_gemm_ex: batched problems need a tensor-core mode and TMA-compatible operands");
    VSG_REQUIRE(K % 32 == 0, "vsg_gemm_ex: batched problems need K %% 32 == 0 (K blocks must not straddle problems)");
    VSG_REQUIRE(a->a_rows > 0 && a->a_cols > 0 && a->w_rows > 0 && a->w_cols > 0, "vsg_gemm_ex: batched problems need the full operand extents");
    VSG_REQUIRE(!a->bias && !a->rowbias && !a->residual && !a->accumulate, "vsg_gemm_ex: batched problems take no bias / residual / accumulate");
  }
  if (mode == 0 || !tma_ok) {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    gemm_simt_kernel<<<grid, 256, 0, st>>>(a->A, a->lda, a->W_hi, a->ldw, ep);
    return check_launch("vsg_gemm(simt)");
  }
  const bool wide = use_bn256(N) && g_force_bn != 128;
  if (mode == 1)
    return wide ? launch_tc_auto<1, 256>(a->A, a->lda, a_rows, a_cols, a->W_hi, nullptr, a->ldw, w_rows, w_cols, ep, st)
                : launch_tc_auto<1, 128>(a->A, a->lda, a_rows, a_cols, a->W_hi, nullptr, a->ldw, w_rows, w_cols, ep, st);
  if (mode == 2) {
    VSG_REQUIRE(a->W_lo && aligned16(a->W_lo), "vsg_gemm: mode 2 (3xTF32) needs the pre-split low part of W");
    return wide ? launch_tc_auto<2, 256>(a->A, a->lda, a_rows, a_cols, a->W_hi, a->W_lo, a->ldw, w_rows, w_cols, ep, st)
                : launch_tc_auto<2, 128>(a->A, a->lda, a_rows, a_cols, a->W_hi, a->W_lo, a->ldw, w_rows, w_cols, ep, st);
  }
  if (mode == 3) {
    VSG_REQUIRE(batch == 1, "vsg_gemm_ex: mode 3 (tf32 + 2 x bf16) takes plain problems only; batched attention problems use mode 2");
    VSG_REQUIRE(a->W_b16 && a->W_lo16 && aligned16(a->W_b16) && aligned16(a->W_lo16) && a->ldw16 >= K && a->ldw16 % 8 == 0,
                "vsg_gemm_ex: mode 3 needs the bf16 copies of W from vsg_split_bf16 (16-byte aligned, ldw16 a multiple of 8)");
    if (g_w_image && a->W_img && a->img_bn == (wide ? 256 : 128) && aligned16(a->W_img)) ep.w_img = (const uint8_t*)a->W_img;
    if (a->dw_w) {
      // A = dwconv(X) computed inside the kernel (grounding conv blocks: depthwise k in {3, 7} over K = 128 channels, then the 1x1 conv)
      VSG_REQUIRE(!wide && N <= 128, "vsg_gemm_ex: the fused depthwise conv needs N <= 128 (128-wide tiles)");
      VSG_REQUIRE(a->dw_b && a->seq_pos && a->seq_rem && aligned16(a->dw_b), "vsg_gemm_ex: fused depthwise conv needs dw_b, seq_pos, seq_rem");
      VSG_REQUIRE(a->dw_k >= 1 && a->dw_k <= 7 && (a->dw_k & 1) && K % 16 == 0 && K * (a->dw_k + 1) <= 2048,
                  "vsg_gemm_ex: fused depthwise conv supports odd k <= 7, K %% 16 == 0 and K * (k + 1) <= 2048 (got k %d, K %d)", a->dw_k, K);
      ep.dw_w = a->dw_w; ep.dw_b = a->dw_b; ep.seq_pos = a->seq_pos; ep.seq_rem = a->seq_rem; ep.dw_k = a->dw_k;
      return launch_tc<3, 128, 1, true>(a->A, a->lda, a_rows, a_cols, a->W_hi, a->W_lo16, a->ldw, w_rows, w_cols, ep, st, a->W_b16, a->ldw16);
    }
    return wide ? launch_tc_auto<3, 256>(a->A, a->lda, a_rows, a_cols, a->W_hi, a->W_lo16, a->ldw, w_rows, w_cols, ep, st, a->W_b16, a->ldw16)
                : launch_tc_auto<3, 128>(a->A, a->lda, a_rows, a_cols, a->W_hi, a->W_lo16, a->ldw, w_rows, w_cols, ep, st, a->W_b16, a->ldw16);
  }
  if (mode == 5) {
    VSG_REQUIRE(batch == 1, "vsg_gemm_ex: mode 5 (fp16x3) takes plain problems only; batched attention problems use mode 2");
    VSG_REQUIRE(a->W_img16 && aligned16(a->W_img16) && a->img16_bn == (wide ? 256 : 128),
                "vsg_gemm_ex: mode 5 needs the fp16 weight-tile image of vsg_build_weight_image_fp16 for tile width %d", wide ? 256 : 128);
    VSG_REQUIRE(a->w_alpha > 0.f, "vsg_gemm_ex: mode 5 needs w_alpha = 1 / (the image's scale)");
    {
      int ex = 0;
      VSG_REQUIRE(a->a_scale == 0.f || (a->a_scale > 0.f && frexpf(a->a_scale, &ex) == 0.5f), "vsg_gemm_ex: a_scale must be 0 (= 1) or a power of two");
    }
    ep.a_scale = a->a_scale > 0.f ? a->a_scale : 1.f;
    ep.w_img = (const uint8_t*)a->W_img16; ep.alpha = a->w_alpha / ep.a_scale;
    if (a->dw_w) {
      VSG_REQUIRE(!wide && N <= 128, "vsg_gemm_ex: the fused depthwise conv needs N <= 128 (128-wide tiles)");
      VSG_REQUIRE(a->dw_b && a->seq_pos && a->seq_rem && aligned16(a->dw_b), "vsg_gemm_ex: fused depthwise conv needs dw_b, seq_pos, seq_rem");
      VSG_REQUIRE(a->dw_k >= 1 && a->dw_k <= 7 && (a->dw_k & 1) && K % 16 == 0 && K * (a->dw_k + 1) <= 2048,
                  "vsg_gemm_ex: fused depthwise conv supports odd k <= 7, K %% 16 == 0 and K * (k + 1) <= 2048 (got k %d, K %d)", a->dw_k, K);
      ep.dw_w = a->dw_w; ep.dw_b = a->dw_b; ep.seq_pos = a->seq_pos; ep.seq_rem = a->seq_rem; ep.dw_k = a->dw_k;
      return launch_tc<5, 128, 1, true>(a->A, a->lda, a_rows, a_cols, a->W_hi, nullptr, a->ldw, w_rows, w_cols, ep, st);
    }
    return wide ? launch_tc_auto<5, 256>(a->A, a->lda, a_rows, a_cols, a->W_hi, nullptr, a->ldw, w_rows, w_cols, ep, st)
                : launch_tc_auto<5, 128>(a->A, a->lda, a_rows, a_cols, a->W_hi, nullptr, a->ldw, w_rows, w_cols, ep, st);
  }
  set_error("vsg_gemm: unknown mode %d", mode);
  return VSG_E_UNSUPPORTED;
}

extern "C" int vsg_gemm(int mode, const float* A, int lda, const float* W_hi, const float* W_lo, int ldw, int M, int N, int K,
                        const float* bias, const float* rowbias, const int32_t* rb_index, int rb_period, int ld_rb, int relu,
                        int accumulate, const float* residual, int ld_res, float* C, int ldc, void* stream) {
  VsgGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = mode; a.A = A; a.lda = lda; a.W_hi = W_hi; a.W_lo = W_lo; a.ldw = ldw; a.M = M; a.N = N; a.K = K;
  a.bias = bias; a.rowbias = rowbias; a.rb_index = rb_index; a.rb_period = rb_period; a.ld_rb = ld_rb; a.relu = relu;
  a.accumulate = accumulate; a.residual = residual; a.ld_res = ld_res; a.C = C; a.ldc = ldc; a.batch = 1; a.batch_inner = 1;
  return vsg_gemm_ex(&a, stream);
}
