// Multi-head attention with head_dim 16 on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// The grounding network's three QANet encoders run nn.MultiheadAttention(128, 8 heads) (models/grd_model_v5.py:90, :100-108, :128-130)
// over ragged sequences: the video encoder (one sequence of T clips), the query encoder (nq sequences of 3 words) and the combined
// encoder (nq sequences of T clips) -- the last one is 4 T^2 128 flops per query, the largest non-GEMM kernel of the VidOR step.
//
// One CTA = (sequence, block of 128 queries, head).  Per block of KC (32 or 64) keys:
//   S  = Q K^T      tcgen05.mma kind::tf32, M = 128, N = KC, K = 16: Q / K tiles staged by the CTA's threads as K-major SWIZZLE_64B
//                   tiles (64-byte rows = the 16 floats of one head), accumulator in TMEM columns [0, 64)
//   P  = exp(...)   the 128 threads read their own row of S with tcgen05.ld (one TMEM lane each), exponentiate and write P as the
//                   A operand of the second product: two K-major SWIZZLE_128B panels of 32 keys
//   O += P V        tcgen05.mma M = 128, N = 16, K = KC against V^T (staged transposed: 16 rows x 32 keys per panel), accumulator in
//                   TMEM columns [KC, KC + 16)
// Softmax is online (running row maximum and sum; when a block raises the maximum the TMEM accumulator is rescaled through
// tcgen05.ld / tcgen05.st), i.e. softmax(s - max) of the reference up to rounding (measured <= 4e-6 against fp64).
// fp32-class precision (`products` = 3, the default of the fp32-class GEMM modes): every operand is split into hi = trunc_tf32(x) (the raw
// fp32 value -- the MMA ignores the low 13 mantissa bits) and lo = x - hi, and each product is lo*hi + hi*lo + hi*hi (3xTF32, error ~2^-21
// per product).  `products` = 1 runs the hi*hi product only (tf32, the reduced-precision modes).
// Shared memory 56 KB (KC = 32) / 96 KB (KC = 64), TMEM 64 / 128 columns -> three / two CTAs per SM, so one CTA's exponentials overlap
// the others' MMA round trips (the kernel is latency-bound: K = 16 and N = 16 MMAs are tiny).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>

namespace vsg {

constexpr int AT_BM = 128;   // queries per CTA (MMA M)
constexpr int AT_DH = 16;    // head dimension
// KC = keys per block (MMA N of the first product, K of the second): 64 -> 96 KB of shared memory, 2 CTAs / SM; 32 -> 56 KB, 4 CTAs / SM
// (twice the MMA round trips per key, but twice the CTAs to overlap them with -- the kernel is latency-bound).
// shared-memory map (bytes, every tile 1024-byte aligned): Q hi/lo [128 rows x 64 B] SWIZZLE_64B | K hi/lo [KC rows x 64 B] SWIZZLE_64B |
// V^T hi/lo KC/32 panels x [16 rows x 128 B] SWIZZLE_128B | P hi/lo KC/32 panels x [128 rows x 128 B] SWIZZLE_128B
template <int KC> struct AtCfg {
  static constexpr int PANELS = KC / 32;
  static constexpr int OFF_QH = 0, OFF_QL = 8192;
  static constexpr int OFF_KH = 16384, OFF_KL = OFF_KH + KC * 64;
  static constexpr int OFF_VH = OFF_KL + KC * 64, OFF_VL = OFF_VH + PANELS * 2048;
  static constexpr int OFF_PH = ((OFF_VL + PANELS * 2048 + 1023) / 1024) * 1024, OFF_PL = OFF_PH + PANELS * 16384;
  static constexpr int SMEM = OFF_PL + PANELS * 16384;
  static constexpr int TMEM_COLS = KC == 64 ? 128 : 64;        // S: columns [0, KC), O: [KC, KC + 16)
  static constexpr int CTAS = KC == 64 ? 2 : 4;
};

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float4 tf32_lo(float4 x) { return make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w)); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <int KC>
__global__ void __launch_bounds__(128, AtCfg<KC>::CTAS)
mha16_tc_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
                const int64_t* __restrict__ seg_off, const int32_t* __restrict__ blk_seg, const int32_t* __restrict__ blk_q0,
                float scale_log2e, int products, float* __restrict__ O, int ldo) {
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  using CF = AtCfg<KC>;
  constexpr int AT_KC = KC, AT_OFF_QH = CF::OFF_QH, AT_OFF_QL = CF::OFF_QL, AT_OFF_KH = CF::OFF_KH, AT_OFF_KL = CF::OFF_KL;
  constexpr int AT_OFF_VH = CF::OFF_VH, AT_OFF_VL = CF::OFF_VL, AT_OFF_PH = CF::OFF_PH, AT_OFF_PL = CF::OFF_PL, AT_TMEM_COLS = CF::TMEM_COLS;
  constexpr int NX = KC / 32;                       // 16-byte K / V items per thread and block
  const int tid = threadIdx.x, warp = tid >> 5;
  const int seg = blk_seg[blockIdx.x], q0 = blk_q0[blockIdx.x], h = blockIdx.y;
  const int64_t r0 = seg_off[seg];
  const int T = (int)(seg_off[seg + 1] - r0);
  const int col0 = h * AT_DH;

  if (tid == 0) { mbar_init(&mma_bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, AT_TMEM_COLS); tmem_relinquish(); }
  // ---- Q tile: row r = query q0 + r, 4 chunks of 16 bytes; SWIZZLE_64B: chunk c of row r sits at chunk c ^ ((r >> 1) & 3) ----
  for (int idx = tid; idx < AT_BM * 4; idx += 128) {
    const int r = idx >> 2, c = idx & 3;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) x = *reinterpret_cast<const float4*>(Q + (r0 + q0 + r) * (int64_t)ldq + col0 + 4 * c);
    const int o = r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
    *reinterpret_cast<float4*>(smem + AT_OFF_QH + o) = x;
    *reinterpret_cast<float4*>(smem + AT_OFF_QL + o) = tf32_lo(x);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_s = tmem, t_o = tmem + AT_KC;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;          // this warp's TMEM lane quadrant
  const uint32_t sbase = smem_u32(smem);
  const int row = tid;                                             // the query row (TMEM lane) this thread owns
  uint32_t phase = 0;
  const int n_blocks = (T + AT_KC - 1) / AT_KC;

  // K / V blocks go global -> registers -> shared memory in two steps, so that the loads of block b+1 are in flight while block b is
  // multiplied and exponentiated (the round-2 ncu capture had a third of all warp samples waiting on these loads).
  // Item idx = it * 128 + tid: key r = idx >> 2 of the block, 16-byte chunk c = idx & 3 (dims 4c .. 4c+3).
  auto load_kv = [&](const float* base, int ld, int k0, float4 (&x)[NX]) {
#pragma unroll
    for (int it = 0; it < NX; ++it) {
      const int idx = it * 128 + tid, r = idx >> 2, c = idx & 3;
      x[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + r < T) x[it] = *reinterpret_cast<const float4*>(base + (r0 + k0 + r) * (int64_t)ld + col0 + 4 * c);
    }
  };
  auto store_k = [&](const float4 (&x)[NX]) {
#pragma unroll
    for (int it = 0; it < NX; ++it) {
      const int idx = it * 128 + tid, r = idx >> 2, c = idx & 3;
      const int o = r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
      *reinterpret_cast<float4*>(smem + AT_OFF_KH + o) = x[it];
      *reinterpret_cast<float4*>(smem + AT_OFF_KL + o) = tf32_lo(x[it]);
    }
  };
  // V^T: row d (0..15) of panel p holds keys 32p .. 32p+31 (128 bytes); SWIZZLE_128B: chunk q of row d sits at chunk q ^ (d & 7)
  auto store_v = [&](const float4 (&x)[NX]) {
#pragma unroll
    for (int it = 0; it < NX; ++it) {
      const int idx = it * 128 + tid, r = idx >> 2, c = idx & 3;
      const float xs[4] = {x[it].x, x[it].y, x[it].z, x[it].w};
      const int p = r >> 5, kk = r & 31;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = 4 * c + j;
        const int o = p * 2048 + d * 128 + (((kk >> 2) ^ (d & 7)) << 4) + (kk & 3) * 4;
        *reinterpret_cast<float*>(smem + AT_OFF_VH + o) = xs[j];
        *reinterpret_cast<float*>(smem + AT_OFF_VL + o) = tf32_lo(xs[j]);
      }
    }
  };
  // S[128 x n_s] = Q K^T for the staged key block (n_s = valid keys rounded up to 16); issued by one thread
  auto issue_s = [&](int n_s) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_s >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
    const uint64_t qh = make_smem_desc<16>(sbase + AT_OFF_QH), ql = make_smem_desc<16>(sbase + AT_OFF_QL);
    const uint64_t kh = make_smem_desc<16>(sbase + AT_OFF_KH), kl = make_smem_desc<16>(sbase + AT_OFF_KL);
    uint32_t acc = 0;
    if (products == 3) {
#pragma unroll
      for (int k = 0; k < 2; ++k) { umma_tf32(t_s, ql + 2 * k, kh + 2 * k, idesc, acc); acc = 1; }
#pragma unroll
      for (int k = 0; k < 2; ++k) umma_tf32(t_s, qh + 2 * k, kl + 2 * k, idesc, 1u);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) { umma_tf32(t_s, qh + 2 * k, kh + 2 * k, idesc, acc); acc = 1; }
    umma_commit(&mma_bar);
  };

  // ---------------- one pass over the key blocks: online softmax ----------------
  // m = running row maximum, l = running sum of exp(s - m); when a block raises the maximum, the output accumulated so far (TMEM) and l
  // are rescaled by exp(m_old - m_new) -- softmax(s) V exactly, without a separate pass over S for the maxima (which cost one more
  // K staging + MMA round trip per block).  The rescale is skipped warp-wide when no row of the warp changed its maximum.
  float m = -INFINITY, l = 0.f;
  float4 kreg[NX], vreg[NX];
  load_kv(K, ldk, 0, kreg);
  load_kv(V, ldv, 0, vreg);
  for (int b = 0; b < n_blocks; ++b) {
    const int k0 = b * AT_KC, valid = min(AT_KC, T - k0), n_s = (valid + 15) & ~15;
    store_k(kreg);
    store_v(vreg);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) { tc_fence_after(); issue_s(n_s); }
    if (b + 1 < n_blocks) { load_kv(K, ldk, k0 + AT_KC, kreg); load_kv(V, ldv, k0 + AT_KC, vreg); }   // in flight during this block
    mbar_wait(&mma_bar, phase); phase ^= 1;
    tc_fence_after();
    float bm = -INFINITY;
#pragma unroll
    for (int half = 0; half < NX; ++half) {
      if (half * 32 < n_s) {                                       // warp-uniform
        uint32_t r[32];
        tmem_ld32(t_s + lane_base + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (half * 32 + j < valid) bm = fmaxf(bm, __uint_as_float(r[j]));
      }
    }
    const float m_new = fmaxf(m, bm);
    if (b > 0 && __any_sync(0xffffffffu, m_new > m)) {             // warp-uniform: tcgen05.ld / .st are warp-collective
      const float alpha = exp2f((m - m_new) * scale_log2e);        // 1 for rows whose maximum did not move
      uint32_t o[16];
      tmem_ld16(t_o + lane_base, o);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
      tmem_st16(t_o + lane_base, o);
      tmem_st_wait();
      l *= alpha;
    }
    m = m_new;
    const float mscaled = m * scale_log2e;
#pragma unroll
    for (int half = 0; half < NX; ++half) {
      uint8_t* ph = smem + AT_OFF_PH + half * 16384 + row * 128;
      uint8_t* pl = smem + AT_OFF_PL + half * 16384 + row * 128;
      if (half * 32 < n_s) {                                       // warp-uniform
        uint32_t r[32];
        tmem_ld32(t_s + lane_base + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 p;
          p.x = (half * 32 + j + 0 < valid) ? exp2f(fmaf(__uint_as_float(r[j + 0]), scale_log2e, -mscaled)) : 0.f;
          p.y = (half * 32 + j + 1 < valid) ? exp2f(fmaf(__uint_as_float(r[j + 1]), scale_log2e, -mscaled)) : 0.f;
          p.z = (half * 32 + j + 2 < valid) ? exp2f(fmaf(__uint_as_float(r[j + 2]), scale_log2e, -mscaled)) : 0.f;
          p.w = (half * 32 + j + 3 < valid) ? exp2f(fmaf(__uint_as_float(r[j + 3]), scale_log2e, -mscaled)) : 0.f;
          l += (p.x + p.y) + (p.z + p.w);
          const int o = ((j >> 2) ^ (row & 7)) << 4;                // SWIZZLE_128B
          *reinterpret_cast<float4*>(ph + o) = p;
          *reinterpret_cast<float4*>(pl + o) = tf32_lo(p);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // O[128 x 16] += P[128 x ksteps*8] V[ksteps*8 x 16]; only the k steps that hold valid keys (the rest of P is stale / zero)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(AT_DH >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
      const int ksteps = (valid + 7) >> 3;
      uint32_t acc = b ? 1u : 0u;
      for (int ks = 0; ks < ksteps; ++ks) {
        const int p = ks >> 2, k = ks & 3;
        const uint64_t ph = make_smem_desc<32>(sbase + AT_OFF_PH + p * 16384) + 2 * k, pl = make_smem_desc<32>(sbase + AT_OFF_PL + p * 16384) + 2 * k;
        const uint64_t vh = make_smem_desc<32>(sbase + AT_OFF_VH + p * 2048) + 2 * k, vl = make_smem_desc<32>(sbase + AT_OFF_VL + p * 2048) + 2 * k;
        if (products == 3) {
          umma_tf32(t_o, pl, vh, idesc, acc); acc = 1;
          umma_tf32(t_o, ph, vl, idesc, 1u);
        }
        umma_tf32(t_o, ph, vh, idesc, acc); acc = 1;
      }
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, phase); phase ^= 1;                        // P / V / K may be overwritten, S recomputed, O rescaled
    tc_fence_after();
  }

  // ---------------- O / l -> global ----------------
  {
    uint32_t r[16];
    tmem_ld16(t_o + lane_base, r);
    tmem_ld_wait();
    if (q0 + row < T) {
      const float inv = 1.f / l;
      float* o = O + (r0 + q0 + row) * (int64_t)ldo + col0;
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv,
                                                        __uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, AT_TMEM_COLS); }
}


// ---------------------------------------------------------------------------------------------------------------------------------
// head_dim 64: the BIG-C decoder's self-attention over the Q queries of a video (models/model_0v10.py:181-186, nn.MultiheadAttention(512, 8))
// -- and any other ragged / fixed-length family of sequences with 64-wide heads -- as ONE kernel: S never leaves the SM (it used to be two
// batched GEMMs + a softmax pass + a V^T transpose through HBM: 2.5 of the 22 ms VidVRD step).
//
// Same anatomy as mha16_tc_kernel (CTA = (sequence, 128 queries, head), blocks of 64 keys, online softmax, accumulators in TMEM), but the
// operands are fp16 hi / lo PAIRS and every product is kind::f16 (K = 16 per instruction): x = hi + lo with hi = fp16_rn(x),
// lo = fp16_rn(x - hi), S = Ql Kh + Qh Kl + Qh Kh, O += Pl Vh + Ph Vl + Ph Vh -- 22 significant bits per operand wherever |x| >= 2^-2 and an
// absolute 2^-25 below (scores and probabilities are O(1) quantities, so that is fp32-class; measured <= 3e-6 against fp64).  `products` = 1
// keeps the hi parts only (11 bits, like tf32: the reduced-precision modes).  |q|, |k|, |v| must stay below the fp16 range (65504).
// Shared memory 96 KB, TMEM 128 columns -> 2 CTAs per SM.
constexpr int A6_DH = 64, A6_KC = 64;
constexpr int A6_QH = 0, A6_QL = 16384, A6_KH = 32768, A6_KL = 40960, A6_VH = 49152, A6_VL = 57344, A6_PH = 65536, A6_PL = 81920;
constexpr int A6_SMEM = 98304;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// float4 -> 4 fp16 (hi) and the fp16 of the remainders (lo), each packed into 8 bytes
__device__ __forceinline__ void split_f16(const float4 x, uint2& hi, uint2& lo) {
  const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(x.x - f0.x, x.y - f0.y), l1 = __floats2half2_rn(x.z - f1.x, x.w - f1.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

// 256 threads: warp w owns TMEM lane quadrant w & 3 (query rows 32 (w & 3) ..) and column half w >> 2 of S / P / O -- two threads per
// query row, which halves the exponential / conversion chain of a block (the kernel is bound by that chain and by the MMA round trips,
// not by the tensor pipe); the halves exchange their block maxima through shared memory.
// FULLK: every sequence length is a multiple of 64 keys (the decoder: Q = 192), so no key is ever masked -- the per-element `col < valid`
// selects, the per-load `key < T` predicates and the variable MMA shapes drop out of the instruction stream (the kernel is instruction-bound).
template <bool FULLK>
__global__ void __launch_bounds__(256, 2)
mha64_tc_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
                const int64_t* __restrict__ seg_off, int fixed_len, const int32_t* __restrict__ blk_seg, const int32_t* __restrict__ blk_q0,
                float scale_log2e, int products, float* __restrict__ O, int ldo) {
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float xch[2][AT_BM];                                  // block maxima / final sums of the two column halves
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  int seg, q0, T;
  int64_t r0;
  if (seg_off) {                                                   // ragged sequences: work list of (sequence, first query) blocks
    seg = blk_seg[blockIdx.x]; q0 = blk_q0[blockIdx.x];
    if (q0 & (AT_BM - 1)) return;                                  // entries of a finer (64-query) work list that do not start a 128-query block
    r0 = seg_off[seg]; T = (int)(seg_off[seg + 1] - r0);
  } else {                                                         // sequences of `fixed_len` rows, back to back
    const int per = (fixed_len + AT_BM - 1) / AT_BM;
    seg = blockIdx.x / per; q0 = (blockIdx.x - seg * per) * AT_BM;
    r0 = (int64_t)seg * fixed_len; T = fixed_len;
  }
  const int col0 = blockIdx.y * A6_DH;
  const bool split = products == 3;

  if (tid == 0) { mbar_init(&mma_bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 128); tmem_relinquish(); }
  // ---- Q tile: row r = query q0 + r, 128-byte rows of 64 fp16; SWIZZLE_128B: 16-byte chunk c of row r sits at chunk c ^ (r & 7).
  //      Item = 4 floats -> 8 bytes: chunk c4 >> 1, half (c4 & 1) ----
#pragma unroll 4
  for (int idx = tid; idx < AT_BM * 16; idx += 256) {
    const int r = idx >> 4, c4 = idx & 15;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) x = *reinterpret_cast<const float4*>(Q + (r0 + q0 + r) * (int64_t)ldq + col0 + 4 * c4);
    uint2 hi, lo;
    split_f16(x, hi, lo);
    const int o = r * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + (c4 & 1) * 8;
    *reinterpret_cast<uint2*>(smem + A6_QH + o) = hi;
    if (split) *reinterpret_cast<uint2*>(smem + A6_QL + o) = lo;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_s = tmem + half * 32, t_o = tmem + A6_KC + half * 32;         // this thread's 32 columns of S and of O
  const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
  const uint32_t sbase = smem_u32(smem);
  const int row = quad * 32 + lane;
  const bool live = q0 + quad * 32 < T;                            // warp-uniform: some row of this warp is a real query
  uint32_t phase = 0;
  const int n_blocks = (T + A6_KC - 1) / A6_KC;

  // K block: 64 keys x 16 float4 items = 4 per thread (item idx = it * 256 + tid: key idx >> 4, chunk idx & 15), coalesced.
  // V block: TRANSPOSED into [dim][key] rows; a thread takes the same 4 dims of a PAIR of keys (2j, 2j + 1), so that every store is one
  // 32-bit word and a warp's 32 key pairs fill one whole 128-byte row: 32 pairs x 16 chunks = 2 pair items per thread.
  auto load_k = [&](int k0, float4 (&x)[4]) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 256 + tid, r = idx >> 4, c4 = idx & 15;
      x[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (FULLK || k0 + r < T) x[it] = *reinterpret_cast<const float4*>(K + (r0 + k0 + r) * (int64_t)ldk + col0 + 4 * c4);
    }
  };
  auto load_v = [&](int k0, float4 (&x)[4]) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int c4 = warp + 8 * it;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        x[2 * it + e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (FULLK || k0 + 2 * lane + e < T) x[2 * it + e] = *reinterpret_cast<const float4*>(V + (r0 + k0 + 2 * lane + e) * (int64_t)ldv + col0 + 4 * c4);
      }
    }
  };
  auto store_k = [&](const float4 (&x)[4]) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 256 + tid, r = idx >> 4, c4 = idx & 15;
      uint2 hi, lo;
      split_f16(x[it], hi, lo);
      const int o = r * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + (c4 & 1) * 8;
      *reinterpret_cast<uint2*>(smem + A6_KH + o) = hi;
      if (split) *reinterpret_cast<uint2*>(smem + A6_KL + o) = lo;
    }
  };
  auto store_v = [&](const float4 (&x)[4]) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int c4 = warp + 8 * it;
      const float a[4] = {x[2 * it].x, x[2 * it].y, x[2 * it].z, x[2 * it].w};
      const float b[4] = {x[2 * it + 1].x, x[2 * it + 1].y, x[2 * it + 1].z, x[2 * it + 1].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 4 * c4 + e;
        const __half2 h = __floats2half2_rn(a[e], b[e]);
        const float2 f = __half22float2(h);
        const int o = d * 128 + ((((2 * lane) >> 3) ^ (d & 7)) << 4) + ((2 * lane) & 7) * 2;
        *reinterpret_cast<uint32_t*>(smem + A6_VH + o) = *reinterpret_cast<const uint32_t*>(&h);
        if (split) {
          const __half2 l = __floats2half2_rn(a[e] - f.x, b[e] - f.y);
          *reinterpret_cast<uint32_t*>(smem + A6_VL + o) = *reinterpret_cast<const uint32_t*>(&l);
        }
      }
    }
  };
  // S[128 x n_s] = Q K^T over the 64 dims (4 K = 16 steps per product)
  auto issue_s = [&](int n_s) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n_s >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);      // kind::f16, fp16 A / B, fp32 D
    const uint64_t qh = make_smem_desc<32>(sbase + A6_QH), ql = make_smem_desc<32>(sbase + A6_QL);
    const uint64_t kh = make_smem_desc<32>(sbase + A6_KH), kl = make_smem_desc<32>(sbase + A6_KL);
    uint32_t acc = 0;
    if (split) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { umma_bf16(tmem, ql + 2 * k, kh + 2 * k, idesc, acc); acc = 1; }
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, qh + 2 * k, kl + 2 * k, idesc, 1u);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { umma_bf16(tmem, qh + 2 * k, kh + 2 * k, idesc, acc); acc = 1; }
    umma_commit(&mma_bar);
  };

  float m = -INFINITY, l = 0.f;                                    // l: this thread's column half only
  float4 kreg[4], vreg[4];
  load_k(0, kreg);
  load_v(0, vreg);
  for (int b = 0; b < n_blocks; ++b) {
    const int k0 = b * A6_KC, valid = FULLK ? A6_KC : min(A6_KC, T - k0), n_s = FULLK ? A6_KC : ((valid + 15) & ~15);
    store_k(kreg);
    store_v(vreg);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) { tc_fence_after(); issue_s(n_s); }
    if (b + 1 < n_blocks) { load_k(k0 + A6_KC, kreg); load_v(k0 + A6_KC, vreg); }      // in flight during this block
    mbar_wait(&mma_bar, phase); phase ^= 1;
    tc_fence_after();
    // this thread's 32 scores of the block, read once
    const bool cols = live && half * 32 < n_s;                     // warp-uniform
    uint32_t sr[32];
    float bm = -INFINITY;
    if (cols) {
      tmem_ld32(t_s + lane_base, sr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (FULLK || half * 32 + j < valid) bm = fmaxf(bm, __uint_as_float(sr[j]));
    }
    xch[half][row] = bm;
    __syncthreads();
    const float m_new = fmaxf(m, fmaxf(xch[0][row], xch[1][row]));
    if (live && b > 0 && __any_sync(0xffffffffu, m_new > m)) {     // warp-uniform: tcgen05.ld / .st are warp-collective
      const float alpha = exp2f((m - m_new) * scale_log2e);
      uint32_t o[32];
      tmem_ld32(t_o + lane_base, o);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
      tmem_st32(t_o + lane_base, o);
      tmem_st_wait();
      l *= alpha;
    }
    m = m_new;
    if (cols) {
      const float mscaled = m * scale_log2e;
      uint8_t* ph = smem + A6_PH + row * 128;
      uint8_t* pl = smem + A6_PL + row * 128;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {                             // 8 probabilities = one 16-byte chunk of fp16
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pv[e] = (FULLK || half * 32 + cc * 8 + e < valid) ? exp2f(fmaf(__uint_as_float(sr[cc * 8 + e]), scale_log2e, -mscaled)) : 0.f;
          l += pv[e];
        }
        uint2 h0, l0, h1, l1;
        split_f16(make_float4(pv[0], pv[1], pv[2], pv[3]), h0, l0);
        split_f16(make_float4(pv[4], pv[5], pv[6], pv[7]), h1, l1);
        const int o = (((half * 4 + cc) ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(ph + o) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        if (split) *reinterpret_cast<uint4*>(pl + o) = make_uint4(l0.x, l0.y, l1.x, l1.y);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // O[128 x 64] += P[128 x n_s] V[n_s x 64]: K = 16 keys per instruction; P is zero for the invalid keys inside the last step
      const uint32_t idesc = (1u << 4) | ((uint32_t)(A6_DH >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
      const uint64_t pH = make_smem_desc<32>(sbase + A6_PH), pL = make_smem_desc<32>(sbase + A6_PL);
      const uint64_t vH = make_smem_desc<32>(sbase + A6_VH), vL = make_smem_desc<32>(sbase + A6_VL);
      const int ksteps = FULLK ? 4 : (n_s >> 4);
      uint32_t acc = b ? 1u : 0u;
#pragma unroll
      for (int ks = 0; ks < (FULLK ? 4 : ksteps); ++ks) {
        if (split) {
          umma_bf16(tmem + A6_KC, pL + 2 * ks, vH + 2 * ks, idesc, acc); acc = 1;
          umma_bf16(tmem + A6_KC, pH + 2 * ks, vL + 2 * ks, idesc, 1u);
        }
        umma_bf16(tmem + A6_KC, pH + 2 * ks, vH + 2 * ks, idesc, acc); acc = 1;
      }
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, phase); phase ^= 1;                        // P / V / K may be overwritten, S recomputed, O rescaled
    tc_fence_after();
  }

  // ---------------- O / l -> global ----------------
  xch[half][row] = l;
  __syncthreads();
  if (live) {
    const float inv = 1.f / (xch[0][row] + xch[1][row]);
    uint32_t r[32];
    tmem_ld32(t_o + lane_base, r);
    tmem_ld_wait();
    if (q0 + row < T) {
      float* o = O + (r0 + q0 + row) * (int64_t)ldo + col0 + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv,
                                                        __uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

}  // namespace vsg

using namespace vsg;

static int g_at_kc = 32;
/* validation / tuning knob: keys per block of the tcgen05 attention kernel (32 or 64); returns the old value */
extern "C" int vsg_mha_tc16_set_kc(int kc) { int old = g_at_kc; if (kc == 32 || kc == 64) g_at_kc = kc; return old; }

template <int KC>
static int launch_mha16(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const int64_t* seg_off, int n_head, float* O,
                        int ldo, const int32_t* blk_seg, const int32_t* blk_q0, int n_blk, int products, void* stream) {
  static PerDeviceFlag attr_done;
  const int dev_ = current_device();
  constexpr int SMEM = AtCfg<KC>::SMEM + 1024;
  if (!attr_done.is_set(dev_)) {
    if (cudaFuncSetAttribute(mha16_tc_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
      set_error("vsg_mha_tc16: cannot raise dynamic shared memory to %d", SMEM);
      return VSG_E_LAUNCH;
    }
    attr_done.set(dev_);
  }
  const float scale_log2e = 1.4426950408889634f / 4.0f;            // 1 / sqrt(16) * log2(e)
  mha16_tc_kernel<KC><<<dim3(n_blk, n_head), 128, SMEM, (cudaStream_t)stream>>>(Q, ldq, K, ldk, V, ldv, seg_off, blk_seg, blk_q0, scale_log2e,
                                                                                products, O, ldo);
  return check_launch("vsg_mha_tc16");
}

extern "C" int vsg_mha_tc16(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const int64_t* seg_off,
                            int n_head, float* O, int ldo, const int32_t* blk_seg, const int32_t* blk_q0, int n_blk, int products,
                            void* stream) {
  VSG_REQUIRE(n_blk >= 0 && n_head > 0 && n_head <= 65535, "vsg_mha_tc16: bad size");
  if (n_blk == 0) return VSG_OK;
  VSG_REQUIRE(Q && K && V && O && seg_off && blk_seg && blk_q0, "vsg_mha_tc16: null pointer");
  VSG_REQUIRE(products == 1 || products == 3, "vsg_mha_tc16: products must be 1 (tf32) or 3 (3xTF32)");
  VSG_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(O) && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0,
              "vsg_mha_tc16: Q / K / V / O must be 16-byte aligned with leading dimensions that are multiples of 4");
  return g_at_kc == 64 ? launch_mha16<64>(Q, ldq, K, ldk, V, ldv, seg_off, n_head, O, ldo, blk_seg, blk_q0, n_blk, products, stream)
                       : launch_mha16<32>(Q, ldq, K, ldk, V, ldv, seg_off, n_head, O, ldo, blk_seg, blk_q0, n_blk, products, stream);
}

extern "C" int vsg_mha_tc64(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const int64_t* seg_off, int n_seg,
                            int fixed_len, int n_head, float* O, int ldo, const int32_t* blk_seg, const int32_t* blk_q0, int n_blk, int products,
                            void* stream) {
  VSG_REQUIRE(n_head > 0 && n_head <= 65535 && n_seg >= 0, "vsg_mha_tc64: bad size");
  VSG_REQUIRE(products == 1 || products == 3, "vsg_mha_tc64: products must be 1 (hi parts only) or 3 (fp16 hi / lo pairs)");
  long long blocks;
  if (seg_off) {
    VSG_REQUIRE(n_blk >= 0 && (n_blk == 0 || (blk_seg && blk_q0)), "vsg_mha_tc64: ragged sequences need the (sequence, first query) work list");
    blocks = n_blk;
  } else {
    VSG_REQUIRE(fixed_len > 0, "vsg_mha_tc64: seg_off == NULL needs fixed_len > 0");
    blocks = (long long)n_seg * ((fixed_len + AT_BM - 1) / AT_BM);
  }
  if (blocks == 0) return VSG_OK;
  VSG_REQUIRE(blocks <= 0x7fffffffLL, "vsg_mha_tc64: too many blocks");
  VSG_REQUIRE(Q && K && V && O, "vsg_mha_tc64: null pointer");
  VSG_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(O) && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0,
              "vsg_mha_tc64: Q / K / V / O must be 16-byte aligned with leading dimensions that are multiples of 4");
  static PerDeviceFlag attr_done;
  const int dev_ = current_device();
  constexpr int SMEM = A6_SMEM + 1024;
  if (!attr_done.is_set(dev_)) {
    if (cudaFuncSetAttribute(mha64_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(mha64_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
      set_error("vsg_mha_tc64: cannot raise dynamic shared memory to %d", SMEM);
      return VSG_E_LAUNCH;
    }
    attr_done.set(dev_);
  }
  const float scale_log2e = 1.4426950408889634f / 8.0f;            // 1 / sqrt(64) * log2(e)
  if (!seg_off && fixed_len % A6_KC == 0)
    mha64_tc_kernel<true><<<dim3((unsigned)blocks, n_head), 256, SMEM, (cudaStream_t)stream>>>(Q, ldq, K, ldk, V, ldv, seg_off, fixed_len, blk_seg, blk_q0,
                                                                                               scale_log2e, products, O, ldo);
  else
    mha64_tc_kernel<false><<<dim3((unsigned)blocks, n_head), 256, SMEM, (cudaStream_t)stream>>>(Q, ldq, K, ldk, V, ldv, seg_off, fixed_len, blk_seg, blk_q0,
                                                                                                scale_log2e, products, O, ldo);
  return check_launch("vsg_mha_tc64");
}
