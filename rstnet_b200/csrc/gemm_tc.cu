// tcgen05 / TMEM / TMA GEMM for the codec's convs, transposed convs and linears.
//
//   D[(i,o), n] = sum_{tap} sum_{c} A[c, i + tap*tap_di, o*o_mul + tap*tap_do] * W[n, tap*Kc + c]
//
// A is a 3-D TMA tensor (c, i, o): the activation buffer itself (no im2col), so a causal conv is a
// K loop over taps.  One CTA owns a 128 x BN output tile with the fp32 accumulator in TMEM.
// Warp roles (448 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// tcgen05.mma issuer, warps 2..9 = activation transform (pre-activation, TF32 hi/lo split of the A
// tile in shared memory; weights arrive pre-split), warps 10..13 = promote the TMEM chunk
// accumulators into fp32 registers, then run the epilogue.
//
// Precision 0 ("3xTF32"): fp32-equivalent products for the encoder side, where RVQ indices must
// match the fp32 reference.  Each fp32 operand x is split in shared memory into
// hi = rna_tf32(x) and lo = rna_tf32(x - hi) (both exactly representable in TF32), and three TF32
// MMAs accumulate a_lo*b_hi + a_hi*b_lo + a_hi*b_hi in fp32 (dropping the 2^-22 lo*lo term).
// Precision 1: one TF32 MMA (hi only).  The transform warps also apply a fused pre-activation (ELU).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/rstnet_b200.h"

namespace rstnet {
extern void count_launch();
using namespace tc;

constexpr int TC_BM = 128;
constexpr int TC_BKE = 32;                 // fp32 elements per 128-byte swizzle row
constexpr int TC_A_BYTES = TC_BM * 128;    // 16 KB
constexpr int TC_THREADS = 448;          // TMA, MMA, 8 transform, 4 drain/epilogue warps
constexpr int TC_XF_THREADS = 256;       // transform threads (warps 2..9)
constexpr int TC_CHUNK_STAGES = 4;       // K elements accumulated in TMEM before promotion = 4 * 32
constexpr int TC_NACC = 2;               // independent TMEM accumulators per buffer (consecutive MMAs alternate: no RAW chain)

struct TcParams {
  float* C;
  float* C2;  // optional second output: act2(pre-activation value), same strides as C
  int act2;
  long long c_i_stride, c_o_stride, c_split_stride;
  const float* R;
  long long r_i_stride, r_o_stride, r_split_stride;
  const float* bias;
  const float* scale;
  int n_split;
  int I_out, O_out, N, Kc;
  int taps, tap_di, tap_do, o_mul, kchunks;
  int pre_act, post_act, i_tiles;
  long long* trace;  // optional [total_k][8] clock64 stamps of CTA 0 (debug / profiling)
  long long* cta_times;  // optional [grid.x][4] %globaltimer: entry, setup done, mainloop+epilogue done, exit (blockIdx.y == 0)
};

template <int BN, int PREC>
struct TcCfg {
  static constexpr bool SPLIT = PREC == 0;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = (TC_A_BYTES + B_BYTES) * (SPLIT ? 2 : 1);
  static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * TC_NACC * BN;  // two buffers x TC_NACC interleaved accumulators
};

// round-to-nearest (ties away) TF32 on the integer pipe: add half an ulp of the 10-bit mantissa to
// the magnitude bits and clear the low 13 bits (cvt.rna.tf32.f32 gives the same value but runs at
// the slow conversion rate).  |x - hi| <= 2^-12 |x|, so with lo = rna(x - hi) the split
// x ~ hi + lo is good to 2^-24 |x|: fp32-equivalent products.
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float4 tf32_hi(float4 v) {
  return make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
}

template <int BN, int PREC>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmWlo, const TcParams p) {
  using Cfg = TcCfg<BN, PREC>;
  constexpr int S = Cfg::STAGES;
  constexpr int CH = TC_CHUNK_STAGES;
  extern __shared__ uint8_t smem_raw[];
  auto gtime = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
  const bool timed = p.cta_times && blockIdx.y == 0 && threadIdx.x == 0;
  if (timed) p.cta_times[blockIdx.x * 4 + 0] = gtime();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES);
  uint64_t* full = bars;               // [S] TMA landed
  uint64_t* ready = bars + S;          // [S] operands transformed (128 arrivals)
  uint64_t* empty = bars + 2 * S;      // [S] MMAs reading the stage retired
  uint64_t* acc_full = bars + 3 * S;   // [2] a K chunk has been accumulated into TMEM buffer b
  uint64_t* acc_empty = bars + 3 * S + 2;  // [2] buffer b drained into registers (128 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int it = blockIdx.x % p.i_tiles, ot = blockIdx.x / p.i_tiles;
  const int i0 = it * TC_BM, n0 = blockIdx.y * BN;
  const int total_k = p.taps * p.kchunks;
  const int nchunks = (total_k + CH - 1) / CH;
  // single-pass TF32 without a pre-activation needs no operand transform: the tensor core reads the top 19 bits
  // of the fp32 words TMA delivered (truncation), so the MMA warp consumes the TMA barrier directly
  const bool xf = Cfg::SPLIT || p.pre_act != ACT_NONE;

  auto a_hi = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto b_hi = [&](int s) { return smem + s * Cfg::STAGE_BYTES + TC_A_BYTES; };
  auto a_lo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + TC_A_BYTES + Cfg::B_BYTES; };
  auto b_lo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * TC_A_BYTES + Cfg::B_BYTES; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], TC_XF_THREADS);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (timed) p.cta_times[blockIdx.x * 4 + 1] = gtime();

  if (warp == 0) {
    // ================= TMA producer (whole warp runs the loop; one elected lane issues)
    for (int kit = 0; kit < total_k; ++kit) {
      const int s = kit % S;
      const uint32_t ph = (kit / S) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      const int tap = kit / p.kchunks, kc = kit % p.kchunks;
      if (elect_one()) {
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[kit * 8 + 0] = clock64();
        mbar_arrive_expect_tx(&full[s], TC_A_BYTES + Cfg::B_BYTES * (Cfg::SPLIT ? 2 : 1));
        tma_load_3d(a_hi(s), &tmA, &full[s], kc * TC_BKE, i0 + tap * p.tap_di, ot * p.o_mul + tap * p.tap_do);
        tma_load_2d(b_hi(s), &tmW, &full[s], tap * p.Kc + kc * TC_BKE, n0);
        if (Cfg::SPLIT) tma_load_2d(b_lo(s), &tmWlo, &full[s], tap * p.Kc + kc * TC_BKE, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread).  The tensor core adds into its fp32 accumulator with
    // truncation, a bias that grows linearly with the number of accumulation steps; each chunk of CH
    // stages therefore starts a fresh TMEM accumulator (two buffers, ping-pong) that the drain warps add
    // into fp32 registers with round-to-nearest (same idea as FP8 GEMM accumulator promotion).
    constexpr uint32_t idesc = instr_desc(2u, TC_BM, BN);
    for (int kit = 0; kit < total_k; ++kit) {
      const int s = kit % S;
      const uint32_t ph = (kit / S) & 1;
      const int chunk = kit / CH, pos = kit % CH, buf = chunk & 1;
      if (pos == 0) mbar_wait(&acc_empty[buf], ((chunk >> 1) & 1) ^ 1);
      mbar_wait(xf ? &ready[s] : &full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[kit * 8 + 3] = clock64();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * TC_NACC * BN);  // accumulator 0 of this buffer
        const uint32_t d1 = d0 + (uint32_t)BN;                            // accumulator 1
        const uint64_t da = smem_desc_sw128(smem_u32(a_hi(s))), db = smem_desc_sw128(smem_u32(b_hi(s)));
        uint64_t dal = 0, dbl = 0;
        if (Cfg::SPLIT) {
          dal = smem_desc_sw128(smem_u32(a_lo(s)));
          dbl = smem_desc_sw128(smem_u32(b_lo(s)));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x (K = 8 tf32 = 32 bytes) per 128-byte row
          const uint64_t adv = (uint64_t)(k * 2);  // +32 bytes in 16-byte units
          const uint32_t fresh = (pos == 0 && k == 0) ? 0u : 1u;
          if (Cfg::SPLIT) {
            mma_tf32(d0, da + adv, db + adv, idesc, fresh);    // a_hi * b_hi
            mma_tf32(d1, dal + adv, db + adv, idesc, fresh);   // a_lo * b_hi   (small terms kept apart)
            mma_tf32(d1, da + adv, dbl + adv, idesc, 1u);      // a_hi * b_lo
          } else {
            mma_tf32((k & 1) ? d1 : d0, da + adv, db + adv, idesc, (pos == 0 && k < 2) ? 0u : 1u);
          }
        }
        tc_commit(&empty[s]);  // frees the stage once these MMAs have read it
        if (pos == CH - 1 || kit == total_k - 1) tc_commit(&acc_full[buf]);
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[kit * 8 + 4] = clock64();
      }
      __syncwarp();
    }
  } else if (warp < 10) {
    // ================= transform warps (2..9): pre-activation + TF32 hi/lo split of the A tile, in place.
    // Weights were rounded / split once on the host side of the plan (W_hi, W_lo arrive by TMA).
    const int tid = threadIdx.x - 64;  // 0..255
    for (int kit = 0; xf && kit < total_k; ++kit) {
      const int s = kit % S;
      const uint32_t ph = (kit / S) & 1;
      mbar_wait(&full[s], ph);
      if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) p.trace[kit * 8 + 1] = clock64();
      float4* ah = reinterpret_cast<float4*>(a_hi(s));
      float4* al = reinterpret_cast<float4*>(a_lo(s));
#pragma unroll
      for (int j = 0; j < TC_A_BYTES / 16 / TC_XF_THREADS; ++j) {
        float4 v = ah[j * TC_XF_THREADS + tid];
        if (p.pre_act != ACT_NONE) {
          v.x = apply_act(v.x, p.pre_act); v.y = apply_act(v.y, p.pre_act);
          v.z = apply_act(v.z, p.pre_act); v.w = apply_act(v.w, p.pre_act);
        }
        const float4 h = tf32_hi(v);
        ah[j * TC_XF_THREADS + tid] = h;
        if (Cfg::SPLIT) al[j * TC_XF_THREADS + tid] = tf32_hi(make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
      }
      fence_proxy_async_smem();
      if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) p.trace[kit * 8 + 2] = clock64();
      mbar_arrive(&ready[s]);
    }
  } else {
    // ================= drain + epilogue warps (10..13): TMEM chunk accumulators -> fp32 registers
    const int q = warp % 4;  // TMEM lane quarter this warp may read
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int buf = chunk & 1;
      mbar_wait(&acc_full[buf], (chunk >> 1) & 1);
      tc_fence_after();
      if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 320) p.trace[chunk * 8 + 5] = clock64();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        uint32_t r2[32];
        const uint32_t col = (uint32_t)(buf * TC_NACC * BN + c0);
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + col, r);
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + col + (uint32_t)BN, r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]) + __uint_as_float(r2[j]);
      }
      tc_fence_before();
      if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 320) p.trace[chunk * 8 + 6] = clock64();
      mbar_arrive(&acc_empty[buf]);
    }
    // ---- epilogue: bias, LayerScale, residual, activation -> global
    const int row = q * 32 + lane;
    const int i = i0 + row;
    if (i < p.I_out) {
      float* crow = p.C + (long long)ot * p.c_o_stride + (long long)i * p.c_i_stride;
      float* crow2 = p.C2 ? p.C2 + (long long)ot * p.c_o_stride + (long long)i * p.c_i_stride : nullptr;
      const float* rrow = p.R ? p.R + (long long)ot * p.r_o_stride + (long long)i * p.r_i_stride : nullptr;
#pragma unroll
      for (int g = 0; g < BN / 4; ++g) {
        const int n = n0 + 4 * g;
        if (n < p.N) {
          float4 v = make_float4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
          if (p.bias) {
            const float4 bb = *reinterpret_cast<const float4*>(p.bias + n);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          }
          if (p.scale) {
            const float4 ss = *reinterpret_cast<const float4*>(p.scale + n);
            v.x *= ss.x; v.y *= ss.y; v.z *= ss.z; v.w *= ss.w;
          }
          long long coff = n, roff = n;
          if (p.n_split > 0) {
            const int j = n / p.n_split, co = n % p.n_split;
            coff = (long long)j * p.c_split_stride + co;
            roff = (long long)j * p.r_split_stride + co;
          }
          if (rrow) {
            const float4 rr = *reinterpret_cast<const float4*>(rrow + roff);
            v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
          }
          if (crow2) {  // e.g. the ELU'd copy a resblock's first conv consumes while the skip reads the raw one
            *reinterpret_cast<float4*>(crow2 + coff) = make_float4(apply_act(v.x, p.act2), apply_act(v.y, p.act2),
                                                                   apply_act(v.z, p.act2), apply_act(v.w, p.act2));
          }
          if (p.post_act != ACT_NONE) {
            v.x = apply_act(v.x, p.post_act); v.y = apply_act(v.y, p.post_act);
            v.z = apply_act(v.z, p.post_act); v.w = apply_act(v.w, p.post_act);
          }
          *reinterpret_cast<float4*>(crow + coff) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (timed) p.cta_times[blockIdx.x * 4 + 2] = gtime();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  if (timed) p.cta_times[blockIdx.x * 4 + 3] = gtime();
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace rstnet
using namespace rstnet;

struct rstnet_tc_plan {
  CUtensorMap tmA, tmW, tmWlo;
  TcParams p;
  dim3 grid;
  int bn, prec;
};

template <int BN, int PREC>
static int tc_launch(const rstnet_tc_plan* pl, cudaStream_t st) {
  using Cfg = TcCfg<BN, PREC>;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gemm_tc_kernel<BN, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    attr = true;
  }
  gemm_tc_kernel<BN, PREC><<<pl->grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(pl->tmA, pl->tmW, pl->tmWlo, pl->p);
  count_launch();
  return check_launch("gemm_tc");
}

extern "C" int rstnet_tc_gemm_create(const rstnet_tc_gemm_desc* d, rstnet_tc_plan** out) {
  RSTNET_REQUIRE(d && out, "tc_gemm_create: null argument");
  RSTNET_REQUIRE(d->A && d->W && d->C, "tc_gemm_create: null tensor pointer");
  RSTNET_REQUIRE(d->precision != 0 || d->W_lo, "tc_gemm_create: precision 0 (3xTF32) needs W_lo");
  RSTNET_REQUIRE(d->Kc > 0 && d->Kc % TC_BKE == 0, "tc_gemm_create: Kc (%d) must be a multiple of %d", d->Kc, TC_BKE);
  RSTNET_REQUIRE(d->N > 0 && d->N % 4 == 0, "tc_gemm_create: N (%d) must be a multiple of 4", d->N);
  RSTNET_REQUIRE(d->taps >= 1 && d->I_out > 0 && d->O_out > 0, "tc_gemm_create: bad shape");
  RSTNET_REQUIRE(d->precision == 0 || d->precision == 1, "tc_gemm_create: precision must be 0 (3xTF32) or 1 (TF32)");
  RSTNET_REQUIRE((uintptr_t)d->A % 16 == 0 && (uintptr_t)d->W % 16 == 0 && (uintptr_t)d->C % 16 == 0 &&
                     d->a_i_stride % 4 == 0 && d->a_o_stride % 4 == 0 && d->c_i_stride % 4 == 0 && d->c_o_stride % 4 == 0 &&
                     d->c_split_stride % 4 == 0 && d->n_split % 4 == 0,
                 "tc_gemm_create: 16-byte alignment required");
  EncodeTiledFn enc = get_encode_fn();
  RSTNET_REQUIRE(enc != nullptr, "tc_gemm_create: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  rstnet_tc_plan* pl = new rstnet_tc_plan();
  const int N = d->N;
  pl->bn = N >= 64 ? 64 : 32;
  // narrow the tile when the grid would leave most of the 148 SMs idle
  const int i_tiles = ceil_div(d->I_out, TC_BM);
  while (pl->bn > 32 && (long long)i_tiles * d->O_out * ceil_div(N, pl->bn) < 148) pl->bn /= 2;
  pl->prec = d->precision;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)d->a_c_extent, (cuuint64_t)d->a_i_extent, (cuuint64_t)d->a_o_extent};
    cuuint64_t gstr[2] = {(cuuint64_t)d->a_i_stride * 4, (cuuint64_t)d->a_o_stride * 4};
    cuuint32_t box[3] = {TC_BKE, TC_BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)d->A, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      set_error("tc_gemm_create: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
      return 3;
    }
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)d->taps * d->Kc, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)d->taps * d->Kc * 4};
    cuuint32_t box[2] = {TC_BKE, (cuuint32_t)pl->bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&pl->tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)d->W, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS && d->W_lo)
      r = enc(&pl->tmWlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)d->W_lo, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    else
      pl->tmWlo = pl->tmW;
    if (r != CUDA_SUCCESS) {
      delete pl;
      set_error("tc_gemm_create: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
      return 3;
    }
  }
  TcParams& p = pl->p;
  p.C2 = d->C2; p.act2 = d->act2;
  p.C = d->C; p.c_i_stride = d->c_i_stride; p.c_o_stride = d->c_o_stride; p.c_split_stride = d->c_split_stride;
  p.R = d->R; p.r_i_stride = d->r_i_stride; p.r_o_stride = d->r_o_stride; p.r_split_stride = d->r_split_stride;
  p.bias = d->bias; p.scale = d->scale; p.n_split = d->n_split;
  p.I_out = d->I_out; p.O_out = d->O_out; p.N = N; p.Kc = d->Kc;
  p.taps = d->taps; p.tap_di = d->tap_di; p.tap_do = d->tap_do; p.o_mul = d->o_mul; p.kchunks = d->Kc / TC_BKE;
  p.pre_act = d->pre_act; p.post_act = d->post_act; p.i_tiles = i_tiles;
  p.trace = nullptr;
  p.cta_times = nullptr;
  pl->grid = dim3((unsigned)(i_tiles * d->O_out), (unsigned)ceil_div(N, pl->bn));
  *out = pl;
  return 0;
}

extern "C" int rstnet_tc_gemm_run(const rstnet_tc_plan* pl, rstnet_stream_t stream) {
  RSTNET_REQUIRE(pl != nullptr, "tc_gemm_run: null plan");
  cudaStream_t st = (cudaStream_t)stream;
  if (pl->prec == 0) {
    if (pl->bn == 64) return tc_launch<64, 0>(pl, st);
    return tc_launch<32, 0>(pl, st);
  }
  if (pl->bn == 64) return tc_launch<64, 1>(pl, st);
  return tc_launch<32, 1>(pl, st);
}

extern "C" void rstnet_tc_gemm_destroy(rstnet_tc_plan* pl) { delete pl; }
/* debug: CTA (0,0) records clock64 stamps per k iteration into trace[total_k][8] (device int64) */
extern "C" void rstnet_tc_gemm_set_trace(rstnet_tc_plan* pl, int64_t* trace, int64_t* cta_times) {
  if (pl) { pl->p.trace = (long long*)trace; pl->p.cta_times = (long long*)cta_times; }
}
extern "C" int rstnet_tc_gemm_grid(const rstnet_tc_plan* pl, int32_t* gx, int32_t* gy, int32_t* bn) {
  if (!pl) return 1;
  *gx = (int)pl->grid.x; *gy = (int)pl->grid.y; *bn = pl->bn;
  return 0;
}

namespace rstnet {
__global__ void tf32_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i], h = tf32_rna(v);
    hi[i] = h;
    lo[i] = tf32_rna(v - h);
  }
}
}  // namespace rstnet

extern "C" int rstnet_tf32_split_f32(const float* x, float* hi, float* lo, int64_t n, rstnet_stream_t stream) {
  RSTNET_REQUIRE(x && hi && lo, "tf32_split: null pointer");
  if (n <= 0) return 0;
  int g = ceil_div(n, 256);
  if (g > 148 * 8) g = 148 * 8;
  rstnet::tf32_split_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n);
  count_launch();
  return check_launch("tf32_split");
}
