// Device-side token sampling shared by the stand-alone sampler (lm_small.cu) and the persistent depth-transformer kernel
// (lm_depth_frame.cu).  Included inside namespace rstnet after `bf16`, b2f / f2b are defined.
#pragma once

// ---------------------------------------------------------------- sampling (utils/sampling.py:85-154)
__device__ __forceinline__ uint32_t hash_u32(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du ^ (d * 0x27D4EB2Fu);
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return h;
}

// logits [rows][V] bf16, candidates restricted to ids < n_valid.  top_k <= 0: argmax (first maximum).
// top_k > 0: the top_k largest logits, weights exp((l - max)/temp), token = argmax_i w_i / Exp(1)_i
// (= torch's exponential-noise multinomial over the top-k probabilities, sampling.py:43-46, 57-59).
//
// Large vocabularies (the 152k text head) first shrink the row to a candidate list: bf16 has 16 key bits, so two
// 256-bin histogram passes give the exact key of the top_k-th largest logit; every logit with key >= that threshold
// (top_k of them plus ties) is compacted into shared memory and the ordered selection below runs on the list instead
// of re-scanning the row top_k times.  Same result as the full scan: (value desc, index asc) order.
constexpr int SAMPLE_CAND = 1024;
constexpr int SAMPLE_HISTS = 16;

__device__ __forceinline__ uint32_t bf16_order_key(bf16 v) {
  const uint32_t u = (uint32_t)__bfloat16_as_ushort(v);
  return (u & 0x8000u) ? (~u & 0xFFFFu) : (u | 0x8000u);
}

// block-wide argmax in the order (value desc, index asc); every thread returns the winner
__device__ __forceinline__ void block_argmax(float& bv, int& bi, float* s_val, int* s_idx) {
  const int tid = threadIdx.x, nthr = blockDim.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  __syncthreads();   // the previous round's readers are done with s_val / s_idx
  if (tid % 32 == 0) { s_val[tid / 32] = bv; s_idx[tid / 32] = bi; }
  __syncthreads();
  if (tid < 32) {
    bv = tid < nthr / 32 ? s_val[tid] : -INFINITY;
    bi = tid < nthr / 32 ? s_idx[tid] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (tid == 0) { s_val[0] = bv; s_idx[0] = bi; }
  }
  __syncthreads();
  bv = s_val[0];
  bi = s_idx[0];
}

__device__ __forceinline__ float gumbel_of(uint32_t seed, uint32_t stepc, uint32_t row, uint32_t id) {
  const uint32_t u = hash_u32(seed, stepc, row, id);
  const float uni = ((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  return -logf(-logf(uni));   // argmax_i (l_i/temp + G_i)  ==  argmax_i softmax(l/temp)_i / Exp(1)_i
}

// top_k == 0: argmax.  1..64: ordered selection of the top-k (below), noise keyed by rank.  65..SAMPLE_CAND: threshold
// select (the k largest by (value desc, index asc), found with the histogram + candidate list), noise keyed by token id.
// top_k < 0: multinomial over all n_valid ids (sample_token with top_k == 0, utils/sampling.py:97-101).
// All threads of the block call it (any block size that is a multiple of 32, <= 1024); writes *token_out.
__device__ __noinline__ void sample_row(const bf16* __restrict__ lr, int n_valid, int top_k, float temp, uint32_t seed, uint32_t stepc,
                                        int row, long long* __restrict__ token_out) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ float top_v[64];
  __shared__ int top_i[64];
  __shared__ int hist[SAMPLE_HISTS][256];
  __shared__ float cand_v[SAMPLE_CAND];
  __shared__ int cand_i[SAMPLE_CAND];
  __shared__ int s_sel[5];   // [0] high-byte bin, [1] count above the threshold key, [2] threshold key, [3] candidate count, [4] ties taken
  const int tid = threadIdx.x, nthr = blockDim.x;

  if (top_k < 0) {   // full multinomial
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    const float inv_t = 1.0f / temp;
    for (int i = tid; i < n_valid; i += nthr) {
      const float sc = b2f(lr[i]) * inv_t + gumbel_of(seed, stepc, (uint32_t)row, (uint32_t)i);
      if (sc > bv) { bv = sc; bi = i; }
    }
    block_argmax(bv, bi, s_val, s_idx);
    if (tid == 0) *token_out = bi;
    return;
  }

  const bool big = top_k > 64;
  const int kk = top_k <= 0 ? 1 : (big ? top_k : top_k);
  int n_items = n_valid;
  bool from_list = false;
  if (big || (kk > 1 && n_valid > 4 * SAMPLE_CAND)) {
    int* myh = hist[(tid / 32) % SAMPLE_HISTS];
    for (int pass = 0; pass < 2; ++pass) {
      for (int i = tid; i < SAMPLE_HISTS * 256; i += nthr) (&hist[0][0])[i] = 0;
      __syncthreads();
      const int b1 = pass ? s_sel[0] : 0;
      for (int i = tid; i < n_valid; i += nthr) {
        const uint32_t k = bf16_order_key(lr[i]);
        if (pass == 0) atomicAdd(&myh[k >> 8], 1);
        else if ((int)(k >> 8) == b1) atomicAdd(&myh[k & 255u], 1);
      }
      __syncthreads();
      if (tid < 256) {
        int c = 0;
#pragma unroll
        for (int h = 0; h < SAMPLE_HISTS; ++h) c += hist[h][tid];
        hist[0][tid] = c;
      }
      __syncthreads();
      if (tid == 0) {
        int above = pass ? s_sel[1] : 0, b = 255;
        while (b > 0 && above + hist[0][b] < kk) { above += hist[0][b]; --b; }
        if (pass == 0) { s_sel[0] = b; s_sel[1] = above; }
        else { s_sel[2] = (s_sel[0] << 8) | b; s_sel[1] = above; s_sel[3] = 0; s_sel[4] = 0; }
      }
      __syncthreads();
    }
    const uint32_t thr = (uint32_t)s_sel[2];
    for (int i = tid; i < n_valid; i += nthr) {
      const bf16 v = lr[i];
      if (bf16_order_key(v) >= thr) {
        const int slot = atomicAdd(&s_sel[3], 1);
        if (slot < SAMPLE_CAND) { cand_v[slot] = b2f(v); cand_i[slot] = i; }
      }
    }
    __syncthreads();
    if (s_sel[3] <= SAMPLE_CAND) { from_list = true; n_items = s_sel[3]; }   // else: massive ties at the threshold
  }

  if (big) {
    const uint32_t thr = (uint32_t)s_sel[2];
    const int above = s_sel[1];
    const int need = kk - above;          // ties at the threshold to take, lowest ids first
    const float inv_t = 1.0f / temp;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    if (from_list) {
      for (int t = tid; t < n_items; t += nthr) {
        const float v = cand_v[t];
        const int id = cand_i[t];
        bool in = bf16_order_key(f2b(v)) > thr;
        if (!in) {
          int rank = 0;
          for (int j = 0; j < n_items; ++j) rank += (cand_v[j] == v && cand_i[j] < id) ? 1 : 0;
          in = rank < need;
        }
        if (in) {
          const float sc = v * inv_t + gumbel_of(seed, stepc, (uint32_t)row, (uint32_t)id);
          if (sc > bv || (sc == bv && id < bi)) { bv = sc; bi = id; }
        }
      }
    } else {
      // more than SAMPLE_CAND logits share the threshold value: walk the row in index order, counting ties
      for (int base = 0; base < n_valid; base += nthr) {
        const int i = base + tid;
        const uint32_t k = i < n_valid ? bf16_order_key(lr[i]) : 0u;
        const bool tie = i < n_valid && k == thr;
        const unsigned bal = __ballot_sync(0xffffffffu, tie);
        const int wpre = __popc(bal & ((1u << (tid % 32)) - 1u));
        __syncthreads();
        if (tid % 32 == 0) s_idx[tid / 32] = __popc(bal);
        __syncthreads();
        int before = s_sel[4];
        for (int w = 0; w < tid / 32; ++w) before += s_idx[w];
        const bool in = i < n_valid && (k > thr || (tie && before + wpre < need));
        if (in) {
          const float sc = b2f(lr[i]) * inv_t + gumbel_of(seed, stepc, (uint32_t)row, (uint32_t)i);
          if (sc > bv || (sc == bv && i < bi)) { bv = sc; bi = i; }
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < nthr / 32; ++w) t += s_idx[w]; s_sel[4] += t; }
        __syncthreads();
      }
    }
    block_argmax(bv, bi, s_val, s_idx);
    if (tid == 0) *token_out = bi;
    return;
  }

  float last_v = INFINITY;
  int last_i = -1;
  for (int r = 0; r < kk; ++r) {
    // largest (value, lowest index) strictly after (last_v, last_i) in the order (value desc, index asc)
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < n_items; i += nthr) {
      const float v = from_list ? cand_v[i] : b2f(lr[i]);
      const int id = from_list ? cand_i[i] : i;
      const bool after = v < last_v || (v == last_v && id > last_i);
      if (after && (v > bv || (v == bv && id < bi))) { bv = v; bi = id; }
    }
    block_argmax(bv, bi, s_val, s_idx);
    if (tid == 0) { top_v[r] = bv; top_i[r] = bi; }
    last_v = bv;
    last_i = bi;
  }
  if (tid == 0) {
    int pick = top_i[0];
    if (top_k > 0) {
      float best = -INFINITY;
      for (int r = 0; r < kk; ++r) {
        const float w = expf((top_v[r] - top_v[0]) / temp);
        const uint32_t u = hash_u32(seed, stepc, (uint32_t)row, (uint32_t)r);
        const float uni = ((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
        const float e = -logf(uni);                                        // Exp(1)
        const float score = w / e;
        if (score > best) { best = score; pick = top_i[r]; }
      }
    }
    *token_out = pick;
  }
}
