// Device side of HF-`generate` token selection (src/models/text_decoder.py:131-144 ->
// transformers GenerationMixin): log-softmax, the three logits processors the reference
// always enables (RepetitionPenalty -> NoRepeatNGram -> MinNewTokens), accumulated beam
// scores and the top-2*num_beams continuation search over num_beams * vocab candidates
// (`_get_top_k_continuations`), plus the KV-cache beam reorder as a slot-table update
// instead of HF's `index_select` copy of every layer's K/V.
// The tiny per-step bookkeeping over B x 2*num_beams candidates stays on the host (beam.py).
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

#define VC_LAUNCH(name, work, stream, ...)        \
  do {                                            \
    vc::KernelScope _ks(name, work, stream);      \
    __VA_ARGS__;                                  \
  } while (0)

namespace {

constexpr int BS_THREADS = 1024;
constexpr int BS_MAX_K = 16;

struct BestPair { float v; int i; };
__device__ __forceinline__ BestPair better(BestPair a, BestPair b) {
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ BestPair block_best(BestPair x, BestPair* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    BestPair y{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o)};
    x = better(x, y);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_red[warp] = x;
  __syncthreads();
  if (warp == 0) {
    x = lane < (blockDim.x >> 5) ? s_red[lane] : BestPair{-INFINITY, 0x7fffffff};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      BestPair y{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o)};
      x = better(x, y);
    }
    if (lane == 0) s_red[32] = x;
  }
  __syncthreads();
  const BestPair r = s_red[32];
  __syncthreads();
  return r;
}
__device__ float block_sum(float x, float* s_f) {
  x = warp_sum(x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_f[warp] = x;
  __syncthreads();
  if (warp == 0) {
    x = lane < (blockDim.x >> 5) ? s_f[lane] : 0.f;
    x = warp_sum(x);
    if (lane == 0) s_f[32] = x;
  }
  __syncthreads();
  const float r = s_f[32];
  __syncthreads();
  return r;
}

// One CTA per running row; the whole vocabulary row lives in shared memory (50257 x 4 B = 196 KB).
__global__ void __launch_bounds__(BS_THREADS, 1) beam_scores_kernel(const float* __restrict__ logits, long long ld, int vocab,
                                                                   const int32_t* __restrict__ seqs, int max_len, int cur_len,
                                                                   const float* __restrict__ running, float rep_penalty, int ngram,
                                                                   int min_new, int eos, int raw, int K, float* __restrict__ out_score,
                                                                   int32_t* __restrict__ out_tok) {
  extern __shared__ float sm[];                       // [vocab]
  __shared__ BestPair s_red[33];
  __shared__ float s_f[33];
  const int row = blockIdx.x;
  const float* src = logits + row * ld;
  const int32_t* seq = seqs + static_cast<long long>(row) * max_len;
  BestPair mx{-INFINITY, 0x7fffffff};
  for (int j = threadIdx.x; j < vocab; j += blockDim.x) {
    const float v = src[j];
    sm[j] = v;
    mx = better(mx, BestPair{v, j});
  }
  __syncthreads();
  if (!raw) {
    // log_softmax: (x - max) - log(sum(exp(x - max)))
    const float m = block_best(mx, s_red).v;
    float part = 0.f;
    for (int j = threadIdx.x; j < vocab; j += blockDim.x) part += expf(sm[j] - m);
    const float lse = logf(block_sum(part, s_f));
    for (int j = threadIdx.x; j < vocab; j += blockDim.x) sm[j] = (sm[j] - m) - lse;
    __syncthreads();
  }
  // RepetitionPenaltyLogitsProcessor: once per distinct previously generated id
  if (rep_penalty != 1.0f && threadIdx.x < cur_len) {
    const int tok = seq[threadIdx.x];
    bool first = true;
    for (int q = 0; q < threadIdx.x; ++q) first = first && (seq[q] != tok);
    if (first) {
      const float v = sm[tok];
      sm[tok] = v < 0.f ? v * rep_penalty : v / rep_penalty;
    }
  }
  __syncthreads();
  // NoRepeatNGramLogitsProcessor
  if (ngram > 0 && cur_len + 1 >= ngram) {
    const int s0 = threadIdx.x;
    if (s0 + ngram <= cur_len) {
      bool match = true;
      for (int q = 0; q < ngram - 1; ++q) match = match && (seq[s0 + q] == seq[cur_len + 1 - ngram + q]);
      if (match) sm[seq[s0 + ngram - 1]] = -INFINITY;
    }
  }
  // MinNewTokensLengthLogitsProcessor (prompt passed as embeds -> prompt length 0)
  if (threadIdx.x == 0 && min_new > 0 && cur_len < min_new) sm[eos] = -INFINITY;
  __syncthreads();
  const float rs = raw ? 0.f : running[row];
  for (int pick = 0; pick < K; ++pick) {
    BestPair b{-INFINITY, 0x7fffffff};
    for (int j = threadIdx.x; j < vocab; j += blockDim.x) b = better(b, BestPair{raw ? sm[j] : sm[j] + rs, j});
    b = block_best(b, s_red);
    if (threadIdx.x == 0) {
      out_score[row * K + pick] = b.v;
      out_tok[row * K + pick] = b.i == 0x7fffffff ? 0 : b.i;
      if (b.i != 0x7fffffff) sm[b.i] = -INFINITY;
    }
    __syncthreads();
  }
}

// per video: merge rows_per_item x K candidates into the top K over flat index beam*vocab + token
__global__ void __launch_bounds__(32) beam_merge_kernel(const float* __restrict__ cand_score, const int32_t* __restrict__ cand_tok,
                                                        int rows_per_item, int K, int vocab, float* __restrict__ top_score,
                                                        int32_t* __restrict__ top_idx) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n = rows_per_item * K;       // <= 16 * 16
  float v[8]; int id[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = lane + q * 32;
    if (c < n) {
      const int r = c / K;
      v[q] = cand_score[(static_cast<long long>(b) * rows_per_item + r) * K + (c - r * K)];
      id[q] = r * vocab + cand_tok[(static_cast<long long>(b) * rows_per_item + r) * K + (c - r * K)];
    } else {
      v[q] = -INFINITY; id[q] = 0x7fffffff;
    }
  }
  for (int pick = 0; pick < K; ++pick) {
    BestPair x{-INFINITY, 0x7fffffff};
#pragma unroll
    for (int q = 0; q < 8; ++q) x = better(x, BestPair{v[q], id[q]});
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      BestPair y{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o)};
      x = better(x, y);
    }
    if (lane == 0) { top_score[b * K + pick] = x.v; top_idx[b * K + pick] = x.i == 0x7fffffff ? 0 : x.i; }
#pragma unroll
    for (int q = 0; q < 8; ++q) if (id[q] == x.i) { v[q] = -INFINITY; id[q] = 0x7fffffff; }
  }
}

__global__ void beam_reorder_kernel(const int32_t* __restrict__ slot_in, int32_t* __restrict__ slot_out, const int32_t* __restrict__ src,
                                    int n_seq, int s_max, int upto) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seq * upto) return;
  const int r = i / upto, p = i - r * upto;
  slot_out[r * s_max + p] = slot_in[src[r] * s_max + p];
}

}  // namespace

int beam_step(const float* logits, long long ld, int vocab, int n_rows, int rows_per_item, const int32_t* seqs, int max_len, int cur_len,
              const float* running, float rep_penalty, int ngram, int min_new, int eos, int raw, int K, float* cand_score,
              int32_t* cand_tok, float* top_score, int32_t* top_idx, cudaStream_t s) {
  VC_REQUIRE(K >= 1 && K <= BS_MAX_K && rows_per_item >= 1 && rows_per_item <= 16 && n_rows % rows_per_item == 0,
             "beam_step: K=%d rows_per_item=%d n_rows=%d", K, rows_per_item, n_rows);
  VC_REQUIRE(cur_len <= max_len && cur_len < BS_THREADS, "beam_step: cur_len=%d", cur_len);
  const int smem = vocab * static_cast<int>(sizeof(float));
  VC_REQUIRE(smem <= 220 * 1024, "beam_step: vocab=%d does not fit in shared memory", vocab);
  static int attr = 0;
  if (smem > attr) {
    VC_CUDA_OK(cudaFuncSetAttribute(beam_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = smem;
  }
  VC_LAUNCH("beam_scores", static_cast<double>(n_rows) * vocab * 4.0, s,
            (beam_scores_kernel<<<n_rows, BS_THREADS, smem, s>>>(logits, ld, vocab, seqs, max_len, cur_len, running, rep_penalty, ngram,
                                                                 min_new, eos, raw, K, cand_score, cand_tok)));
  VC_CUDA_OK(cudaGetLastError());
  VC_LAUNCH("beam_merge", 0.0, s,
            (beam_merge_kernel<<<n_rows / rows_per_item, 32, 0, s>>>(cand_score, cand_tok, rows_per_item, K, vocab, top_score, top_idx)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int beam_reorder(const int32_t* slot_in, int32_t* slot_out, const int32_t* src_rows, int n_seq, int s_max, int upto, cudaStream_t s) {
  if (n_seq <= 0 || upto <= 0) return 0;
  VC_REQUIRE(upto <= s_max, "beam_reorder: upto=%d > s_max=%d", upto, s_max);
  const int total = n_seq * upto;
  VC_LAUNCH("beam_reorder", total * 8.0, s, (beam_reorder_kernel<<<(total + 255) / 256, 256, 0, s>>>(slot_in, slot_out, src_rows, n_seq, s_max, upto)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vc
