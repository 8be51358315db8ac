// Device side of HF-`generate` token selection (src/models/text_decoder.py:131-144 ->
// transformers GenerationMixin): log-softmax, the three logits processors the reference
// always enables (RepetitionPenalty -> NoRepeatNGram -> MinNewTokens), accumulated beam
// scores and the top-2*num_beams continuation search over num_beams * vocab candidates
// (`_get_top_k_continuations`), plus the KV-cache beam reorder as a slot-table update
// instead of HF's `index_select` copy of every layer's K/V.
// The per-step bookkeeping over B x 2*num_beams candidates (running / finished hypotheses, early-stop heuristic) is one small
// kernel per step (beam_update_kernel), so the whole beam loop is a fixed sequence of launches: CUDA-graph capturable.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <algorithm>
#include <cmath>

namespace vc {

#define VC_LAUNCH(name, work, stream, ...)        \
  do {                                            \
    vc::KernelScope _ks(name, work, stream);      \
    __VA_ARGS__;                                  \
  } while (0)

namespace {

constexpr int BS_THREADS = 1024;
constexpr int BS_MAX_K = 16;
constexpr int BS_LIST = 256;     // scores at or above the selection threshold that beam_scores_kernel orders in one warp

struct BestPair { float v; int i; };
__device__ __forceinline__ BestPair better(BestPair a, BestPair b) {
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ BestPair warp_best(BestPair x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    BestPair y{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o)};
    x = better(x, y);
  }
  return x;
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
  extern __shared__ __align__(16) float sm[];         // [vocab]
  __shared__ BestPair s_red[33];
  __shared__ float s_f[33];
  const int row = blockIdx.x;
  const float* src = logits + row * ld;
  const int32_t* seq = seqs + static_cast<long long>(row) * max_len;
  // the row -> shared memory: 16-byte asynchronous copies (nothing staged in registers, every copy of the thread in flight at
  // once) when the row starts on a 16-byte boundary, element loads otherwise; the last vocab % 4 elements always by element
  const int n4 = (reinterpret_cast<uintptr_t>(src) & 15) == 0 ? vocab >> 2 : 0;
  {
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
    for (int j = threadIdx.x; j < n4; j += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + j * 16), "l"(src + 4 * j) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int j = 4 * n4 + threadIdx.x; j < vocab; j += blockDim.x) sm[j] = src[j];
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  float4* sm4 = reinterpret_cast<float4*>(sm);
  const int v4 = vocab >> 2;                          // float4 passes over shared memory + a tail of vocab % 4
  if (!raw) {
    // log_softmax: (x - max) - log(sum(exp(x - max)))
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < v4; j += blockDim.x) {
      const float4 x = sm4[j];
      mx = fmaxf(fmaxf(mx, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
    }
    for (int j = 4 * v4 + threadIdx.x; j < vocab; j += blockDim.x) mx = fmaxf(mx, sm[j]);
    const float m = block_best(BestPair{mx, 0}, s_red).v;
    float part = 0.f;
    for (int j = threadIdx.x; j < v4; j += blockDim.x) {
      const float4 x = sm4[j];
      part += (expf(x.x - m) + expf(x.y - m)) + (expf(x.z - m) + expf(x.w - m));
    }
    for (int j = 4 * v4 + threadIdx.x; j < vocab; j += blockDim.x) part += expf(sm[j] - m);
    const float lse = logf(block_sum(part, s_f));
    for (int j = threadIdx.x; j < v4; j += blockDim.x) {
      float4 x = sm4[j];
      x.x = (x.x - m) - lse; x.y = (x.y - m) - lse; x.z = (x.z - m) - lse; x.w = (x.w - m) - lse;
      sm4[j] = x;
    }
    for (int j = 4 * v4 + threadIdx.x; j < vocab; j += blockDim.x) sm[j] = (sm[j] - m) - lse;
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
  // ---- top K of the row by (score descending, token ascending).
  // A threshold first: the K-th largest of the 1024 per-thread maxima is a lower bound of the row's K-th largest score (K
  // different elements reach it), so the top K are among the elements >= that threshold — a handful.  Only the threads whose
  // own maximum reaches it rescan their elements; one warp then orders the short list.  (K full passes over the row with a
  // block-wide argmax each cost 200 us per step at 320 rows.)
  __shared__ BestPair s_wtop[32 * BS_MAX_K];
  __shared__ BestPair s_list[BS_LIST];
  __shared__ float s_thr;
  __shared__ int s_n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  BestPair loc{-INFINITY, 0x7fffffff};
  for (int j = threadIdx.x; j < v4; j += blockDim.x) {           // the thread's elements: float4 groups j, plus the tail
    const float4 x = sm4[j];
    loc = better(loc, BestPair{raw ? x.x : x.x + rs, 4 * j});
    loc = better(loc, BestPair{raw ? x.y : x.y + rs, 4 * j + 1});
    loc = better(loc, BestPair{raw ? x.z : x.z + rs, 4 * j + 2});
    loc = better(loc, BestPair{raw ? x.w : x.w + rs, 4 * j + 3});
  }
  for (int j = 4 * v4 + threadIdx.x; j < vocab; j += blockDim.x) loc = better(loc, BestPair{raw ? sm[j] : sm[j] + rs, j});
  if (threadIdx.x == 0) s_n = 0;
  {
    BestPair cur = loc;
    for (int pick = 0; pick < K; ++pick) {
      const BestPair w = warp_best(cur);
      if (lane == 0) s_wtop[warp * K + pick] = w;
      if (cur.i == w.i) cur = BestPair{-INFINITY, 0x7fffffff};
    }
  }
  __syncthreads();
  if (warp == 0) {
    BestPair mine[BS_MAX_K];                          // entries lane, lane + 32, ... of the 32 * K warp results
#pragma unroll
    for (int q = 0; q < BS_MAX_K; ++q) mine[q] = q < K ? s_wtop[q * 32 + lane] : BestPair{-INFINITY, 0x7fffffff};
    BestPair w{-INFINITY, 0x7fffffff};
    for (int pick = 0; pick < K; ++pick) {
      BestPair x{-INFINITY, 0x7fffffff};
#pragma unroll
      for (int q = 0; q < BS_MAX_K; ++q) x = better(x, mine[q]);
      w = warp_best(x);
#pragma unroll
      for (int q = 0; q < BS_MAX_K; ++q) if (mine[q].i == w.i) mine[q] = BestPair{-INFINITY, 0x7fffffff};
    }
    if (lane == 0) s_thr = w.i == 0x7fffffff ? -INFINITY : w.v;
  }
  __syncthreads();
  const float thr = s_thr;
  if (loc.i != 0x7fffffff && loc.v >= thr) {
    auto take = [&](float x, int j) {
      const float v = raw ? x : x + rs;
      if (v >= thr) {
        const int pos = atomicAdd(&s_n, 1);
        if (pos < BS_LIST) s_list[pos] = BestPair{v, j};
      }
    };
    for (int j = threadIdx.x; j < v4; j += blockDim.x) {
      const float4 x = sm4[j];
      take(x.x, 4 * j); take(x.y, 4 * j + 1); take(x.z, 4 * j + 2); take(x.w, 4 * j + 3);
    }
    for (int j = 4 * v4 + threadIdx.x; j < vocab; j += blockDim.x) take(sm[j], j);
  }
  __syncthreads();
  const int n_list = s_n;
  if (n_list >= K && n_list <= BS_LIST && thr > -INFINITY) {
    if (warp == 0) {
      BestPair mine[BS_LIST / 32];
#pragma unroll
      for (int q = 0; q < BS_LIST / 32; ++q) mine[q] = q * 32 + lane < n_list ? s_list[q * 32 + lane] : BestPair{-INFINITY, 0x7fffffff};
      for (int pick = 0; pick < K; ++pick) {
        BestPair x{-INFINITY, 0x7fffffff};
#pragma unroll
        for (int q = 0; q < BS_LIST / 32; ++q) x = better(x, mine[q]);
        const BestPair w = warp_best(x);
        if (lane == 0) { out_score[row * K + pick] = w.v; out_tok[row * K + pick] = w.i; }
#pragma unroll
        for (int q = 0; q < BS_LIST / 32; ++q) if (mine[q].i == w.i) mine[q] = BestPair{-INFINITY, 0x7fffffff};
      }
    }
    return;
  }
  // degenerate rows (fewer than K finite scores, or very many scores tied at the threshold): K passes over the row
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

// ---------------------------------------------------------------------------------------------- beam bookkeeping
// transformers `_beam_search` per-step state update (installed generation/utils.py: _get_running_beams_for_next_iteration,
// _update_finished_beams, _check_early_stop_heuristic, _beam_search_has_unfinished_sequences) for ONE video per
// CTA (32 threads; lane k = candidate k of the top 2*num_beams).  Arithmetic mirrors the torch formulation op for op
// (score + stop * -1e9, score / len**lp + ... ), selection is top-n by (value descending, index ascending).
constexpr float BM_NEG = -1.0e9f;
constexpr int BM_MAX_LEN = 256;

__device__ __forceinline__ int rank_among(float v, int self, int n, bool valid) {
  // position of (v, self) in the descending order of the n values held by lanes 0..n-1 (ties: lower lane first)
  int r = 0;
  for (int j = 0; j < n; ++j) {
    const float vj = __shfl_sync(0xffffffffu, v, j);
    r += (vj > v || (vj == v && j < self)) ? 1 : 0;
  }
  return valid ? r : 0x7fffffff;
}

__global__ void __launch_bounds__(32) beam_update_kernel(VcBeamState st, const float* __restrict__ top_score, const int32_t* __restrict__ top_idx,
                                                         int vocab, int cur_len, float den) {
  __shared__ int32_t s_run[16][BM_MAX_LEN], s_fin[16][BM_MAX_LEN];
  __shared__ int s_pick[16], s_pbeam[16], s_ptok[16], s_sel[16], s_sbeam[16], s_stok[16];
  __shared__ float s_fs[16];
  __shared__ int s_fd[16];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int nb = st.nb, K = 2 * nb, L = st.max_len;
  const int new_len = cur_len + 1;
  // has HF's loop already ended?  (evaluated from the flags the PREVIOUS step's CTAs left; sticky)
  int stopped = *st.stopped;
  if (cur_len > 0) {
    const int any_unsat = st.flags[2 * (cur_len - 1)], all_hit = st.flags[2 * (cur_len - 1) + 1];
    stopped |= !(any_unsat != 0 && all_hit == 0);
  }
  for (int i = lane; i < nb * L; i += 32) {
    s_run[i / L][i % L] = st.running_seqs[static_cast<long long>(b) * nb * L + i];
    s_fin[i / L][i % L] = st.fin_seqs[static_cast<long long>(b) * nb * L + i];
  }
  const bool cand = lane < K;
  const float score = cand ? top_score[b * K + lane] : -INFINITY;
  const int flat = cand ? top_idx[b * K + lane] : 0;
  const int beam = flat / vocab, tok = flat - beam * vocab;
  const bool hit = cand && (tok == st.eos || new_len >= L);
  __syncwarp();
  // ---- running beams of the next iteration: top nb of (score, stopped candidates pushed down by -1e9)
  const float run_v = score + (hit ? 1.f : 0.f) * BM_NEG;
  const int r_run = rank_among(run_v, lane, K, cand);
  if (r_run < nb) {
    s_pick[r_run] = lane; s_pbeam[r_run] = beam; s_ptok[r_run] = tok;
    st.running_scores[b * nb + r_run] = run_v;
    st.src_rows[b * nb + r_run] = b * nb + beam;
    st.next_tok[b * nb + r_run] = tok;
  }
  // ---- finished pool: candidates in the first nb slots that just stopped, merged with the pool, top nb by normalised score
  const int unsat = st.unsatisfied[b];
  const bool newly = hit && lane < nb;
  float f = score / den;
  f = f + (unsat ? 0.f : 1.f) * BM_NEG;
  f = f + (newly ? 0.f : 1.f) * BM_NEG;
  // merged order: lanes 0..nb-1 = pool entries, lanes nb..nb+K-1 = candidates  (torch.cat([fin, cand]))
  const float pool_v = lane < nb ? st.fin_scores[b * nb + lane] : -INFINITY;
  const float cand_f = __shfl_sync(0xffffffffu, f, lane >= nb ? lane - nb : 0);
  const bool m_valid = lane < nb + K;
  const float m_v = lane < nb ? pool_v : (m_valid ? cand_f : -INFINITY);
  const int r_fin = rank_among(m_v, lane, nb + K, m_valid);
  const int pool_done = lane < nb ? st.fin_done[b * nb + lane] : 0;
  const int pool_len = lane < nb ? st.fin_len[b * nb + lane] : 0;
  const int c_newly = __shfl_sync(0xffffffffu, newly ? 1 : 0, lane >= nb ? lane - nb : 0);
  const int c_beam = __shfl_sync(0xffffffffu, beam, lane >= nb ? lane - nb : 0), c_tok = __shfl_sync(0xffffffffu, tok, lane >= nb ? lane - nb : 0);
  if (r_fin < nb) {
    s_sel[r_fin] = lane; s_sbeam[r_fin] = c_beam; s_stok[r_fin] = c_tok;
    s_fs[r_fin] = m_v;
    s_fd[r_fin] = lane < nb ? pool_done : c_newly;
  }
  __syncwarp();
  const int sel_len = (r_fin < nb) ? (lane < nb ? pool_len : new_len) : 0;
  // ---- write the new running sequences (old sequence of the source beam + the chosen token)
  for (int i = lane; i < nb * L; i += 32) {
    const int r = i / L, p = i - r * L;
    st.running_seqs[static_cast<long long>(b) * nb * L + i] = p == cur_len ? s_ptok[r] : s_run[s_pbeam[r]][p];
  }
  // ---- write the new finished pool unless HF's loop has already ended
  if (!stopped) {
    for (int i = lane; i < nb * L; i += 32) {
      const int r = i / L, p = i - r * L;
      const int src = s_sel[r];
      st.fin_seqs[static_cast<long long>(b) * nb * L + i] = src < nb ? s_fin[src][p] : (p == cur_len ? s_stok[r] : s_run[s_sbeam[r]][p]);
    }
    if (r_fin < nb) {
      st.fin_scores[b * nb + r_fin] = m_v;
      st.fin_done[b * nb + r_fin] = s_fd[r_fin];
      st.fin_len[b * nb + r_fin] = sel_len;
    }
  }
  // ---- early-stop heuristic on the (possibly frozen) pool
  float fs = 0.f; int fd = 0;
  if (lane < nb) {
    fs = stopped ? st.fin_scores[b * nb + lane] : s_fs[lane];
    fd = stopped ? st.fin_done[b * nb + lane] : s_fd[lane];
  }
  float mn = lane < nb ? fs : INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  const float best_running = __shfl_sync(0xffffffffu, run_v, s_pick[0]) / den;
  const bool gt = lane < nb && best_running > (fd ? mn : BM_NEG);
  const unsigned any_gt = __ballot_sync(0xffffffffu, gt);
  const int unsat_new = unsat && any_gt != 0;
  const unsigned hits = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) {
    st.unsatisfied[b] = unsat_new;
    if (unsat_new) atomicOr(&st.flags[2 * cur_len], 1);
    if (hits != ((K >= 32) ? 0xffffffffu : ((1u << K) - 1))) atomicAnd(&st.flags[2 * cur_len + 1], 0);
    if (b == 0) *st.stopped = stopped;
  }
}

__global__ void beam_init_kernel(VcBeamState st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_rows = st.B * st.nb;
  if (i < n_rows * st.max_len) { st.running_seqs[i] = st.eos; st.fin_seqs[i] = st.eos; }
  if (i < n_rows) {
    st.running_scores[i] = (i % st.nb) == 0 ? 0.f : BM_NEG;
    st.fin_scores[i] = BM_NEG; st.fin_done[i] = 0; st.fin_len[i] = 0;
  }
  if (i < st.B) st.unsatisfied[i] = 1;
  if (i <= st.max_len) { st.flags[2 * i] = 0; st.flags[2 * i + 1] = 1; }
  if (i == 0) *st.stopped = 0;
}

// best finished hypothesis per video, cropped to its length, eos padded
__global__ void beam_finalize_kernel(VcBeamState st, int32_t* __restrict__ ids_out, int32_t* __restrict__ len_out) {
  const int b = blockIdx.x;
  const int n = st.fin_len[b * st.nb];
  for (int p = threadIdx.x; p < st.max_len; p += blockDim.x)
    ids_out[b * st.max_len + p] = p < n ? st.fin_seqs[static_cast<long long>(b) * st.nb * st.max_len + p] : st.eos;
  if (threadIdx.x == 0) len_out[b] = n;
}

}  // namespace

int beam_init(const VcBeamState* st, cudaStream_t s) {
  VC_REQUIRE(st != nullptr && st->nb >= 1 && st->nb <= 8 && st->max_len >= 1 && st->max_len <= BM_MAX_LEN && st->B >= 1,
             "beam_init: B=%d nb=%d max_len=%d", st ? st->B : 0, st ? st->nb : 0, st ? st->max_len : 0);
  const int total = std::max(st->B * st->nb * st->max_len, st->max_len + 1);
  VC_LAUNCH("beam_init", 0.0, s, (beam_init_kernel<<<(total + 255) / 256, 256, 0, s>>>(*st)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int beam_update(const VcBeamState* st, const float* top_score, const int32_t* top_idx, int vocab, int cur_len, float length_penalty, cudaStream_t s) {
  VC_REQUIRE(st != nullptr && st->nb >= 1 && st->nb <= 8 && st->max_len <= BM_MAX_LEN && cur_len >= 0 && cur_len < st->max_len,
             "beam_update: nb=%d max_len=%d cur_len=%d", st ? st->nb : 0, st ? st->max_len : 0, cur_len);
  const float den = static_cast<float>(std::pow(static_cast<double>(cur_len + 1), static_cast<double>(length_penalty)));
  VC_LAUNCH("beam_update", 0.0, s, (beam_update_kernel<<<st->B, 32, 0, s>>>(*st, top_score, top_idx, vocab, cur_len, den)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int beam_finalize(const VcBeamState* st, int32_t* ids_out, int32_t* len_out, cudaStream_t s) {
  VC_LAUNCH("beam_finalize", 0.0, s, (beam_finalize_kernel<<<st->B, 64, 0, s>>>(*st, ids_out, len_out)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int beam_step(const float* logits, long long ld, int vocab, int n_rows, int rows_per_item, const int32_t* seqs, int max_len, int cur_len,
              const float* running, float rep_penalty, int ngram, int min_new, int eos, int raw, int K, float* cand_score,
              int32_t* cand_tok, float* top_score, int32_t* top_idx, cudaStream_t s) {
  VC_REQUIRE(K >= 1 && K <= BS_MAX_K && rows_per_item >= 1 && rows_per_item <= 16 && n_rows % rows_per_item == 0,
             "beam_step: K=%d rows_per_item=%d n_rows=%d", K, rows_per_item, n_rows);
  VC_REQUIRE(cur_len <= max_len && cur_len < BS_THREADS, "beam_step: cur_len=%d", cur_len);
  const int smem = vocab * static_cast<int>(sizeof(float));
  VC_REQUIRE(smem <= 220 * 1024, "beam_step: vocab=%d does not fit in shared memory", vocab);
  {
    static int attr[64] = {0};                 // per device: a second GPU in the same process needs its own opt-in
    int dev = 0;
    VC_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || smem > attr[dev]) {
      VC_CUDA_OK(cudaFuncSetAttribute(beam_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      if (dev >= 0 && dev < 64) attr[dev] = smem;
    }
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
