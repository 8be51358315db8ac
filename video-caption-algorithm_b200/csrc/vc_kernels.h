// Internal host-side declarations shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/vcb200.h"

namespace vc {

void set_error(const char* fmt, ...);

// Counts every kernel launch (bench.py "gpu_launches") and, while profiling is on,
// brackets it with CUDA events on the launch stream.
struct KernelScope {
  KernelScope(const char* name, double work, cudaStream_t stream);
  ~KernelScope();
  cudaStream_t stream_;
  int slot_;
};

// "Done once per device" flags for cudaFuncSetAttribute: the attribute lives per (function, device), so a second GPU in the same
// process (InferenceConfig.device = "cuda:1") needs its own opt-in to > 48 KB of dynamic shared memory.
struct PerDeviceOnce {
  bool done[64] = {false};
  // true if the caller still has to set the attribute on the current device (and marks it done)
  bool first(int* dev_out = nullptr) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (dev_out) *dev_out = dev;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// Launch with the programmatic-stream-serialization attribute (the kernel must begin with pdl_wait()).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// resize_kernels.cu — Pillow-exact antialiased bilinear resize (uint8 HWC), coefficient tables from the host
int resize_bilinear_u8(const uint8_t* src, int n, int H, int W, uint8_t* tmp, uint8_t* dst, int OH, int OW, const int32_t* kx, const int32_t* bx,
                       int ksize_x, const int32_t* ky, const int32_t* by, int ksize_y, cudaStream_t s);

// gemm_tcgen05.cu
// internal epilogues (continue the public VC_EPI_* numbering): LayerNorm folded into the product, residual update + row statistics
enum { VC_EPI_LNF_BIAS = 6, VC_EPI_LNF_GELU_ERF = 7, VC_EPI_LNF_GELU_TANH = 8, VC_EPI_RESID_STATS = 9 };
struct GemmExtra {
  const float* cs;       // LNF: column sums of the folded weights [N]
  const float* stats;    // LNF: float2 (mean, rstd) per row of A [M]
  float* xres;           // RESID_STATS: fp32 residual stream [M, N], x += bf16(acc + bias)
  void* xb_out;          // RESID_STATS: bf16 copy of the new residual [M, N]
  float* pstats;         // RESID_STATS: float2 partial (sum, sum of squares) [N/32][M], one slot per 32-column chunk
  float* stats_out;      // RESID_STATS, optional (few rows): float2 (mean, rstd) [M] written by the last CTA to finish ...
  unsigned int* done;    // ... which needs a device counter that is zero before the launch (it is reset to zero at the end)
  float eps;
  int split_k;           // RESID_STATS with 64-column tiles: 0 = chosen from the shape, 1 / 2 / 4 = forced (tests, A/B)
};
int gemm_bf16_ex(const void* A, const void* W, const float* bias, int M, int N, int K, int mode, void* out, int ldo,
                 const float* aux, int rows_per_group, int max_ctas, const GemmExtra* ex, cudaStream_t stream);
// tile width (256 / 64 columns) the GEMM picks for an [M, N] output; partial-statistics slots per row its RESID_STATS epilogue
// writes (= ln_stats_finalize's `parts`)
int gemm_tile_width(int M, int N);
int gemm_resid_parts(int N);
// float2 (mean, rstd) per row from `parts` partial (sum, sum of squares) pairs over N columns in total
int ln_stats_finalize(const float* pstats, int parts, int M, int N, float eps, float* stats, cudaStream_t s);
// xb = bf16(x), stats = (mean, rstd) per row of the fp32 rows x [M, dim]
int rowstats_cast(const float* x, void* xb_bf16, float* stats, int M, int dim, float eps, cudaStream_t s);
int gemm_bf16(const void* A, const void* W, const float* bias, int M, int N, int K, int mode, void* out, int ldo,
              const float* aux, int rows_per_group, int max_ctas, cudaStream_t stream);


// vit_kernels.cu
int preprocess_u8(const uint8_t* frames, const float* lut, void* out, int n, int H, int W, int layout, int patch, int k_pad,
                  cudaStream_t s);
int patchify_f32(const float* video, void* out, int n, int H, int W, int patch, int k_pad, cudaStream_t s);
int layernorm_f32_bf16(const float* x, const float* g, const float* b, void* out, int rows, int dim, float eps,
                       cudaStream_t s);
// fp32 rows picked with a stride (class tokens / last positions): out = LN(x[r*row_stride + row_offset])
int layernorm_rows(const float* x, long long row_stride, long long row_offset, const float* g, const float* b, float* out_f32,
                   void* out_bf16, int rows, int dim, float eps, cudaStream_t s);
// t = x[r] + delta[r] (+ delta2[r]) (bf16, may be null); out = LN(t); x[r] = t if write_x; r = i*row_stride + row_offset
int add_layernorm_rows(float* x, const void* delta_bf16, const void* delta2_bf16, int write_x, long long row_stride, long long row_offset,
                       const float* g, const float* b, float* out_f32, void* out_bf16, int rows, int dim, float eps, cudaStream_t s);
int vit_attention(const void* qkv, void* out, int n_frames, int tokens, int heads, int head_dim, cudaStream_t s);
int vit_attention_mma_sync(const void* qkv, void* out, int n_frames, int tokens, int heads, int head_dim, cudaStream_t s);
// vit_attention_tc.cu: tcgen05 / TMEM version (tokens <= 256, head_dim 64); vit_attention() dispatches to it
bool vit_attention_tc_supported(int tokens, int heads, int head_dim);
int vit_attention_tc(const void* qkv, void* out, int n_frames, int tokens, int heads, cudaStream_t s);
// attention output of the class-token query only: out bf16 [n_frames, heads*64]
int vit_cls_attention(const void* qkv, void* out, int n_frames, int tokens, int heads, int head_dim, cudaStream_t s);
// full attention of query rows row0..row0+n_rows-1 of every frame, written in place into out [n_frames*tokens, D]
int vit_row_attention(const void* qkv, void* out, int n_frames, int tokens, int heads, int row0, int n_rows, cudaStream_t s);
int cls_rows_init(float* x, const float* cls_pos0, int n_frames, int tokens, int dim, cudaStream_t s);
int pool_prefix(const float* cls, int B, int T, int dim, const float* head_w, const float* head_b, int video_dim,
                float ln_scale, float in_weight, const float* mapper_w, const float* mapper_b, int mapper_out,
                float* feat_out, float* prefix_out, cudaStream_t s);
int vit_pool_temporal(const void* feat, int is_bf16, int bsz, int T, int tokens, int C, int gap, float* out, cudaStream_t s);
int linear_bias_f32(const float* x, const float* w, const float* b, float* y, int rows, int in_f, int out_f, cudaStream_t s);

// gpt2_kernels.cu
int gpt_add_pos(const float* embeds, const float* wpe, float* h, int n_seq, int L, int past_len, int dim, cudaStream_t s);
// h = embeds + wpe[past_len + l]; xn = LayerNorm(h) as bf16 — one kernel
int gpt_add_pos_ln(const float* embeds, const float* wpe, float* h, void* xn_bf16, const float* gamma, const float* beta, int n_seq, int L,
                   int past_len, int dim, float eps, cudaStream_t s);
// qkv: bf16 [rows, 3H] (prefill) or, when P != null, fp32 split-K partials [ksplit][rows][3H] + bias (decode step)
int gpt_attention(const void* qkv, const float* P, int ksplit, const float* bias, void* out, const VcKvCache* cache, int layer,
                  int n_seq, int L, int past_len, cudaStream_t s);

// skinny_gemm.cu (decode step, M <= 128)
// out = bf16(gelu_new(x W^T + bias)), K split inside the CTA (no partial buffer); K % 256 == 0
int skinny_gemm_gelu(const void* x_bf16, const void* w_bf16, const float* bias, void* out_bf16, int M, int N, int K, cudaStream_t s);
int skinny_ksplit(int N, int K);
int skinny_ksplit(int N, int K, int row_tiles);   // fewer slices when the row tiles already supply the CTAs
int skinny_gemm(const void* x, const void* W, float* P, int M, int N, int K, int ksplit, cudaStream_t s);
int resid_ln(float* h, const float* P, int ksplit, const float* bias, const float* gamma, const float* beta, void* xn, int rows, int dim,
             float eps, cudaStream_t s);
int bias_act(const float* P, int ksplit, const float* bias, void* out, int M, int N, int gelu, cudaStream_t s);
int argmax_f32(const float* logits, long long ld, int rows, int vocab, int32_t* out, cudaStream_t s);
int embed_tokens(const void* wte_bf16, const int32_t* ids, int n, int dim, float* out, cudaStream_t s);
int greedy_init(int32_t* ids_out, int32_t* len_out, int32_t* finished, int n_seq, int max_new, int eos, cudaStream_t s);
// argmax of logits rows + the bookkeeping of benchmark_baseline.py:210-227 + wte gather for the next step
int greedy_select(const float* logits, long long ld, int vocab, int n_seq, int step, int max_new, int eos, int32_t* finished,
                  int32_t* ids_out, int32_t* len_out, const int32_t* forced, const void* wte_bf16, int dim, float* next_embeds,
                  int32_t* next_ids, cudaStream_t s);
int build_prefill_embeds(const float* prefix, const void* wte_bf16, const int32_t* prompt_ids, int n_seq, int P, int Lp,
                         int dim, float* out, cudaStream_t s);

// decode_chain.cu — the few-row GPT-2 forward as 5 dependent kernels per layer (LayerNorm folded into the consumer
// products, cluster split-K for the N = H products, lm_head with per-CTA argmax candidates)
struct ChainBuffers {
  float* h; void* hb; void* qkv; void* att; void* hid; float* stat; float* cand_v; int* cand_i;
};
bool chain_supported(const VcGptWeights* w, int rows);
int chain_lmhead_ctas();
// h = embeds + wpe[pos]; hb = bf16(h); stat[0] = row statistics
int chain_add_pos_stats(const VcGptWeights* w, const float* embeds, const ChainBuffers& b, int n_seq, int L, int past_len, cudaStream_t s);
// all layers + lm_head (candidates always; logits fp32 [n_seq, ld] when non-null) on rows prepared in h / hb / stat[0]
int chain_layers(const VcGptWeights* w, const ChainBuffers& b, int n_seq, int L, int past_len, VcKvCache* cache, float* logits, long long ld,
                 cudaStream_t s);
// argmax over candidates + greedy bookkeeping (benchmark_baseline.py:210-227) + next step's h / hb / stat[0] rows at next_pos
int chain_select(const VcGptWeights* w, const ChainBuffers& b, int n_seq, int step, int max_new, int eos, int32_t* finished, int32_t* ids_out,
                 int32_t* len_out, const int32_t* forced, int next_pos, bool feed_next, int32_t* next_ids, cudaStream_t s);
int chain_argmax(const ChainBuffers& b, int n_seq, int32_t* out, cudaStream_t s);
int chain_set_trace(void* buf, int max_records);

// beam_kernels.cu
int beam_step(const float* logits, long long ld, int vocab, int n_rows, int rows_per_item, const int32_t* seqs, int max_len, int cur_len,
              const float* running, float rep_penalty, int ngram, int min_new, int eos, int raw, int K, float* cand_score,
              int32_t* cand_tok, float* top_score, int32_t* top_idx, cudaStream_t s);
int beam_init(const VcBeamState* st, cudaStream_t s);
int beam_update(const VcBeamState* st, const float* top_score, const int32_t* top_idx, int vocab, int cur_len, float length_penalty, cudaStream_t s);
int beam_finalize(const VcBeamState* st, int32_t* ids_out, int32_t* len_out, cudaStream_t s);
int beam_reorder(const int32_t* slot_in, int32_t* slot_out, const int32_t* src_rows, int n_seq, int s_max, int upto, cudaStream_t s);

}  // namespace vc
