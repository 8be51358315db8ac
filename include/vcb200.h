/* vcb200.h — C ABI of the B200-native video-caption hot path.
 *
 * Drop-in boundary.  The reference is 100 % Python; its plugin
 * surface for this path is (1) the module attributes `model.encoder(video)`,
 * `model.decoder.mapper(x)`, `model.decoder.model(inputs_embeds=…, past_key_values=…)`
 * used by core/engine.py:43-61 and core/scripts/benchmark_baseline.py:162-289,
 * (2) the backend switch in core/models/model_loader.py:21-28 and (3) the CuPy
 * operator hooks core/operators/cupy_vit_pool.py:127-186 and
 * core/operators/cupy_linear_mapper.py:137-184.  A maintainer binds this library
 * with ctypes; the Python adapter in
 * video-caption-algorithm_b200/ presents surface (1) on top of it.
 *
 * Conventions (same as the reference's hooks): every pointer is DEVICE memory
 * owned by the caller (torch), nothing is allocated inside, work is enqueued on
 * the stream passed in (the reference uses torch's current stream:
 * cupy_vit_pool.py:162-163), calls return 0 on success and a negative code on
 * error with the text in vc_last_error().  Unlike the reference's hooks there is
 * NO fallback: an error is an error (north_star: "no CPU fallback").
 */
#ifndef VCB200_H
#define VCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vc_stream_t; /* cudaStream_t */

/* GEMM epilogues (vc_gemm_bf16) */
enum {
  VC_EPI_BIAS = 0,           /* bf16 out = acc + bias                                  */
  VC_EPI_BIAS_GELU_ERF = 1,  /* bf16 out = gelu_erf(acc + bias)   torchvision nn.GELU() */
  VC_EPI_BIAS_GELU_TANH = 2, /* bf16 out = gelu_tanh(acc + bias)  timm patch / gelu_new */
  VC_EPI_BIAS_RESID_F32 = 3, /* fp32 out += acc + bias            residual stream       */
  VC_EPI_BIAS_F32 = 4,       /* fp32 out = acc + bias                                  */
  VC_EPI_PATCH_EMBED = 5     /* fp32 out[g*(rpg+1)+1+r] = acc + bias + pos[1+r]         */
};

/* ---- ViT encoder weights: packed once from the model state_dict ------------------------ */
typedef struct {
  const float* ln1_g; const float* ln1_b;   /* LayerNorm eps 1e-6 */
  const void* qkv_w;  const float* qkv_b;   /* bf16 [3D, D], rows q|k|v (in_proj_weight / attn.qkv) */
  const void* proj_w; const float* proj_b;  /* bf16 [D, D] */
  const float* ln2_g; const float* ln2_b;
  const void* fc1_w;  const float* fc1_b;   /* bf16 [mlp, D] */
  const void* fc2_w;  const float* fc2_b;   /* bf16 [D, mlp] */
  /* LayerNorm folded into the consumer product (see VcGptLayer): ln_1 into qkv, ln_2 into fc1.  NULL = not packed
   * (vc_vit_encode then runs the stand-alone residual-add + LayerNorm passes). */
  const void* qkv_wf; const float* qkv_cs; const float* qkv_bf;
  const void* fc1_wf; const float* fc1_cs; const float* fc1_bf;
} VcVitLayer;

typedef struct {
  int32_t dim, layers, heads, mlp, tokens, patch_k; /* patch_k: 3*p*p padded to a multiple of 64 */
  int32_t gelu_tanh;                                 /* 0: erf (torchvision path) 1: tanh (timm path, video_encoder.py:134) */
  int32_t video_dim;                                 /* 256 */
  const void* patch_w;  const float* patch_b;        /* bf16 [dim, patch_k] (conv_proj flattened c,i,j) */
  const float* cls_pos0;                             /* fp32 [dim]: class_token + pos_embedding[0] */
  const float* pos;                                  /* fp32 [tokens, dim] */
  const float* lnf_g; const float* lnf_b;            /* encoder.ln / norm */
  const float* head_w; const float* head_b;          /* fp32 encoder.proj [video_dim, dim] */
  const VcVitLayer* layer;                           /* host array [layers] */
} VcVitWeights;

/* ---- GPT-2 decoder weights --------------------------------------------------------------- */
typedef struct {
  const float* ln1_g; const float* ln1_b;    /* eps 1e-5 */
  const void* attn_w;  const float* attn_b;  /* bf16 [3H, H] = c_attn.weight^T (Conv1D is [in,out]) */
  const void* aproj_w; const float* aproj_b; /* bf16 [H, H]  = attn.c_proj.weight^T */
  const float* ln2_g; const float* ln2_b;
  const void* fc_w;    const float* fc_b;    /* bf16 [4H, H] */
  const void* mproj_w; const float* mproj_b; /* bf16 [H, 4H] */
  /* LayerNorm folded into the consumer product (decode chain): W' = bf16(gamma (.) W) [out, H], cs[n] = sum_k W'[n][k],
   * bf[n] = b[n] + sum_k beta[k] W[n][k]  so that  LN(h) W^T + b = rstd (h W'^T - mean cs) + bf.  NULL = not packed. */
  const void* attn_wf; const float* attn_cs; const float* attn_bf;   /* ln_1 into c_attn */
  const void* fc_wf;   const float* fc_cs;   const float* fc_bf;     /* ln_2 into c_fc   */
} VcGptLayer;

typedef struct {
  int32_t dim, layers, heads, vocab, n_pos, vocab_pad; /* vocab_pad: rows of wte padded to a multiple of 64 (zero rows) */
  const void* wte;            /* bf16 [vocab_pad, H] (tied lm_head) */
  const float* wpe;           /* fp32 [n_pos, H] */
  const float* lnf_g; const float* lnf_b;
  const void* lmh_w;          /* bf16 [vocab_pad, H] = bf16(ln_f.gamma (.) wte): ln_f folded into the tied lm_head (NULL = not packed) */
  const float* lmh_cs; const float* lmh_b;   /* fp32 [vocab_pad]: column sums of lmh_w, wte . ln_f.beta */
  const VcGptLayer* layer;    /* host array [layers] */
} VcGptWeights;

/* contiguous bf16 KV cache [layers][2][n_seq][heads][s_max][head_dim] + per-(seq,pos) slot table for beams */
typedef struct {
  void* kv;                 /* bf16 */
  int32_t* slot;            /* int32 [n_seq, s_max]: physical sequence row holding position p of logical row s */
  int32_t layers, n_seq, heads, s_max, head_dim;
} VcKvCache;

/* ---- library state ----------------------------------------------------------------------- */
const char* vc_last_error(void);
int vc_abi_version(void);
int vc_num_sms(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long vc_launch_count(void);
/* optional in-kernel timeline of the decode chain (tools/trace_decode.py): device buffer of 8 + 8*max_records uint64, buf[0] = record
 * counter (zero it first); every record = {kernel id | cta << 8, %globaltimer at entry, after the dependency wait, after the
 * activations arrived, after the MMAs, after the reduction barrier, -, at exit}.  NULL switches it off. */
int vc_debug_trace(void* buf_u64, int max_records);
/* optional per-kernel CUDA-event timing on the launch stream (bench.py roofline pass) */
int vc_prof_begin(void);
int vc_prof_end(int max_rows, char* names /*[max_rows][48]*/, float* total_ms, int* calls, double* work);

/* ---- a1  preprocessing: core/preprocessing/frame_loader.py:34-47 ------------------------ */
/* a1' — frame resize.  Replaces `transforms.Resize((image_size, image_size))` applied to the PIL image of every frame
 * (core/preprocessing/frame_loader.py:34-45; torchvision -> PIL.Image.resize(BILINEAR) -> Pillow ImagingResample, 8 bpc):
 * antialiased separable bilinear resample, horizontal pass first, uint8 intermediate, 22-bit fixed-point weights.
 * src_hwc: uint8 [n, in_h, in_w, 3]; dst_hwc: uint8 [n, out_h, out_w, 3]; scratch: uint8 [n, in_h, out_w, 3] (may be NULL
 * when in_h == out_h).  kx/ky: int32 [out, ksize] weights, bounds: int32 [out, 2] = (first input index, taps), built on the
 * host with Pillow's rule (video-caption-algorithm_b200/resample.py: pillow_bilinear_coeffs).  Byte-exact vs Pillow. */
int vc_resize_bilinear_u8(const uint8_t* src_hwc, int n_frames, int in_h, int in_w, uint8_t* scratch, uint8_t* dst_hwc, int out_h, int out_w,
                          const int32_t* kx, const int32_t* bounds_x, int ksize_x, const int32_t* ky, const int32_t* bounds_y, int ksize_y,
                          vc_stream_t stream);

/* uint8 HWC frames -> bf16 through the 3x256 LUT of ToTensor+Normalize.
 * layout 0: [n,3,H,W] (the reference tensor, bf16-rounded); layout 1: patch-major
 * [n*(H/p)*(W/p), k_pad] with column c*p*p + i*p + j — the A operand of the patch-embed GEMM. */
int vc_preprocess_u8(const uint8_t* frames_hwc, const float* lut3x256, void* out_bf16, int n_frames, int H, int W,
                     int layout, int patch, int k_pad, vc_stream_t stream);

/* fp32 [n,3,H,W] normalised tensor (what core/engine.py:43 hands to model.encoder) -> bf16 patch-major rows */
int vc_patchify_f32(const float* video_chw, void* out_bf16, int n_frames, int H, int W, int patch, int k_pad, vc_stream_t stream);

/* ---- a2  ViT encoder building blocks + whole encoder: src/models/video_encoder.py:288-326 */
int vc_gemm_bf16(const void* A, const void* W, const float* bias, int M, int N, int K, int epilogue, void* out, int ldo,
                 const float* aux, int rows_per_group, vc_stream_t stream);
/* The residual half of a transformer block as ONE product (video_encoder.py:168-171 `x = x + attn(...)` / `x + mlp(...)`,
 * transformers GPT2Block residual adds):  x[M,N] (fp32, in place) += bf16(A W^T + bias);  xb = bf16(x);  and the rows' partial
 * (sum, sum of squares) per 32-column chunk in pstats (float2 [N/32][M]) for the LayerNorm folded into the NEXT product.
 * stats_out / done (both or neither): with few rows the last CTA turns the partials into float2 (mean, rstd) [M]; `done` is a
 * device counter that is zero before the call.  split_k: 0 = chosen from the shape; 1, 2, 4 = forced (64-column tiles only). */
int vc_gemm_resid_stats(const void* A, const void* W, const float* bias, int M, int N, int K, float* x, void* xb_bf16,
                        float* pstats, float* stats_out, unsigned int* done, float eps, int split_k, vc_stream_t stream);
int vc_layernorm_f32_bf16(const float* x, const float* gamma, const float* beta, void* out_bf16, int rows, int dim,
                          float eps, vc_stream_t stream);
int vc_vit_attention(const void* qkv_bf16, void* out_bf16, int n_frames, int tokens, int heads, int head_dim,
                     vc_stream_t stream);
/* the mma.sync kernel vc_vit_attention falls back to for tokens > 256 (ViT-L/14: 257); exported for A/B tests */
int vc_vit_attention_mma_sync(const void* qkv_bf16, void* out_bf16, int n_frames, int tokens, int heads, int head_dim,
                              vc_stream_t stream);
size_t vc_vit_workspace_bytes(const VcVitWeights* w, int chunk_frames);
/* patches: bf16 [n_frames*(tokens-1), patch_k]; cls_out: fp32 [n_frames, dim] = final-LN class token per frame */
int vc_vit_encode(const VcVitWeights* w, const void* patches_bf16, int n_frames, int chunk_frames, void* workspace,
                  size_t workspace_bytes, float* cls_out, vc_stream_t stream);

/* ---- a3-a6  pool + proj + prefix norm + mapper: video_encoder.py:256-258,316; engine.py:45-50;
 *      text_decoder.py:36-45,69; replaces cupy_vit_pool.py / cupy_linear_mapper.py ---------- */
int vc_pool_prefix(const float* cls_tokens /*[B*T, dim]*/, int B, int T, int dim, const float* head_w, const float* head_b,
                   int video_dim, float ln_scale, float in_weight, const float* mapper_w /*[out, video_dim]*/,
                   const float* mapper_b, int mapper_out, float* feat_out /*[B,video_dim]*/, float* prefix_out /*[B,mapper_out]*/,
                   vc_stream_t stream);
/* the two CuPy hooks as stand-alone operators (same contracts, torch-owned tensors) */
int vc_vit_pool_temporal(const void* feat, int is_bf16, int bsz, int timesteps, int tokens, int channels, int gap,
                         float* out /*[bsz,channels]*/, vc_stream_t stream);
int vc_linear_bias_f32(const float* x, const float* w, const float* b, float* y, int rows, int in_features, int out_features,
                       vc_stream_t stream);

/* ---- a7-a9  GPT-2 forward with KV cache (transformers GPT2LMHeadModel.forward) ----------- */
size_t vc_gpt_workspace_bytes(const VcGptWeights* w, int n_seq, int max_new_rows);
/* rows = n_seq * L new positions (L>=1), fp32 embeds [n_seq, L, H] WITHOUT position embedding;
 * past_len = positions already in the cache.  logits_out (fp32 [n_seq, vocab]) may be NULL;
 * next_ids (int32 [n_seq]) = argmax of the last position (ties -> lowest index) may be NULL. */
int vc_gpt2_forward(const VcGptWeights* w, const float* embeds, int n_seq, int L, int past_len, VcKvCache* cache,
                    void* workspace, size_t workspace_bytes, float* logits_out, int32_t* next_ids, vc_stream_t stream);
/* wte gather for fed-back tokens: out fp32 [n, H] */
int vc_gpt2_embed_tokens(const VcGptWeights* w, const int32_t* ids, int n, float* out, vc_stream_t stream);

/* ---- a8  greedy loop of core/scripts/benchmark_baseline.py:160-240, no host sync ---------- */
/* prefix fp32 [n_seq, P, H]; prompt_ids int32 [Lp] (shared by all rows); ids_out int32 [n_seq, max_new]
 * padded with eos; len_out int32 [n_seq]; forced_ids (teacher forcing, int32 [n_seq,max_new]) may be NULL;
 * step_logits (fp32 [max_new, n_seq, vocab]) may be NULL. */
int vc_greedy_decode(const VcGptWeights* w, const float* prefix, int n_seq, int P, const int32_t* prompt_ids, int Lp,
                     int max_new, int eos, VcKvCache* cache, void* workspace, size_t workspace_bytes, int32_t* ids_out,
                     int32_t* len_out, const int32_t* forced_ids, float* step_logits, vc_stream_t stream);

/* ---- decode-step building block: split-K weight-streaming GEMM for M <= 128 live sequences.
 *      partial fp32 [ksplit][M][N]; consumers sum the slices in order (ksplit 0 = library's choice) ---- */
int vc_skinny_ksplit(int N, int K);
int vc_skinny_gemm_partial(const void* x_bf16, const void* w_bf16, float* partial, int M, int N, int K, int ksplit, vc_stream_t stream);

/* ---- a10  HF generate on device: log-softmax + RepetitionPenalty/NoRepeatNGram/MinNewTokens processors +
 *      running beam scores + top-K continuations per video (K = 2*num_beams); raw_logits=1 is HF greedy
 *      (processors on raw logits, K=1).  seqs int32 [n_rows, max_len] = tokens generated so far. ---- */
int vc_beam_step(const float* logits, long long ld, int vocab, int n_rows, int rows_per_item, const int32_t* seqs, int max_len, int cur_len,
                 const float* running_scores, float repetition_penalty, int no_repeat_ngram, int min_new_tokens, int eos, int raw_logits,
                 int K, float* cand_score /*[n_rows,K]*/, int32_t* cand_tok /*[n_rows,K]*/, float* top_score /*[n_rows/rows_per_item,K]*/,
                 int32_t* top_idx /*flat beam*vocab+token*/, vc_stream_t stream);
/* KV-cache beam reorder as an index update: slot_out[r][p] = slot_in[src_rows[r]][p] for p < upto */
int vc_beam_reorder(const int32_t* slot_in, int32_t* slot_out, const int32_t* src_rows, int n_seq, int s_max, int upto, vc_stream_t stream);

/* Beam bookkeeping of transformers `_beam_search` on the device (running / finished hypotheses, early-stop heuristic), one small
 * kernel per step, so the whole beam loop is a fixed launch sequence (CUDA-graph capturable, no host round trip).  All arrays are
 * caller-owned device memory.  Per step: vc_beam_step (scores + top 2*nb continuations) -> vc_beam_update -> vc_beam_reorder with
 * src_rows -> next forward with next_tok.  After the last step: vc_beam_finalize. */
typedef struct {
  int32_t B, nb, max_len, eos;
  float* running_scores;   /* [B, nb]            (the `running_scores` input of vc_beam_step) */
  int32_t* running_seqs;   /* [B * nb, max_len]  (the `seqs` input of vc_beam_step) */
  int32_t* fin_seqs;       /* [B, nb, max_len] */
  float* fin_scores;       /* [B, nb] */
  int32_t* fin_done;       /* [B, nb] */
  int32_t* fin_len;        /* [B, nb] */
  int32_t* unsatisfied;    /* [B] */
  int32_t* flags;          /* [max_len + 1][2]: per step {some video unsatisfied, every candidate hit a stop} */
  int32_t* stopped;        /* [1]: HF's loop would have ended (the finished pool is frozen from then on) */
  int32_t* src_rows;       /* [B * nb] out: row each running beam continues from */
  int32_t* next_tok;       /* [B * nb] out: token each running beam feeds next */
} VcBeamState;
int vc_beam_init(const VcBeamState* st, vc_stream_t stream);
int vc_beam_update(const VcBeamState* st, const float* top_score, const int32_t* top_idx, int vocab, int cur_len, float length_penalty,
                   vc_stream_t stream);
int vc_beam_finalize(const VcBeamState* st, int32_t* ids_out /*[B,max_len] eos padded*/, int32_t* len_out /*[B]*/, vc_stream_t stream);

/* ---- token selection on given logits (bit-exact vs torch.argmax / topk) ------------------- */
int vc_argmax_f32(const float* logits, int rows, int vocab, int32_t* out, vc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VCB200_H */
