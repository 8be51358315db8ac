// C-ABI entry points (include/vcb200.h) and the host-side orchestration of the two
// hot loops: the ViT encoder over B*T frames (src/models/video_encoder.py:288-326)
// and the GPT-2 forward / greedy loop (core/scripts/benchmark_baseline.py:160-240).
// Everything is enqueued on the caller's stream; nothing here synchronises or allocates.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>

namespace vc {

// ---------------------------------------------------------------- error text
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------- launch accounting / profiling
static std::atomic<long long> g_launches{0};
static bool g_prof_on = false;
struct ProfRec { const char* name; double work; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;

KernelScope::KernelScope(const char* name, double work, cudaStream_t stream) : stream_(stream), slot_(-1) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  // no timing events inside a stream capture: they would become graph nodes, and cudaEventElapsedTime on them fails later
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool capturing = cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone;
  if (g_prof_on && !capturing) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r{name, work, nullptr, nullptr};
    if (cudaEventCreate(&r.a) == cudaSuccess && cudaEventCreate(&r.b) == cudaSuccess) {
      cudaEventRecord(r.a, stream);
      g_prof.push_back(r);
      slot_ = static_cast<int>(g_prof.size()) - 1;
    }
  }
}
KernelScope::~KernelScope() {
  if (slot_ >= 0) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof[slot_].b, stream_);
  }
}

static inline cudaStream_t S(vc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct VitBuffers {
  float* x; void* xn; void* qkv; void* att; void* hid; void* delta; void* delta2; float* x_cls; float* stats; float* pstats; size_t total;
};
static VitBuffers carve_vit(const VcVitWeights* w, int chunk_frames, void* base) {
  const size_t M = static_cast<size_t>(chunk_frames) * w->tokens;
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  VitBuffers b;
  b.x = reinterpret_cast<float*>(p + off);  off += align_up(M * w->dim * 4, 1024);
  b.xn = p + off;                           off += align_up(M * w->dim * 2, 1024);
  b.qkv = p + off;                          off += align_up(M * w->dim * 3 * 2, 1024);
  b.att = p + off;                          off += align_up(M * w->dim * 2, 1024);
  b.hid = p + off;                          off += align_up(M * w->mlp * 2, 1024);
  b.delta = p + off;                        off += align_up(M * w->dim * 2, 1024);
  b.delta2 = p + off;                       off += align_up(M * w->dim * 2, 1024);
  b.x_cls = reinterpret_cast<float*>(p + off); off += align_up(static_cast<size_t>(chunk_frames) * w->dim * 4, 1024);
  b.stats = reinterpret_cast<float*>(p + off); off += align_up(M * 8, 1024);                                     // (mean, rstd) per row
  b.pstats = reinterpret_cast<float*>(p + off); off += align_up(M * 8 * gemm_resid_parts(w->dim), 1024);      // partial (sum, sum sq)
  b.total = off;
  return b;
}

struct GptBuffers {
  float* h; void* xn; void* qkv; void* att; void* hid; float* emb; float* logits; int32_t* finished; int32_t* next; float* partial;
  float* stat; float* cand_v; int* cand_i; float* pstats; unsigned int* done;
  size_t total;
};
constexpr int kDecodeMaxRows = 1024;  // n_seq * new positions handled by the weight-streaming kernels; beyond: tcgen05 GEMM path
static GptBuffers carve_gpt(const VcGptWeights* w, int n_seq, int max_rows, void* base) {
  const size_t R = static_cast<size_t>(max_rows);
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  GptBuffers b;
  b.h = reinterpret_cast<float*>(p + off);      off += align_up(R * w->dim * 4, 1024);
  b.xn = p + off;                               off += align_up(R * w->dim * 2, 1024);
  b.qkv = p + off;                              off += align_up(R * w->dim * 3 * 2, 1024);
  b.att = p + off;                              off += align_up(R * w->dim * 2, 1024);
  b.hid = p + off;                              off += align_up(R * w->dim * 4 * 2, 1024);
  b.emb = reinterpret_cast<float*>(p + off);    off += align_up(R * w->dim * 4, 1024);
  b.logits = reinterpret_cast<float*>(p + off); off += align_up(static_cast<size_t>(n_seq) * w->vocab_pad * 4, 1024);
  b.finished = reinterpret_cast<int32_t*>(p + off); off += align_up(static_cast<size_t>(n_seq) * 4, 1024);
  b.next = reinterpret_cast<int32_t*>(p + off);     off += align_up(static_cast<size_t>(n_seq) * 4, 1024);
  {
    // fp32 split-K partials of the fallback chain's GEMMs (widest product); partial LayerNorm statistics [16][rows] and the
    // lm_head's per-CTA argmax candidates [n_seq][ctas] of the decode chain
    const size_t rows = R < static_cast<size_t>(kDecodeMaxRows) ? R : static_cast<size_t>(kDecodeMaxRows);
    const int H = w->dim;
    size_t m = static_cast<size_t>(skinny_ksplit(3 * H, H)) * 3 * H;
    m = std::max(m, static_cast<size_t>(skinny_ksplit(H, H)) * H);
    m = std::max(m, static_cast<size_t>(skinny_ksplit(4 * H, H)) * 4 * H);
    m = std::max(m, static_cast<size_t>(skinny_ksplit(H, 4 * H)) * H);
    b.partial = reinterpret_cast<float*>(p + off);  off += align_up(m * rows * 4, 1024);
    b.stat = reinterpret_cast<float*>(p + off);     off += align_up(static_cast<size_t>(16) * R * 8, 1024);
    const size_t nc = static_cast<size_t>(n_seq) * 256;
    b.cand_v = reinterpret_cast<float*>(p + off);   off += align_up(nc * 4, 1024);
    b.cand_i = reinterpret_cast<int*>(p + off);     off += align_up(nc * 4, 1024);
    // partial row statistics of the tcgen05 chain's residual epilogues (all rows: that chain also serves > 1024-row prefills)
    b.pstats = reinterpret_cast<float*>(p + off);   off += align_up(static_cast<size_t>(gemm_resid_parts(H)) * R * 8, 1024);
    b.done = reinterpret_cast<unsigned int*>(p + off); off += 1024;
  }
  b.total = off;
  return b;
}
static ChainBuffers chain_buffers(const GptBuffers& b) { return ChainBuffers{b.h, b.xn, b.qkv, b.att, b.hid, b.stat, b.cand_v, b.cand_i}; }
// VC_DECODE_CHAIN=1 (A/B switch): the round-1 split-K kernel chain instead of decode_chain.cu
static bool use_chain_v2(const VcGptWeights* w, int rows) {
  static const bool v1 = getenv("VC_DECODE_CHAIN") != nullptr && atoi(getenv("VC_DECODE_CHAIN")) == 1;
  return !v1 && chain_supported(w, rows) && chain_lmhead_ctas() <= 256;
}

}  // namespace vc

using namespace vc;

extern "C" {

const char* vc_last_error(void) { return g_err; }
int vc_abi_version(void) { return 5; }
int vc_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}
long long vc_launch_count(void) { return g_launches.load(); }
int vc_debug_trace(void* buf_u64, int max_records) { return chain_set_trace(buf_u64, max_records); }

int vc_prof_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = true;
  return 0;
}
int vc_prof_end(int max_rows, char* names, float* total_ms, int* calls, double* work) {
  VC_CUDA_OK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  std::map<std::string, int> index;
  int n = 0;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { ms = 0.f; (void)cudaGetLastError(); }   // never leave an error behind
    auto it = index.find(r.name);
    int i;
    if (it == index.end()) {
      if (n >= max_rows) continue;
      i = n++;
      index[r.name] = i;
      std::snprintf(names + i * 48, 48, "%s", r.name);
      total_ms[i] = 0.f; calls[i] = 0; work[i] = 0.0;
    } else {
      i = it->second;
    }
    total_ms[i] += ms; calls[i] += 1; work[i] += r.work;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return n;
}

int vc_resize_bilinear_u8(const uint8_t* src_hwc, int n_frames, int in_h, int in_w, uint8_t* scratch, uint8_t* dst_hwc, int out_h, int out_w,
                          const int32_t* kx, const int32_t* bounds_x, int ksize_x, const int32_t* ky, const int32_t* bounds_y, int ksize_y,
                          vc_stream_t stream) {
  return resize_bilinear_u8(src_hwc, n_frames, in_h, in_w, scratch, dst_hwc, out_h, out_w, kx, bounds_x, ksize_x, ky, bounds_y, ksize_y, S(stream));
}
int vc_preprocess_u8(const uint8_t* frames_hwc, const float* lut3x256, void* out_bf16, int n_frames, int H, int W, int layout,
                     int patch, int k_pad, vc_stream_t stream) {
  return preprocess_u8(frames_hwc, lut3x256, out_bf16, n_frames, H, W, layout, patch, k_pad, S(stream));
}

int vc_patchify_f32(const float* video_chw, void* out_bf16, int n_frames, int H, int W, int patch, int k_pad, vc_stream_t stream) {
  return patchify_f32(video_chw, out_bf16, n_frames, H, W, patch, k_pad, S(stream));
}

int vc_gemm_bf16(const void* A, const void* W, const float* bias, int M, int N, int K, int epilogue, void* out, int ldo,
                 const float* aux, int rows_per_group, vc_stream_t stream) {
  return gemm_bf16(A, W, bias, M, N, K, epilogue, out, ldo, aux, rows_per_group, 0, S(stream));
}

int vc_gemm_resid_stats(const void* A, const void* W, const float* bias, int M, int N, int K, float* x, void* xb_bf16, float* pstats,
                        float* stats_out, unsigned int* done, float eps, int split_k, vc_stream_t stream) {
  VC_REQUIRE(x != nullptr && xb_bf16 != nullptr && pstats != nullptr && bias != nullptr, "gemm_resid_stats: null operand");
  VC_REQUIRE((stats_out == nullptr) == (done == nullptr), "gemm_resid_stats: stats_out and done go together");
  VC_REQUIRE(split_k == 0 || split_k == 1 || split_k == 2 || split_k == 4, "gemm_resid_stats: split_k=%d", split_k);
  GemmExtra ex{nullptr, nullptr, x, xb_bf16, pstats, stats_out, done, eps, split_k};
  return gemm_bf16_ex(A, W, bias, M, N, K, VC_EPI_RESID_STATS, xb_bf16, N, nullptr, 0, 0, &ex, S(stream));
}

int vc_layernorm_f32_bf16(const float* x, const float* gamma, const float* beta, void* out_bf16, int rows, int dim, float eps,
                          vc_stream_t stream) {
  return layernorm_f32_bf16(x, gamma, beta, out_bf16, rows, dim, eps, S(stream));
}

int vc_vit_attention(const void* qkv_bf16, void* out_bf16, int n_frames, int tokens, int heads, int head_dim, vc_stream_t stream) {
  return vit_attention(qkv_bf16, out_bf16, n_frames, tokens, heads, head_dim, S(stream));
}

int vc_vit_attention_mma_sync(const void* qkv_bf16, void* out_bf16, int n_frames, int tokens, int heads, int head_dim, vc_stream_t stream) {
  return vit_attention_mma_sync(qkv_bf16, out_bf16, n_frames, tokens, heads, head_dim, S(stream));
}

size_t vc_vit_workspace_bytes(const VcVitWeights* w, int chunk_frames) { return carve_vit(w, chunk_frames, nullptr).total; }

int vc_vit_encode(const VcVitWeights* w, const void* patches_bf16, int n_frames, int chunk_frames, void* workspace,
                  size_t workspace_bytes, float* cls_out, vc_stream_t stream) {
  VC_REQUIRE(w != nullptr && w->layer != nullptr, "vit_encode: null weights");
  VC_REQUIRE(chunk_frames > 0, "vit_encode: chunk_frames=%d", chunk_frames);
  VC_REQUIRE(w->dim == w->heads * 64, "vit_encode: dim=%d heads=%d (head_dim must be 64)", w->dim, w->heads);
  if (n_frames <= 0) return 0;
  VitBuffers b = carve_vit(w, chunk_frames, workspace);
  VC_REQUIRE(workspace != nullptr && workspace_bytes >= b.total, "vit_encode: workspace %zu < %zu bytes", workspace_bytes, b.total);
  cudaStream_t s = S(stream);
  const int D = w->dim, N = w->tokens, P = w->tokens - 1;
  const int gelu = w->gelu_tanh ? VC_EPI_BIAS_GELU_TANH : VC_EPI_BIAS_GELU_ERF;
  for (int f0 = 0; f0 < n_frames; f0 += chunk_frames) {
    const int nf = (n_frames - f0 < chunk_frames) ? (n_frames - f0) : chunk_frames;
    const int M = nf * N;
    const __nv_bfloat16* patches = static_cast<const __nv_bfloat16*>(patches_bf16) + static_cast<size_t>(f0) * P * w->patch_k;
    int e;
    // tokens: [cls + pos0 | conv_proj(patches) + bias + pos]   (video_encoder.py:90-95, Encoder.forward)
    if ((e = cls_rows_init(b.x, w->cls_pos0, nf, N, D, s))) return e;
    if ((e = gemm_bf16(patches, w->patch_w, w->patch_b, nf * P, D, w->patch_k, VC_EPI_PATCH_EMBED, b.x, D, w->pos, P, 0, s))) return e;
    bool folded = getenv("VC_VIT_UNFOLDED") == nullptr;      // A/B switch: the stand-alone LayerNorm passes of round 1
    for (int l = 0; l < w->layers; ++l) folded = folded && w->layer[l].qkv_wf != nullptr && w->layer[l].fc1_wf != nullptr;
    if (folded) {
      // No LayerNorm pass and no separate residual add.  The residual stream x (fp32) is read-modify-written by the proj and
      // fc2 epilogues (x += bf16(acc + bias)); both emit bf16(x_new) — the A operand of the next product, whose weights
      // carry the LayerNorm affine (packing.fold_layernorm) — and partial row statistics; (mean, rstd) are applied in the
      // consumer's epilogue.  Only the first block's input needs a pass of its own (rows come from two producers).
      const int parts = gemm_resid_parts(D);       // one partial-statistics slot per row and 32-column chunk
      const int gelu_f = w->gelu_tanh ? VC_EPI_LNF_GELU_TANH : VC_EPI_LNF_GELU_ERF;
      if ((e = rowstats_cast(b.x, b.xn, b.stats, M, D, 1e-6f, s))) return e;
      for (int l = 0; l < w->layers; ++l) {
        const VcVitLayer& L = w->layer[l];
        GemmExtra ex{L.qkv_cs, b.stats, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0};
        if ((e = gemm_bf16_ex(b.xn, L.qkv_wf, L.qkv_bf, M, 3 * D, D, VC_EPI_LNF_BIAS, b.qkv, 3 * D, nullptr, 0, 0, &ex, s))) return e;
        if (l + 1 == w->layers) break;
        if ((e = vit_attention(b.qkv, b.att, nf, N, w->heads, 64, s))) return e;
        GemmExtra ep{nullptr, nullptr, b.x, b.xn, b.pstats, nullptr, nullptr, 0.f, 0};
        if ((e = gemm_bf16_ex(b.att, L.proj_w, L.proj_b, M, D, D, VC_EPI_RESID_STATS, nullptr, D, nullptr, 0, 0, &ep, s))) return e;
        if ((e = ln_stats_finalize(b.pstats, parts, M, D, 1e-6f, b.stats, s))) return e;
        GemmExtra e1{L.fc1_cs, b.stats, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0};
        if ((e = gemm_bf16_ex(b.xn, L.fc1_wf, L.fc1_bf, M, w->mlp, D, gelu_f, b.hid, w->mlp, nullptr, 0, 0, &e1, s))) return e;
        if ((e = gemm_bf16_ex(b.hid, L.fc2_w, L.fc2_b, M, D, w->mlp, VC_EPI_RESID_STATS, nullptr, D, nullptr, 0, 0, &ep, s))) return e;
        if ((e = ln_stats_finalize(b.pstats, parts, M, D, 1e-6f, b.stats, s))) return e;
      }
    } else {
    // The bias-added outputs of proj / fc2 go to `delta` / `delta2` in bf16 (write-only epilogues); the LayerNorm kernels
    // fold them into the fp32 residual stream, which is stored once per block (by the next block's LN1).
    bool pending = false;
    for (int l = 0; l < w->layers; ++l) {
      const VcVitLayer& L = w->layer[l];
      const bool last = l + 1 == w->layers;
      if ((e = add_layernorm_rows(b.x, pending ? b.delta : nullptr, pending ? b.delta2 : nullptr, 1, 1, 0, L.ln1_g, L.ln1_b, nullptr, b.xn, M, D,
                                  1e-6f, s)))
        return e;
      if ((e = gemm_bf16(b.xn, L.qkv_w, L.qkv_b, M, 3 * D, D, VC_EPI_BIAS, b.qkv, 3 * D, nullptr, 0, 0, s))) return e;
      if (last) break;
      if ((e = vit_attention(b.qkv, b.att, nf, N, w->heads, 64, s))) return e;
      if ((e = gemm_bf16(b.att, L.proj_w, L.proj_b, M, D, D, VC_EPI_BIAS, b.delta, D, nullptr, 0, 0, s))) return e;
      if ((e = add_layernorm_rows(b.x, b.delta, nullptr, 0, 1, 0, L.ln2_g, L.ln2_b, nullptr, b.xn, M, D, 1e-6f, s))) return e;
      if ((e = gemm_bf16(b.xn, L.fc1_w, L.fc1_b, M, w->mlp, D, gelu, b.hid, w->mlp, nullptr, 0, 0, s))) return e;
      if ((e = gemm_bf16(b.hid, L.fc2_w, L.fc2_b, M, D, w->mlp, VC_EPI_BIAS, b.delta2, D, nullptr, 0, 0, s))) return e;
      pending = true;
    }
    }
    // LAST block: only the class token of each frame is consumed downstream (video_encoder.py:256-258), and every op after
    // the attention is row-wise, so attention runs for the class-token query alone and proj / LN2 / MLP / final LN run on
    // nf rows instead of nf*N.  Identical arithmetic per row; ~6 % less encoder work (bench.py subtracts it from the FLOPs).
    {
      const VcVitLayer& L = w->layer[w->layers - 1];
      VC_CUDA_OK(cudaMemcpy2DAsync(b.x_cls, static_cast<size_t>(D) * 4, b.x, static_cast<size_t>(N) * D * 4, static_cast<size_t>(D) * 4, nf,
                                   cudaMemcpyDeviceToDevice, s));
      if ((e = vit_cls_attention(b.qkv, b.att, nf, N, w->heads, 64, s))) return e;
      if ((e = gemm_bf16(b.att, L.proj_w, L.proj_b, nf, D, D, VC_EPI_BIAS, b.delta, D, nullptr, 0, 0, s))) return e;
      if ((e = add_layernorm_rows(b.x_cls, b.delta, nullptr, 0, 1, 0, L.ln2_g, L.ln2_b, nullptr, b.xn, nf, D, 1e-6f, s))) return e;
      if ((e = gemm_bf16(b.xn, L.fc1_w, L.fc1_b, nf, w->mlp, D, gelu, b.hid, w->mlp, nullptr, 0, 0, s))) return e;
      if ((e = gemm_bf16(b.hid, L.fc2_w, L.fc2_b, nf, D, w->mlp, VC_EPI_BIAS, b.delta2, D, nullptr, 0, 0, s))) return e;
      if ((e = add_layernorm_rows(b.x_cls, b.delta, b.delta2, 0, 1, 0, w->lnf_g, w->lnf_b, cls_out + static_cast<size_t>(f0) * D, nullptr, nf, D,
                                  1e-6f, s)))
        return e;
    }
  }
  return 0;
}

int vc_pool_prefix(const float* cls_tokens, int B, int T, int dim, const float* head_w, const float* head_b, int video_dim,
                   float ln_scale, float in_weight, const float* mapper_w, const float* mapper_b, int mapper_out, float* feat_out,
                   float* prefix_out, vc_stream_t stream) {
  return pool_prefix(cls_tokens, B, T, dim, head_w, head_b, video_dim, ln_scale, in_weight, mapper_w, mapper_b, mapper_out, feat_out,
                     prefix_out, S(stream));
}
int vc_vit_pool_temporal(const void* feat, int is_bf16, int bsz, int timesteps, int tokens, int channels, int gap, float* out,
                         vc_stream_t stream) {
  return vit_pool_temporal(feat, is_bf16, bsz, timesteps, tokens, channels, gap, out, S(stream));
}
int vc_linear_bias_f32(const float* x, const float* w, const float* b, float* y, int rows, int in_features, int out_features,
                       vc_stream_t stream) {
  return linear_bias_f32(x, w, b, y, rows, in_features, out_features, S(stream));
}

size_t vc_gpt_workspace_bytes(const VcGptWeights* w, int n_seq, int max_new_rows) { return carve_gpt(w, n_seq, max_new_rows, nullptr).total; }

// Few-row forward (a decode step, or the prefill of a caption batch: n_seq * L <= 1024 rows) as a chain of small kernels:
// every product streams its weights once through the split-K skinny kernel, partial sums are folded into the next kernel of
// the chain, and every kernel is a programmatic dependent launch so its launch latency and ramp overlap the previous
// kernel's tail.  None of these kernels needs a whole SM, so the chain runs beside the resident CTAs of the encoder GEMM
// of the next batch (CaptionPipeline); the tcgen05 GEMM path below would wait for a free SM at every product.
static int gpt_step_skinny(const VcGptWeights* w, const float* embeds, int n_seq, int L, int past_len, VcKvCache* cache, const GptBuffers& b,
                           float* logits_out, cudaStream_t s) {
  const int H = w->dim, M = n_seq * L;
  const int mt = (M + 63) / 64;
  const int ks_attn = skinny_ksplit(3 * H, H, mt), ks_ap = skinny_ksplit(H, H, mt), ks_fc = skinny_ksplit(4 * H, H, mt), ks_mp = skinny_ksplit(H, 4 * H, mt);
  // VC_DECODE_FUSED_FC=1 (A/B switch, read per call): fc1 with bias + gelu_new fused, K split inside the CTA.  Measured equal
  // to the split-K product + bias_act pair (10.30 vs 10.13 ms per 20 tokens): what the saved launch gains, the whole-K
  // activation block every CTA then pulls through L1 costs.
  const bool fused_fc = H % 256 == 0 && getenv("VC_DECODE_FUSED_FC") != nullptr;
  int e;
  if (H % 128 == 0 && H <= 1024) {
    if ((e = gpt_add_pos_ln(embeds, w->wpe, b.h, b.xn, w->layer[0].ln1_g, w->layer[0].ln1_b, n_seq, L, past_len, H, 1e-5f, s))) return e;
  } else {
    if ((e = gpt_add_pos(embeds, w->wpe, b.h, n_seq, L, past_len, H, s))) return e;
    if ((e = layernorm_f32_bf16(b.h, w->layer[0].ln1_g, w->layer[0].ln1_b, b.xn, M, H, 1e-5f, s))) return e;
  }
  for (int l = 0; l < w->layers; ++l) {
    const VcGptLayer& Ly = w->layer[l];
    if ((e = skinny_gemm(b.xn, Ly.attn_w, b.partial, M, 3 * H, H, ks_attn, s))) return e;
    if ((e = gpt_attention(nullptr, b.partial, ks_attn, Ly.attn_b, b.att, cache, l, n_seq, L, past_len, s))) return e;
    if ((e = skinny_gemm(b.att, Ly.aproj_w, b.partial, M, H, H, ks_ap, s))) return e;
    if ((e = resid_ln(b.h, b.partial, ks_ap, Ly.aproj_b, Ly.ln2_g, Ly.ln2_b, b.xn, M, H, 1e-5f, s))) return e;
    if (fused_fc) {
      if ((e = skinny_gemm_gelu(b.xn, Ly.fc_w, Ly.fc_b, b.hid, M, 4 * H, H, s))) return e;
    } else {
      if ((e = skinny_gemm(b.xn, Ly.fc_w, b.partial, M, 4 * H, H, ks_fc, s))) return e;
      if ((e = bias_act(b.partial, ks_fc, Ly.fc_b, b.hid, M, 4 * H, 1, s))) return e;
    }
    if ((e = skinny_gemm(b.hid, Ly.mproj_w, b.partial, M, H, 4 * H, ks_mp, s))) return e;
    const bool last = l + 1 == w->layers;
    const float* ng = last ? w->lnf_g : w->layer[l + 1].ln1_g;
    const float* nb = last ? w->lnf_b : w->layer[l + 1].ln1_b;
    if ((e = resid_ln(b.h, b.partial, ks_mp, Ly.mproj_b, ng, nb, b.xn, M, H, 1e-5f, s))) return e;
  }
  // ln_f + tied lm_head on the last position of every sequence only (HF computes all positions; unused).  One K slice, so
  // the "partial" is the logits row itself.
  if (L > 1 && (e = layernorm_rows(b.h, L, L - 1, w->lnf_g, w->lnf_b, nullptr, b.xn, n_seq, H, 1e-5f, s))) return e;
  return skinny_gemm(b.xn, w->wte, logits_out, n_seq, w->vocab_pad, H, 1, s);
}

// Forwards of many rows (the prefill of a caption batch, grouped decode steps, beam steps: 128-1280+ rows) on the tcgen05 GEMM with
// 64-column tiles: QKV (LayerNorm folded) -> attention -> proj (residual + statistics in the epilogue) -> finalize -> fc1 (folded,
// gelu_new) -> fc2 (residual + statistics) -> finalize: the encoder's scheme on GPT-2's folded weights; every
// launch is a programmatic dependent launch.  lm_head: ln_f on the last rows + tied lm_head GEMM (fp32 logits).
static bool use_gemm_chain(const VcGptWeights* w, int rows) {
  static const int min_rows = getenv("VC_GEMM_ROWS") != nullptr ? atoi(getenv("VC_GEMM_ROWS")) : 128;
  if (min_rows <= 0 || rows < min_rows) return false;
  if (w->dim % 64 != 0 || w->vocab_pad % 32 != 0) return false;
  for (int l = 0; l < w->layers; ++l)
    if (w->layer[l].attn_wf == nullptr || w->layer[l].fc_wf == nullptr) return false;
  return true;
}
static int gpt_forward_gemm_chain(const VcGptWeights* w, const float* embeds, int n_seq, int L, int past_len, VcKvCache* cache,
                                  const GptBuffers& b, float* logits_out, cudaStream_t s) {
  const int H = w->dim, M = n_seq * L;
  int e;
  VC_CUDA_OK(cudaMemsetAsync(b.done, 0, sizeof(unsigned int), s));     // the residual epilogues' "CTAs done" counter
  if ((e = gpt_add_pos(embeds, w->wpe, b.h, n_seq, L, past_len, H, s))) return e;
  if ((e = rowstats_cast(b.h, b.xn, b.stat, M, H, 1e-5f, s))) return e;
  // residual update + statistics; the last CTA to finish also turns the partials into (mean, rstd): no finalize kernel.
  // (Tried: the CONSUMER's epilogue warps combining the partials while its MMAs run — the producer loses its 3.5 us tail, the
  // consumer gains 4.3 us of scattered 8-byte loads in all 16 epilogue warps: 582 us per step either way.)
  GemmExtra er{nullptr, nullptr, b.h, b.xn, b.pstats, b.stat, b.done, 1e-5f, 0};
  for (int l = 0; l < w->layers; ++l) {
    const VcGptLayer& Ly = w->layer[l];
    GemmExtra eq{Ly.attn_cs, b.stat, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0};
    if ((e = gemm_bf16_ex(b.xn, Ly.attn_wf, Ly.attn_bf, M, 3 * H, H, VC_EPI_LNF_BIAS, b.qkv, 3 * H, nullptr, 0, 0, &eq, s))) return e;
    if ((e = gpt_attention(b.qkv, nullptr, 0, nullptr, b.att, cache, l, n_seq, L, past_len, s))) return e;
    if ((e = gemm_bf16_ex(b.att, Ly.aproj_w, Ly.aproj_b, M, H, H, VC_EPI_RESID_STATS, nullptr, H, nullptr, 0, 0, &er, s))) return e;
    GemmExtra ef{Ly.fc_cs, b.stat, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0};
    if ((e = gemm_bf16_ex(b.xn, Ly.fc_wf, Ly.fc_bf, M, 4 * H, H, VC_EPI_LNF_GELU_TANH, b.hid, 4 * H, nullptr, 0, 0, &ef, s))) return e;
    if ((e = gemm_bf16_ex(b.hid, Ly.mproj_w, Ly.mproj_b, M, H, 4 * H, VC_EPI_RESID_STATS, nullptr, H, nullptr, 0, 0, &er, s))) return e;
  }
  if ((e = layernorm_rows(b.h, L, L - 1, w->lnf_g, w->lnf_b, nullptr, b.xn, n_seq, H, 1e-5f, s))) return e;
  return gemm_bf16(b.xn, w->wte, nullptr, n_seq, w->vocab_pad, H, VC_EPI_BIAS_F32, logits_out, w->vocab_pad, nullptr, 0, 0, s);
}

static int gpt_forward_impl(const VcGptWeights* w, const float* embeds, int n_seq, int L, int past_len, VcKvCache* cache,
                            const GptBuffers& b, float* logits_out, cudaStream_t s) {
  const int H = w->dim, M = n_seq * L;
  int e;
  if (use_chain_v2(w, M)) {
    const ChainBuffers cb = chain_buffers(b);
    if ((e = chain_add_pos_stats(w, embeds, cb, n_seq, L, past_len, s))) return e;
    return chain_layers(w, cb, n_seq, L, past_len, cache, logits_out, w->vocab_pad, s);
  }
  // forwards of >= 128 rows (VC_GEMM_ROWS=n moves the threshold, 0 switches it off): the tcgen05 chain above
  if (use_gemm_chain(w, M)) return gpt_forward_gemm_chain(w, embeds, n_seq, L, past_len, cache, b, logits_out, s);
  // VC_PREFILL_TCGEN05=1 (A/B switch, read per call): multi-position forwards take the plain tcgen05 GEMM path as in round 1
  const bool chain_ok = L == 1 ? n_seq <= 256 || getenv("VC_PREFILL_TCGEN05") == nullptr : getenv("VC_PREFILL_TCGEN05") == nullptr;
  if (chain_ok && M <= kDecodeMaxRows && w->vocab_pad % 64 == 0) return gpt_step_skinny(w, embeds, n_seq, L, past_len, cache, b, logits_out, s);
  if ((e = gpt_add_pos(embeds, w->wpe, b.h, n_seq, L, past_len, H, s))) return e;
  for (int l = 0; l < w->layers; ++l) {
    const VcGptLayer& Ly = w->layer[l];
    if ((e = layernorm_f32_bf16(b.h, Ly.ln1_g, Ly.ln1_b, b.xn, M, H, 1e-5f, s))) return e;
    if ((e = gemm_bf16(b.xn, Ly.attn_w, Ly.attn_b, M, 3 * H, H, VC_EPI_BIAS, b.qkv, 3 * H, nullptr, 0, 0, s))) return e;
    if ((e = gpt_attention(b.qkv, nullptr, 0, nullptr, b.att, cache, l, n_seq, L, past_len, s))) return e;
    if ((e = gemm_bf16(b.att, Ly.aproj_w, Ly.aproj_b, M, H, H, VC_EPI_BIAS_RESID_F32, b.h, H, nullptr, 0, 0, s))) return e;
    if ((e = layernorm_f32_bf16(b.h, Ly.ln2_g, Ly.ln2_b, b.xn, M, H, 1e-5f, s))) return e;
    if ((e = gemm_bf16(b.xn, Ly.fc_w, Ly.fc_b, M, 4 * H, H, VC_EPI_BIAS_GELU_TANH, b.hid, 4 * H, nullptr, 0, 0, s))) return e;
    if ((e = gemm_bf16(b.hid, Ly.mproj_w, Ly.mproj_b, M, H, 4 * H, VC_EPI_BIAS_RESID_F32, b.h, H, nullptr, 0, 0, s))) return e;
  }
  // ln_f + tied lm_head on the last position of every row only (HF computes all positions; unused)
  if ((e = layernorm_rows(b.h, L, L - 1, w->lnf_g, w->lnf_b, nullptr, b.xn, n_seq, H, 1e-5f, s))) return e;
  if ((e = gemm_bf16(b.xn, w->wte, nullptr, n_seq, w->vocab_pad, H, VC_EPI_BIAS_F32, logits_out, w->vocab_pad, nullptr, 0, 0, s))) return e;
  return 0;
}

int vc_gpt2_forward(const VcGptWeights* w, const float* embeds, int n_seq, int L, int past_len, VcKvCache* cache, void* workspace,
                    size_t workspace_bytes, float* logits_out, int32_t* next_ids, vc_stream_t stream) {
  VC_REQUIRE(w != nullptr && w->layer != nullptr && cache != nullptr, "gpt2_forward: null argument");
  VC_REQUIRE(n_seq > 0 && L > 0 && past_len >= 0, "gpt2_forward: n_seq=%d L=%d past_len=%d", n_seq, L, past_len);
  VC_REQUIRE(past_len + L <= w->n_pos, "gpt2_forward: position %d exceeds n_positions=%d", past_len + L, w->n_pos);
  GptBuffers b = carve_gpt(w, n_seq, n_seq * L, workspace);
  VC_REQUIRE(workspace != nullptr && workspace_bytes >= b.total, "gpt2_forward: workspace %zu < %zu bytes", workspace_bytes, b.total);
  if (use_chain_v2(w, n_seq * L)) {
    // logits are materialised only when the caller asks for them; next_ids come from the lm_head's per-CTA candidates
    int e = gpt_forward_impl(w, embeds, n_seq, L, past_len, cache, b, logits_out, S(stream));
    if (e) return e;
    if (next_ids) return chain_argmax(chain_buffers(b), n_seq, next_ids, S(stream));
    return 0;
  }
  float* lg = logits_out ? logits_out : b.logits;
  int e = gpt_forward_impl(w, embeds, n_seq, L, past_len, cache, b, lg, S(stream));
  if (e) return e;
  if (next_ids) return argmax_f32(lg, w->vocab_pad, n_seq, w->vocab, next_ids, S(stream));
  return 0;
}

int vc_gpt2_embed_tokens(const VcGptWeights* w, const int32_t* ids, int n, float* out, vc_stream_t stream) {
  return embed_tokens(w->wte, ids, n, w->dim, out, S(stream));
}

int vc_greedy_decode(const VcGptWeights* w, const float* prefix, int n_seq, int P, const int32_t* prompt_ids, int Lp, int max_new,
                     int eos, VcKvCache* cache, void* workspace, size_t workspace_bytes, int32_t* ids_out, int32_t* len_out,
                     const int32_t* forced_ids, float* step_logits, vc_stream_t stream) {
  VC_REQUIRE(w != nullptr && w->layer != nullptr && cache != nullptr, "greedy_decode: null argument");
  VC_REQUIRE(n_seq > 0 && P >= 0 && Lp > 0 && max_new > 0, "greedy_decode: n_seq=%d P=%d Lp=%d max_new=%d", n_seq, P, Lp, max_new);
  const int L0 = P + Lp;
  VC_REQUIRE(L0 + max_new - 1 <= cache->s_max, "greedy_decode: %d positions exceed cache s_max=%d", L0 + max_new - 1, cache->s_max);
  GptBuffers b = carve_gpt(w, n_seq, n_seq * L0, workspace);
  VC_REQUIRE(workspace != nullptr && workspace_bytes >= b.total, "greedy_decode: workspace %zu < %zu bytes", workspace_bytes, b.total);
  cudaStream_t s = S(stream);
  int e;
  if ((e = greedy_init(ids_out, len_out, b.finished, n_seq, max_new, eos, s))) return e;
  if ((e = build_prefill_embeds(prefix, w->wte, prompt_ids, n_seq, P, Lp, w->dim, b.emb, s))) return e;
  // Per forward the faster of the two few-row chains is taken (decode_chain.cu up to 64 rows, the split-K chain beyond: the
  // prefill of 64 captions is 320 rows).  decode_chain.cu's selection kernel writes the next step's input rows itself
  // (wte[token] + wpe, bf16 copy, row statistics); after any other forward they are built from b.emb.
  const ChainBuffers cb = chain_buffers(b);
  bool rows_ready = false;
  for (int step = 0; step < max_new; ++step) {
    const int L = step == 0 ? L0 : 1;
    const int past = step == 0 ? 0 : L0 + step - 1;
    const bool last = step == max_new - 1;
    if (use_chain_v2(w, n_seq * L)) {
      float* lg = step_logits ? step_logits + static_cast<size_t>(step) * n_seq * w->vocab_pad : nullptr;
      if (!rows_ready && (e = chain_add_pos_stats(w, b.emb, cb, n_seq, L, past, s))) return e;
      if ((e = chain_layers(w, cb, n_seq, L, past, cache, lg, w->vocab_pad, s))) return e;
      const bool next_chain = !last && use_chain_v2(w, n_seq);
      if ((e = chain_select(w, cb, n_seq, step, max_new, eos, b.finished, ids_out, len_out, forced_ids, L0 + step, next_chain, b.next, s))) return e;
      rows_ready = next_chain;
      continue;
    }
    float* lg = step_logits ? step_logits + static_cast<size_t>(step) * n_seq * w->vocab_pad : b.logits;
    if ((e = gpt_forward_impl(w, b.emb, n_seq, L, past, cache, b, lg, s))) return e;
    if ((e = greedy_select(lg, w->vocab_pad, w->vocab, n_seq, step, max_new, eos, b.finished, ids_out, len_out, forced_ids, w->wte,
                           w->dim, last ? nullptr : b.emb, b.next, s)))
      return e;
    rows_ready = false;
  }
  return 0;
}

int vc_beam_step(const float* logits, long long ld, int vocab, int n_rows, int rows_per_item, const int32_t* seqs, int max_len, int cur_len,
                 const float* running_scores, float repetition_penalty, int no_repeat_ngram, int min_new_tokens, int eos, int raw_logits,
                 int K, float* cand_score, int32_t* cand_tok, float* top_score, int32_t* top_idx, vc_stream_t stream) {
  return beam_step(logits, ld, vocab, n_rows, rows_per_item, seqs, max_len, cur_len, running_scores, repetition_penalty, no_repeat_ngram,
                   min_new_tokens, eos, raw_logits, K, cand_score, cand_tok, top_score, top_idx, S(stream));
}

int vc_beam_init(const VcBeamState* st, vc_stream_t stream) { return beam_init(st, S(stream)); }
int vc_beam_update(const VcBeamState* st, const float* top_score, const int32_t* top_idx, int vocab, int cur_len, float length_penalty,
                   vc_stream_t stream) {
  return beam_update(st, top_score, top_idx, vocab, cur_len, length_penalty, S(stream));
}
int vc_beam_finalize(const VcBeamState* st, int32_t* ids_out, int32_t* len_out, vc_stream_t stream) {
  return beam_finalize(st, ids_out, len_out, S(stream));
}

int vc_beam_reorder(const int32_t* slot_in, int32_t* slot_out, const int32_t* src_rows, int n_seq, int s_max, int upto, vc_stream_t stream) {
  return beam_reorder(slot_in, slot_out, src_rows, n_seq, s_max, upto, S(stream));
}

int vc_skinny_gemm_partial(const void* x_bf16, const void* w_bf16, float* partial, int M, int N, int K, int ksplit, vc_stream_t stream) {
  return skinny_gemm(x_bf16, w_bf16, partial, M, N, K, ksplit > 0 ? ksplit : skinny_ksplit(N, K), S(stream));
}
int vc_skinny_ksplit(int N, int K) { return skinny_ksplit(N, K); }

int vc_argmax_f32(const float* logits, int rows, int vocab, int32_t* out, vc_stream_t stream) {
  return argmax_f32(logits, vocab, rows, vocab, out, S(stream));
}

}  // extern "C"
