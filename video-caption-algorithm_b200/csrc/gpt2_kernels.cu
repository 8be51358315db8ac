// GPT-2 decode-side kernels: position add, KV-cache attention (prefill and single-query
// steps, with a per-position slot table so beam reorder never copies the cache), argmax /
// greedy bookkeeping on device (no per-token host sync, unlike
// core/scripts/benchmark_baseline.py:194-221), token embedding gather.
// Arithmetic follows transformers GPT2Attention / GPT2Model.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <algorithm>

namespace vc {

#define VC_LAUNCH(name, work, stream, ...)        \
  do {                                            \
    vc::KernelScope _ks(name, work, stream);      \
    __VA_ARGS__;                                  \
  } while (0)

// h[s, l, :] = embeds[s, l, :] + wpe[past_len + l, :]     (GPT2Model.forward: inputs_embeds + position_embeds)
__global__ void add_pos_kernel(const float* __restrict__ e, const float* __restrict__ wpe, float* __restrict__ h, int n_seq, int L,
                               int past_len, int dim4) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(n_seq) * L * dim4;
  if (i >= total) return;
  const int c = static_cast<int>(i % dim4);
  const int l = static_cast<int>((i / dim4) % L);
  const float4 a = reinterpret_cast<const float4*>(e)[i];
  const float4 p = __ldg(reinterpret_cast<const float4*>(wpe) + static_cast<long long>(past_len + l) * dim4 + c);
  reinterpret_cast<float4*>(h)[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}
int gpt_add_pos(const float* embeds, const float* wpe, float* h, int n_seq, int L, int past_len, int dim, cudaStream_t s) {
  VC_REQUIRE(dim % 4 == 0, "add_pos: dim %% 4");
  const long long total = static_cast<long long>(n_seq) * L * (dim / 4);
  if (total == 0) return 0;
  VC_LAUNCH("gpt_add_pos", total * 32.0, s,
            VC_CUDA_OK(launch_pdl(add_pos_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, embeds, wpe, h, n_seq, L, past_len,
                                  dim / 4)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// h = embeds + wpe[pos]; xn = LayerNorm(h) (bf16): the first two ops of a forward in one kernel, one warp per row
// (dim <= 1024, dim % 128 == 0), so a decode step starts with one dependent launch instead of two.
__global__ void __launch_bounds__(256) add_pos_ln_kernel(const float* __restrict__ e, const float* __restrict__ wpe, float* __restrict__ h,
                                                         __nv_bfloat16* __restrict__ xn, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int rows, int L, int past_len, int dim, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int nv = dim >> 7;                               // float4 per lane
  float4 g[8], bt[8], pe[8];
  const int l = row < rows ? row % L : 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {                                        // constants: before the wait
      g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
      bt[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
      pe[i] = __ldg(reinterpret_cast<const float4*>(wpe + static_cast<long long>(past_len + l) * dim) + lane + 32 * i);
    }
  pdl_wait();
  pdl_launch_dependents();
  if (row >= rows) return;
  float4 v[8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      const float4 a = reinterpret_cast<const float4*>(e + static_cast<long long>(row) * dim)[lane + 32 * i];
      v[i] = make_float4(a.x + pe[i].x, a.y + pe[i].y, a.z + pe[i].z, a.w + pe[i].w);
      reinterpret_cast<float4*>(h + static_cast<long long>(row) * dim)[lane + 32 * i] = v[i];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  const float mean = warp_sum(sum) / static_cast<float>(dim);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(dim) + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      uint2 w;
      w.x = pack_bf16((v[i].x - mean) * rstd * g[i].x + bt[i].x, (v[i].y - mean) * rstd * g[i].y + bt[i].y);
      w.y = pack_bf16((v[i].z - mean) * rstd * g[i].z + bt[i].z, (v[i].w - mean) * rstd * g[i].w + bt[i].w);
      reinterpret_cast<uint2*>(xn + static_cast<long long>(row) * dim)[lane + 32 * i] = w;
    }
}
int gpt_add_pos_ln(const float* embeds, const float* wpe, float* h, void* xn, const float* gamma, const float* beta, int n_seq, int L,
                   int past_len, int dim, float eps, cudaStream_t s) {
  VC_REQUIRE(dim % 128 == 0 && dim <= 1024, "add_pos_ln: dim=%d", dim);
  const int rows = n_seq * L;
  if (rows == 0) return 0;
  VC_LAUNCH("gpt_add_pos_ln", rows * dim * 10.0, s,
            VC_CUDA_OK(launch_pdl(add_pos_ln_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, embeds, wpe, h, static_cast<__nv_bfloat16*>(xn), gamma, beta,
                                  rows, L, past_len, dim, eps)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ KV-cache attention
// cache layout: kv[layer][k|v][seq][head][s_max][64] bf16.  slot[seq][pos] (optional) is the
// physical seq row that holds position pos of logical row seq (beam search indirection).
// One CTA per (seq, head): first appends the L new K/V rows, then each warp handles query
// positions l = warp, warp+4, ...; lanes split the keys for the scores (one 128-byte row per
// lane) and split head_dim for the weighted V sum (coalesced 128-byte rows).
constexpr int GHD = 64;
constexpr int ATT_MAX_S = 1024;   // GPT-2 n_positions

// 8 consecutive qkv values of one row as packed bf16: either from the bf16 GEMM output (prefill) or
// bias + the fixed-order sum of the skinny GEMM's split-K partials (decode step), rounded to bf16.
__device__ __forceinline__ uint4 fetch_qkv8(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ P, int ksplit,
                                            const float* __restrict__ bias, size_t plane, size_t row, int ld, int col) {
  if (P == nullptr) return *reinterpret_cast<const uint4*>(qkv + row * ld + col);
  float4 a = __ldg(reinterpret_cast<const float4*>(bias + col)), b = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
  for (int s0 = 0; s0 < ksplit; s0 += 6) {               // fixed summation order; 12 loads in flight
    float4 u[6], v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (s0 + i < ksplit) {
        const float* p = P + (s0 + i) * plane + row * ld + col;
        u[i] = *reinterpret_cast<const float4*>(p);
        v[i] = *reinterpret_cast<const float4*>(p + 4);
      }
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (s0 + i < ksplit) {
        a.x += u[i].x; a.y += u[i].y; a.z += u[i].z; a.w += u[i].w;
        b.x += v[i].x; b.y += v[i].y; b.z += v[i].z; b.w += v[i].w;
      }
  }
  return make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

__global__ void __launch_bounds__(128) gpt_attention_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ P, int ksplit,
                                                            const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                            __nv_bfloat16* __restrict__ kv, const int32_t* __restrict__ slot,
                                                            int layer, int n_seq, int rows_total, int heads, int s_max, int L, int past_len) {
  __shared__ float s_p[4][ATT_MAX_S];
  pdl_wait();
  pdl_launch_dependents();
  const int seq = blockIdx.x / heads, head = blockIdx.x - seq * heads;
  const int H = heads * GHD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long plane = static_cast<long long>(n_seq) * heads * s_max * GHD;        // one of K or V for one layer
  __nv_bfloat16* kbase = kv + (static_cast<long long>(layer) * 2 + 0) * plane;
  __nv_bfloat16* vbase = kv + (static_cast<long long>(layer) * 2 + 1) * plane;
  const long long own = (static_cast<long long>(seq) * heads + head) * s_max * GHD;
  const size_t plane_p = static_cast<size_t>(rows_total) * 3 * H;

  // append new K/V rows (16-byte chunks: 8 per row)
  for (int i = tid; i < L * 16; i += blockDim.x) {
    const int l = i >> 4, which = (i >> 3) & 1, chunk = i & 7;
    const uint4 v = fetch_qkv8(qkv, P, ksplit, bias, plane_p, static_cast<size_t>(seq) * L + l, 3 * H, (1 + which) * H + head * GHD + chunk * 8);
    __nv_bfloat16* dst = (which ? vbase : kbase) + own + static_cast<long long>(past_len + l) * GHD;
    reinterpret_cast<uint4*>(dst)[chunk] = v;
  }
  __syncthreads();

  const float scale = 0.125f;   // head_dim^-0.5
  for (int l = warp; l < L; l += 4) {
    const int n_keys = past_len + l + 1;   // causal
    // q in registers: every lane holds the full 64-dim query (bf16 pairs)
    float q[GHD];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = fetch_qkv8(qkv, P, ksplit, bias, plane_p, static_cast<size_t>(seq) * L + l, 3 * H, head * GHD + c * 8);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
      q[c * 8 + 0] = a.x; q[c * 8 + 1] = a.y; q[c * 8 + 2] = b.x; q[c * 8 + 3] = b.y;
      q[c * 8 + 4] = cc.x; q[c * 8 + 5] = cc.y; q[c * 8 + 6] = d.x; q[c * 8 + 7] = d.y;
    }
    float mx = -INFINITY;
    for (int j = lane; j < n_keys; j += 32) {
      const int phys = (slot != nullptr && j < past_len) ? slot[static_cast<long long>(seq) * s_max + j] : seq;
      const uint4* kp = reinterpret_cast<const uint4*>(kbase + (static_cast<long long>(phys) * heads + head) * s_max * GHD + static_cast<long long>(j) * GHD);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = kp[c];
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
        acc = fmaf(q[c * 8 + 0], a.x, acc); acc = fmaf(q[c * 8 + 1], a.y, acc);
        acc = fmaf(q[c * 8 + 2], b.x, acc); acc = fmaf(q[c * 8 + 3], b.y, acc);
        acc = fmaf(q[c * 8 + 4], cc.x, acc); acc = fmaf(q[c * 8 + 5], cc.y, acc);
        acc = fmaf(q[c * 8 + 6], d.x, acc); acc = fmaf(q[c * 8 + 7], d.y, acc);
      }
      acc *= scale;
      s_p[warp][j] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n_keys; j += 32) {
      const float p = __expf(s_p[warp][j] - mx);
      s_p[warp][j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < n_keys; ++j) {
      const int phys = (slot != nullptr && j < past_len) ? slot[static_cast<long long>(seq) * s_max + j] : seq;
      const uint32_t u = *(reinterpret_cast<const uint32_t*>(vbase + (static_cast<long long>(phys) * heads + head) * s_max * GHD + static_cast<long long>(j) * GHD) + lane);
      const float2 v = unpack_bf16(u);
      const float p = s_p[warp][j];
      o0 = fmaf(p, v.x, o0);
      o1 = fmaf(p, v.y, o1);
    }
    const float inv = 1.f / sum;
    *(reinterpret_cast<uint32_t*>(out + (static_cast<long long>(seq) * L + l) * H + head * GHD) + lane) = pack_bf16(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

// Short-context variant (past_len + L <= 256, i.e. every caption): all K/V rows of this (seq, head) are staged
// in shared memory with every thread issuing independent 16-byte loads, so a decode step pays about two
// memory round trips instead of one per key.  Rows are padded to 144 B (conflict-free 16-byte reads).
constexpr int ATT_SMEM_MAX_S = 256;
constexpr int ATT_ROW_B = 144;
__global__ void __maxnreg__(144) gpt_attention_smem_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ P, int ksplit,
                                                                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                                 __nv_bfloat16* __restrict__ kv, const int32_t* __restrict__ slot, int layer,
                                                                 int n_seq, int rows_total, int heads, int s_max, int L, int past_len) {
  extern __shared__ __align__(16) uint8_t s_kv[];            // K rows then V rows, ATT_ROW_B bytes each
  __shared__ float s_p[4][ATT_SMEM_MAX_S];
  const int seq = blockIdx.x / heads, head = blockIdx.x - seq * heads;
  const int H = heads * GHD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_total = past_len + L;
  const long long plane = static_cast<long long>(n_seq) * heads * s_max * GHD;
  __nv_bfloat16* kbase = kv + (static_cast<long long>(layer) * 2 + 0) * plane;
  __nv_bfloat16* vbase = kv + (static_cast<long long>(layer) * 2 + 1) * plane;
  const long long own = (static_cast<long long>(seq) * heads + head) * s_max * GHD;
  const size_t plane_p = static_cast<size_t>(rows_total) * 3 * H;
  uint8_t* sK = s_kv;
  uint8_t* sV = s_kv + n_total * ATT_ROW_B;
  float* sQ = reinterpret_cast<float*>(s_kv + 2 * n_total * ATT_ROW_B);   // [L][64] fp32 (bf16-rounded values)
  // cached rows (positions < past_len): asynchronous 16-byte copies, all in flight at once
  auto stage_cached = [&]() {
    for (int i = tid; i < past_len * 16; i += blockDim.x) {
      const int j = i >> 4, which = (i >> 3) & 1, chunk = i & 7;
      const int phys = slot != nullptr ? slot[static_cast<long long>(seq) * s_max + j] : seq;
      const __nv_bfloat16* src = (which ? vbase : kbase) + (static_cast<long long>(phys) * heads + head) * s_max * GHD + static_cast<long long>(j) * GHD;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32((which ? sV : sK) + j * ATT_ROW_B + chunk * 16)),
                   "l"(reinterpret_cast<const uint4*>(src) + chunk)
                   : "memory");
    }
  };
  // Dependents are released at entry (decode_chain.cu: dc_fullk_kernel), so before the wait only data that is a whole
  // forward call old may be read.  Without a slot table the cached rows were written by this layer's attention kernels of
  // EARLIER forward calls (dozens of kernels back, and no two full-K products of the chain can be resident together): they
  // are requested before the wait and arrive while the QKV product is still running.  With a slot table (beam search) the
  // table itself may be one kernel old.
  pdl_launch_dependents();
  if (slot == nullptr) stage_cached();
  pdl_wait();
  if (slot != nullptr) stage_cached();
  // the L new rows (appended to the cache) and the L query rows
  for (int i = tid; i < L * 24; i += blockDim.x) {
    if (i >= L * 16) {                                                   // query rows
      const int qi = i - L * 16, l = qi >> 3, chunk = qi & 7;
      const uint4 u = fetch_qkv8(qkv, P, ksplit, bias, plane_p, static_cast<size_t>(seq) * L + l, 3 * H, head * GHD + chunk * 8);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
      float4* dst = reinterpret_cast<float4*>(sQ + l * GHD + chunk * 8);
      dst[0] = make_float4(a.x, a.y, b.x, b.y);
      dst[1] = make_float4(cc.x, cc.y, d.x, d.y);
      continue;
    }
    const int l = i >> 4, which = (i >> 3) & 1, chunk = i & 7, j = past_len + l;
    const uint4 v = fetch_qkv8(qkv, P, ksplit, bias, plane_p, static_cast<size_t>(seq) * L + l, 3 * H, (1 + which) * H + head * GHD + chunk * 8);
    reinterpret_cast<uint4*>((which ? vbase : kbase) + own + static_cast<long long>(j) * GHD)[chunk] = v;   // append to the cache
    *reinterpret_cast<uint4*>((which ? sV : sK) + j * ATT_ROW_B + chunk * 16) = v;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const float scale = 0.125f;
  for (int l = warp; l < L; l += 4) {
    const int n_keys = past_len + l + 1;
    float q[GHD];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 f = reinterpret_cast<const float4*>(sQ + l * GHD)[c];   // broadcast read
      q[c * 4 + 0] = f.x; q[c * 4 + 1] = f.y; q[c * 4 + 2] = f.z; q[c * 4 + 3] = f.w;
    }
    float mx = -INFINITY;
    for (int j = lane; j < n_keys; j += 32) {
      const uint4* kp = reinterpret_cast<const uint4*>(sK + j * ATT_ROW_B);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = kp[c];
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
        acc = fmaf(q[c * 8 + 0], a.x, acc); acc = fmaf(q[c * 8 + 1], a.y, acc);
        acc = fmaf(q[c * 8 + 2], b.x, acc); acc = fmaf(q[c * 8 + 3], b.y, acc);
        acc = fmaf(q[c * 8 + 4], cc.x, acc); acc = fmaf(q[c * 8 + 5], cc.y, acc);
        acc = fmaf(q[c * 8 + 6], d.x, acc); acc = fmaf(q[c * 8 + 7], d.y, acc);
      }
      acc *= scale;
      s_p[warp][j] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n_keys; j += 32) {
      const float p = __expf(s_p[warp][j] - mx);
      s_p[warp][j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < n_keys; ++j) {
      const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(sV + j * ATT_ROW_B + lane * 4));
      const float p = s_p[warp][j];
      o0 = fmaf(p, v.x, o0);
      o1 = fmaf(p, v.y, o1);
    }
    const float inv = 1.f / sum;
    *(reinterpret_cast<uint32_t*>(out + (static_cast<long long>(seq) * L + l) * H + head * GHD) + lane) = pack_bf16(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

int gpt_attention(const void* qkv, const float* P, int ksplit, const float* bias, void* out, const VcKvCache* c, int layer, int n_seq,
                  int L, int past_len, cudaStream_t s) {
  VC_REQUIRE(c->head_dim == GHD, "gpt_attention: head_dim=%d (only 64 is built)", c->head_dim);
  VC_REQUIRE(past_len + L <= c->s_max && c->s_max <= ATT_MAX_S, "gpt_attention: %d+%d positions exceed cache s_max=%d", past_len, L, c->s_max);
  VC_REQUIRE(n_seq <= c->n_seq && layer < c->layers, "gpt_attention: cache too small");
  if (n_seq <= 0 || L <= 0) return 0;
  const double bytes = static_cast<double>(n_seq) * c->heads * GHD * 2.0 * (2.0 * L + 2.0 * (past_len + L));
  if (past_len + L <= ATT_SMEM_MAX_S) {
    const int smem = 2 * (past_len + L) * ATT_ROW_B + L * GHD * 4;
    {
      // per device (a second GPU in the same process needs its own opt-in); the largest shared-memory carve-out so that these
      // CTAs can become resident beside the chain's large-shared-memory kernels without the SM re-splitting L1 / shared
      static bool attr[64] = {false};
      int dev = 0;
      VC_CUDA_OK(cudaGetDevice(&dev));
      if (dev < 0 || dev >= 64 || !attr[dev]) {
        VC_CUDA_OK(cudaFuncSetAttribute(gpt_attention_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_MAX_S * (2 * ATT_ROW_B + GHD * 4)));
        VC_CUDA_OK(cudaFuncSetAttribute(gpt_attention_smem_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        if (dev >= 0 && dev < 64) attr[dev] = true;
      }
    }
    VC_LAUNCH("gpt_attention", bytes, s,
              // one new position per sequence: only one warp computes, so the CTA is that one warp (32 CTAs per SM instead of
              // 16, no idle threads): n_seq * heads CTAs of a 256-sequence step stay in a single wave (19.4 -> 16.6 ms per 20 tokens)
              VC_CUDA_OK(launch_pdl(gpt_attention_smem_kernel, dim3(n_seq * c->heads), dim3(L == 1 ? 32 : 128), static_cast<size_t>(smem), s,
                                    static_cast<const __nv_bfloat16*>(qkv), P, ksplit, bias, static_cast<__nv_bfloat16*>(out),
                                    static_cast<__nv_bfloat16*>(c->kv), static_cast<const int32_t*>(c->slot), layer, c->n_seq, n_seq * L, c->heads,
                                    c->s_max, L, past_len)));
    VC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  VC_LAUNCH("gpt_attention", bytes, s,
            VC_CUDA_OK(launch_pdl(gpt_attention_kernel, dim3(n_seq * c->heads), dim3(128), 0, s, static_cast<const __nv_bfloat16*>(qkv), P, ksplit, bias,
                                  static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(c->kv), static_cast<const int32_t*>(c->slot), layer,
                                  c->n_seq, n_seq * L, c->heads, c->s_max, L, past_len)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ argmax (ties -> lowest index, torch.argmax)
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__device__ int block_argmax(const float* __restrict__ row, int vocab) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    // 16-byte loads, 8 in flight per thread: a 50k-entry row costs two memory round trips instead of a dozen
    const int n4 = vocab >> 2;
    for (int j0 = threadIdx.x; j0 < n4; j0 += blockDim.x * 8) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = j0 + i * blockDim.x;
        if (j < n4) v[i] = reinterpret_cast<const float4*>(row)[j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = j0 + i * blockDim.x;
        if (j < n4) {                                   // increasing index order, strict '>' keeps the lowest index of a tie
          if (v[i].x > best) { best = v[i].x; bi = 4 * j; }
          if (v[i].y > best) { best = v[i].y; bi = 4 * j + 1; }
          if (v[i].z > best) { best = v[i].z; bi = 4 * j + 2; }
          if (v[i].w > best) { best = v[i].w; bi = 4 * j + 3; }
        }
      }
    }
    for (int j = (n4 << 2) + threadIdx.x; j < vocab; j += blockDim.x) {
      const float v = row[j];
      if (v > best || (v == best && j < bi)) { best = v; bi = j; }
    }
  } else {
    for (int j = threadIdx.x; j < vocab; j += blockDim.x) {
      const float v = row[j];
      if (v > best || (v == best && j < bi)) { best = v; bi = j; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    argmax_combine(best, bi, ov, oi);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { s_v[warp] = best; s_i[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < nw ? s_v[lane] : -INFINITY;
    bi = lane < nw ? s_i[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      argmax_combine(best, bi, ov, oi);
    }
    if (lane == 0) s_i[0] = (bi == 0x7fffffff) ? 0 : bi;
  }
  __syncthreads();
  return s_i[0];
}

__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ logits, long long ld, int vocab, int32_t* __restrict__ out) {
  const int r = blockIdx.x;
  const int idx = block_argmax(logits + r * ld, vocab);
  if (threadIdx.x == 0) out[r] = idx;
}
int argmax_f32(const float* logits, long long ld, int rows, int vocab, int32_t* out, cudaStream_t s) {
  if (rows <= 0) return 0;
  VC_REQUIRE(vocab > 0, "argmax: vocab=%d", vocab);
  VC_LAUNCH("argmax", static_cast<double>(rows) * vocab * 4.0, s, (argmax_kernel<<<rows, 1024, 0, s>>>(logits, ld, vocab, out)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ embedding gather (wte bf16 -> fp32)
__global__ void embed_kernel(const __nv_bfloat16* __restrict__ wte, const int32_t* __restrict__ ids, int dim, float* __restrict__ out) {
  const int r = blockIdx.x;
  const long long id = ids[r];
  for (int c = threadIdx.x; c < dim; c += blockDim.x) out[static_cast<long long>(r) * dim + c] = __bfloat162float(wte[id * dim + c]);
}
int embed_tokens(const void* wte, const int32_t* ids, int n, int dim, float* out, cudaStream_t s) {
  if (n <= 0) return 0;
  VC_LAUNCH("embed_tokens", n * dim * 6.0, s, (embed_kernel<<<n, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(wte), ids, dim, out)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// inputs_embeds = cat([prefix, wte(prompt)], dim=1)   (text_decoder.py:60-74, benchmark_baseline.py:176-180)
__global__ void prefill_embeds_kernel(const float* __restrict__ prefix, const __nv_bfloat16* __restrict__ wte,
                                      const int32_t* __restrict__ prompt, int P, int Lp, int dim, float* __restrict__ out) {
  const int s = blockIdx.x / (P + Lp), l = blockIdx.x - s * (P + Lp);
  float* o = out + (static_cast<long long>(s) * (P + Lp) + l) * dim;
  if (l < P) {
    const float* src = prefix + (static_cast<long long>(s) * P + l) * dim;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) o[c] = src[c];
  } else {
    const long long id = prompt[l - P];
    for (int c = threadIdx.x; c < dim; c += blockDim.x) o[c] = __bfloat162float(wte[id * dim + c]);
  }
}
int build_prefill_embeds(const float* prefix, const void* wte, const int32_t* prompt_ids, int n_seq, int P, int Lp, int dim, float* out,
                         cudaStream_t s) {
  if (n_seq <= 0) return 0;
  VC_LAUNCH("prefill_embeds", 0.0, s,
            (prefill_embeds_kernel<<<n_seq * (P + Lp), 256, 0, s>>>(prefix, static_cast<const __nv_bfloat16*>(wte), prompt_ids, P, Lp, dim, out)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ greedy bookkeeping
__global__ void greedy_init_kernel(int32_t* ids_out, int32_t* len_out, int32_t* finished, int n_seq, int max_new, int eos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_seq * max_new) ids_out[i] = eos;
  if (i < n_seq) { len_out[i] = 0; finished[i] = 0; }
}
int greedy_init(int32_t* ids_out, int32_t* len_out, int32_t* finished, int n_seq, int max_new, int eos, cudaStream_t s) {
  const int total = std::max(n_seq * max_new, n_seq);
  if (total <= 0) return 0;
  VC_LAUNCH("greedy_init", 0.0, s, (greedy_init_kernel<<<(total + 255) / 256, 256, 0, s>>>(ids_out, len_out, finished, n_seq, max_new, eos)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// benchmark_baseline.py:210-227: next = argmax(logits[:, -1]); finished rows -> eos; unfinished rows
// append the token (including their first eos) and become finished on eos; next input = wte(next).
__global__ void __launch_bounds__(1024) greedy_select_kernel(const float* __restrict__ logits, long long ld, int vocab, int step, int max_new,
                                                             int eos, int32_t* __restrict__ finished, int32_t* __restrict__ ids_out,
                                                             int32_t* __restrict__ len_out, const int32_t* __restrict__ forced,
                                                             const __nv_bfloat16* __restrict__ wte, int dim, float* __restrict__ next_embeds,
                                                             int32_t* __restrict__ next_ids) {
  pdl_wait();
  pdl_launch_dependents();
  const int r = blockIdx.x;
  int tok = block_argmax(logits + r * ld, vocab);
  const bool was_finished = finished[r] != 0;
  if (was_finished) tok = eos;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!was_finished) {
      ids_out[static_cast<long long>(r) * max_new + step] = tok;
      len_out[r] += 1;
      if (tok == eos) finished[r] = 1;
    }
    if (next_ids != nullptr) next_ids[r] = tok;
  }
  const long long feed = (forced != nullptr) ? forced[static_cast<long long>(r) * max_new + step] : tok;
  if (next_embeds != nullptr)
    for (int c = threadIdx.x; c < dim; c += blockDim.x) next_embeds[static_cast<long long>(r) * dim + c] = __bfloat162float(wte[feed * dim + c]);
}
int greedy_select(const float* logits, long long ld, int vocab, int n_seq, int step, int max_new, int eos, int32_t* finished,
                  int32_t* ids_out, int32_t* len_out, const int32_t* forced, const void* wte, int dim, float* next_embeds,
                  int32_t* next_ids, cudaStream_t s) {
  if (n_seq <= 0) return 0;
  VC_LAUNCH("greedy_select", static_cast<double>(n_seq) * vocab * 4.0, s,
            VC_CUDA_OK(launch_pdl(greedy_select_kernel, dim3(n_seq), dim3(1024), 0, s, logits, ld, vocab, step, max_new, eos, finished, ids_out, len_out,
                                  forced, static_cast<const __nv_bfloat16*>(wte), dim, next_embeds, next_ids)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vc
