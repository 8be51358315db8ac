// Persistent cooperative GPT-2 decode kernel: ALL decode steps of a greedy caption loop (or one
// forward step for the beam / HF path) in ONE launch.
//
// Why: one decode step is ~88 dependent tiny operations (12 layers x {QKV, attention, proj, LN, fc1,
// fc2, LN} + lm_head + token selection).  As separate launches inside a CUDA graph each costs 4-6 us
// of launch / ramp / drain, so the step took ~510 us while streaming its 284 MB takes 43 us at the
// measured HBM rate.  Here the 148 CTAs stay resident, phases are separated by a device-wide barrier
// (one red.release + one ld.acquire spin, ~0.5 us), and the weight stream is decoupled from the
// dependency chain: a producer warp TMA-prefetches weight slabs for upcoming phases (they do not depend
// on activations) while the CTA is still waiting on the barrier of the current one.
//
// Warp roles per CTA (256 threads): warp 0 = weight TMA producer (runs ahead across phases and steps),
// warp 1 = activation TMA producer (gated by the grid barrier), warp 2 = tcgen05.mma issuer + TMEM
// owner, warps 4-7 = epilogue (tcgen05.ld) and the row-parallel phases (LayerNorm, attention, token
// selection).
//
// GEMM phases: D[seq, feature] = acts[seq, K-slice] * W[feature tile, K-slice]^T with the ACTIVATIONS
// as the A operand (M = 128 TMEM lanes, rows >= n_seq zero-filled by TMA) and the weight tile as B, so
// the feature-tile width NT is any multiple of 16 and (NT, K-split) can be chosen per GEMM to give
// ~148 work items: QKV 64x4, proj 32x6, fc1 32x1, fc2 64x12, lm_head 128x1 for GPT-2 small.  K-split
// partials are fp32 in L2 and are summed in a fixed order by the consumer phase (deterministic).
//
// Arithmetic follows transformers GPT2Model / GPT2Attention and the greedy bookkeeping
// of core/scripts/benchmark_baseline.py:210-227.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <cudaTypedefs.h>
#include <algorithm>
#include <cstdlib>

namespace vc {

int make_tmap_bf16_kmajor(CUtensorMap* tm, const void* ptr, int rows, int K, int box_rows);

namespace {

constexpr int DK_THREADS = 256;
constexpr int DK_MAX_LAYERS = 24;
constexpr int DK_W_STAGES = 5;
constexpr int DK_STAGE_BYTES = 16384;            // W stage: up to 128 rows x 128 B (or several 64-k chunks of a narrower tile)
constexpr int DK_A_RING_BYTES = 6 * 16384;       // activation ring: 12 x 8 KB (n_seq <= 64: 64-row slabs) or 6 x 16 KB (128-row slabs)
constexpr int DK_A_MAX_STAGES = 12;
constexpr int DK_MAX_S = 1024;                   // GPT-2 n_positions
constexpr int DK_GROUPS = 8;                     // attention work groups (half warps) per CTA
constexpr int DK_MISC_BYTES = 4096;
constexpr int DK_SMEM = DK_W_STAGES * DK_STAGE_BYTES + DK_A_RING_BYTES + DK_GROUPS * DK_MAX_S * 4 + DK_MISC_BYTES + 1024;

enum { G_QKV = 0, G_APROJ = 1, G_FC1 = 2, G_FC2 = 3, G_LMHEAD = 4 };

struct DkLayer {
  const float *ln1_g, *ln1_b, *attn_b, *aproj_b, *ln2_g, *ln2_b, *fc_b, *mproj_b;
};

struct DkParams {
  CUtensorMap wmap[DK_MAX_LAYERS * 4 + 1];   // [layer*4 + g]; last used = wte (lm_head)
  CUtensorMap amap[3];                        // xn, att, hid  (rows = n_seq, box 64 x 128; rows >= n_seq are zero-filled)
  DkLayer layer[DK_MAX_LAYERS];
  int H, heads, layers, vocab, vocab_pad, n_seq;
  int nt[5], ks[5];
  const float* wpe; const float* lnf_g; const float* lnf_b; const __nv_bfloat16* wte;
  float* h; __nv_bfloat16* xn; __nv_bfloat16* att; __nv_bfloat16* hid; float* part; float* cand_v; int* cand_i;
  __nv_bfloat16* kv; const int32_t* slot; int cache_n_seq, s_max;
  unsigned int* bar;
  const float* emb;   // inputs_embeds of the first step of this launch, fp32 [n_seq, H], without position embedding
  int past0;          // positions already in the cache before that step
  int n_steps;
  int greedy, step0, max_new, eos;
  int32_t* finished; int32_t* ids_out; int32_t* len_out; const int32_t* forced; int32_t* next_ids;
  float* logits; long long logits_step_stride;   // optional fp32 [step][n_seq][vocab_pad]
  unsigned long long* prof;   // optional: CTA 0 records %globaltimer when it observes each phase complete
  int max_phases;     // debugging aid: stop every role after this many phases (INT_MAX = run to the end)
};

struct Gd {
  int wmap, amap, N, K, NT, ks, chunks, n_items, cps, g, layer;
};

__device__ __forceinline__ Gd get_gd(const DkParams& p, int l, int g) {
  Gd d;
  d.g = g; d.layer = l;
  const int H = p.H;
  switch (g) {
    case G_QKV:   d.N = 3 * H; d.K = H;     d.amap = 0; d.wmap = l * 4 + 0; break;
    case G_APROJ: d.N = H;     d.K = H;     d.amap = 1; d.wmap = l * 4 + 1; break;
    case G_FC1:   d.N = 4 * H; d.K = H;     d.amap = 0; d.wmap = l * 4 + 2; break;
    case G_FC2:   d.N = H;     d.K = 4 * H; d.amap = 2; d.wmap = l * 4 + 3; break;
    default:      d.N = p.vocab_pad; d.K = H; d.amap = 0; d.wmap = p.layers * 4; break;
  }
  d.NT = p.nt[g]; d.ks = p.ks[g];
  d.chunks = d.K / 64 / d.ks;
  d.n_items = (d.N / d.NT) * d.ks;
  d.cps = DK_STAGE_BYTES / (d.NT * 128);   // 64-k chunks of this tile per W stage
  return d;
}

// index of a GEMM phase in the launch-wide phase sequence: [embed] then per step {7 per layer, lm_head, (select)}
__device__ __forceinline__ int gemm_phase_index(const DkParams& p, int s, int l, int g) {
  const int L = p.layers;
  const int per_step = 7 * L + 1 + (p.greedy ? 1 : 0);
  return 1 + s * per_step + (l < L ? 7 * l + (g == G_QKV ? 0 : g == G_APROJ ? 2 : g == G_FC1 ? 4 : 5) : 7 * L);
}

// ------------------------------------------------------------------ device-wide barrier (monotonic counter)
__device__ __forceinline__ void grid_arrive(unsigned int* bar) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until `phases_done` phases have been completed by every CTA: relaxed polls (with a short back-off, there
// are ~300 pollers on this one L2 line), then ONE gpu-scope acquire fence
__device__ __forceinline__ void grid_wait(const unsigned int* bar, unsigned int phases_done) {
  const unsigned int target = phases_done * gridDim.x;
#if VC_WATCHDOG
  unsigned int spins = 0;
  while (ld_relaxed_u32(bar) < target) {
    __nanosleep(20);
    if (++spins > (1u << 25)) __trap();
  }
#else
  while (ld_relaxed_u32(bar) < target) __nanosleep(20);
#endif
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void bar_sync_epi() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float2 ldcg_f2(const float* p) { return __ldcg(reinterpret_cast<const float2*>(p)); }

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ------------------------------------------------------------------ row phases (128 epilogue threads)
// sum over the 4 epilogue warps
__device__ __forceinline__ float epi_block_sum(float v, float* s_red, int te) {
  v = warp_sum(v);
  if ((te & 31) == 0) s_red[te >> 5] = v;
  bar_sync_epi();
  const float t = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
  bar_sync_epi();
  return t;
}

// The row phases are __noinline__ with by-value arguments: each gets its own register allocation (inlined into the
// 255-register main kernel they spilled), and no access goes through a generic pointer to the parameter space.
struct RowCtx {
  float* h; __nv_bfloat16* xn; const float* part; int n_seq, H;
};

// v[] = the new residual row (thread te owns float2 columns 2*te + 256*i): write h, LayerNorm, write xn (bf16)
__device__ __forceinline__ void ln_tail(const RowCtx p, int r, float2 (&v)[4], const float* gamma, const float* beta, float* s_red, int te) {
  const int H = p.H, nv = H >> 8;
  float sum = 0.f;
  float2 gm[4], bt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nv) {
      const int c = 2 * te + 256 * i;
      gm[i] = __ldg(reinterpret_cast<const float2*>(gamma + c));
      bt[i] = __ldg(reinterpret_cast<const float2*>(beta + c));
      *reinterpret_cast<float2*>(p.h + static_cast<size_t>(r) * H + c) = v[i];
      sum += v[i].x + v[i].y;
    }
  const float mean = epi_block_sum(sum, s_red, te) / static_cast<float>(H);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean;
      sq += a * a + b * b;
    }
  const float rstd = rsqrtf(epi_block_sum(sq, s_red, te) / static_cast<float>(H) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nv) {
      const int c = 2 * te + 256 * i;
      *reinterpret_cast<uint32_t*>(p.xn + static_cast<size_t>(r) * H + c) =
          pack_bf16((v[i].x - mean) * rstd * gm[i].x + bt[i].x, (v[i].y - mean) * rstd * gm[i].y + bt[i].y);
    }
}

// h[r] += bias + sum_s P[s][r]; xn[r] = LN(h[r])        (one row per CTA)
constexpr int DK_MAX_KS = 12;
__device__ __forceinline__ void row_resid_ln(const RowCtx p, int ks, const float* bias, const float* gamma, const float* beta,
                                          float* s_red, int te) {
  const int r = blockIdx.x;
  if (r >= p.n_seq) return;
  const int H = p.H, nv = H >> 8;
  const size_t plane = static_cast<size_t>(p.n_seq) * H;
  // per column pair: h, bias and all ks partials in flight at once (one L2 round trip), then a fixed-order sum
  float2 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nv) {
      const int c = 2 * te + 256 * i;
      const size_t o = static_cast<size_t>(r) * H + c;
      float2 q[DK_MAX_KS];
      float2 a = ldcg_f2(p.h + o);
#pragma unroll
      for (int s = 0; s < DK_MAX_KS; ++s)
        if (s < ks) q[s] = ldcg_f2(p.part + s * plane + o);
      const float2 b = __ldg(reinterpret_cast<const float2*>(bias + c));
      a.x += b.x; a.y += b.y;
#pragma unroll
      for (int s = 0; s < DK_MAX_KS; ++s)
        if (s < ks) { a.x += q[s].x; a.y += q[s].y; }
      v[i] = a;
    }
  ln_tail(p, r, v, gamma, beta, s_red, te);
}

// h[r] = emb[r] + wpe[pos]; xn[r] = LN1 of layer 0
__device__ __forceinline__ void row_embed(const RowCtx p, const float* wpe, const float* g0, const float* b0, const float* emb_row,
                                          const __nv_bfloat16* wte_row, int pos, float* s_red, int te) {
  const int r = blockIdx.x;
  const int H = p.H, nv = H >> 8;
  float2 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nv) {
      const int c = 2 * te + 256 * i;
      float2 a;
      if (emb_row != nullptr) a = ldcg_f2(emb_row + c);
      else a = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(wte_row + c)));
      const float2 w = __ldg(reinterpret_cast<const float2*>(wpe + static_cast<size_t>(pos) * H + c));
      v[i] = make_float2(a.x + w.x, a.y + w.y);
    }
  ln_tail(p, r, v, g0, b0, s_red, te);
}

// bf16-rounded (bias + sum_s P[s][row][col..col+3]) of the QKV product, as the prefill GEMM would have stored it;
// q, k and v quads of one (row, head, lane) together so that all 3*ks partial loads are in flight at once.
__device__ __forceinline__ float4 round4_bf16(float4 a) {
  const float2 lo = unpack_bf16(pack_bf16(a.x, a.y)), hi = unpack_bf16(pack_bf16(a.z, a.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
struct AttCtx {
  const float* part; const float* bias; __nv_bfloat16* kv; const int32_t* slot; __nv_bfloat16* att;
  int n_seq, H, heads, cache_n_seq, s_max, ks;
};
__device__ __forceinline__ void qkv_triple(const AttCtx& p, int ks, const float* bias, int row, int col, float4& q4, float4& k4, float4& v4) {
  const int ld = 3 * p.H;
  const size_t plane = static_cast<size_t>(p.n_seq) * ld;
  const float* base = p.part + static_cast<size_t>(row) * ld + col;
  float4 a = __ldg(reinterpret_cast<const float4*>(bias + col));
  float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + p.H));
  float4 c = __ldg(reinterpret_cast<const float4*>(bias + col + 2 * p.H));
  for (int s0 = 0; s0 < ks; s0 += 4) {       // 12 independent 16-byte loads in flight per batch, fixed summation order
    float4 pq[4], pk[4], pv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (s0 + i < ks) {
        pq[i] = __ldcg(reinterpret_cast<const float4*>(base + (s0 + i) * plane));
        pk[i] = __ldcg(reinterpret_cast<const float4*>(base + (s0 + i) * plane + p.H));
        pv[i] = __ldcg(reinterpret_cast<const float4*>(base + (s0 + i) * plane + 2 * p.H));
      }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (s0 + i < ks) {
        a.x += pq[i].x; a.y += pq[i].y; a.z += pq[i].z; a.w += pq[i].w;
        b.x += pk[i].x; b.y += pk[i].y; b.z += pk[i].z; b.w += pk[i].w;
        c.x += pv[i].x; c.y += pv[i].y; c.z += pv[i].z; c.w += pv[i].w;
      }
  }
  q4 = round4_bf16(a); k4 = round4_bf16(b); v4 = round4_bf16(c);
}

__device__ __forceinline__ float half_sum(float v) {   // over the 16 lanes of a half warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// single-query attention over the KV cache, one HALF warp per (sequence, head) so that 148 x 8 groups cover the
// 768 items of 64 sequences x 12 heads in one round; appends the new K/V row.  Lane hl owns dims 4hl..4hl+3.
__device__ __forceinline__ void row_attention(const AttCtx p, int layer, int past, float* s_p, float* s_q, int wq, int lane) {
  const int heads = p.heads, H = p.H;
  const int n_items = p.n_seq * heads;
  const size_t plane = static_cast<size_t>(p.cache_n_seq) * heads * p.s_max * 64;
  __nv_bfloat16* kbase = p.kv + (static_cast<size_t>(layer) * 2 + 0) * plane;
  __nv_bfloat16* vbase = p.kv + (static_cast<size_t>(layer) * 2 + 1) * plane;
  const float* bias = p.bias;
  const int ks = p.ks;
  const int grp = wq * 2 + (lane >> 4), hl = lane & 15;
  float* sp = s_p + grp * DK_MAX_S;
  float* sq = s_q + grp * 64;
  const int first = blockIdx.x * DK_GROUPS + grp, stride = gridDim.x * DK_GROUPS;
  // both halves of a warp run the same number of rounds (the shuffles below are warp-wide instructions)
  const int rounds = (n_items - (blockIdx.x * DK_GROUPS + wq * 2) + stride - 1) / stride;
  for (int rd = 0; rd < rounds; ++rd) {
    const int item_raw = first + rd * stride;
    const bool valid = item_raw < n_items;
    const int item = valid ? item_raw : 0;
    const int seq = item / heads, head = item - seq * heads;
    const size_t seq_base = static_cast<size_t>(seq) * p.s_max;
    const size_t own = (static_cast<size_t>(seq) * heads + head) * p.s_max * 64 + static_cast<size_t>(past) * 64;
    float4 q4, k4, v4;
    qkv_triple(p, ks, bias, seq, head * 64 + 4 * hl, q4, k4, v4);
    if (valid) {
      *reinterpret_cast<uint2*>(kbase + own + 4 * hl) = make_uint2(pack_bf16(k4.x, k4.y), pack_bf16(k4.z, k4.w));
      *reinterpret_cast<uint2*>(vbase + own + 4 * hl) = make_uint2(pack_bf16(v4.x, v4.y), pack_bf16(v4.z, v4.w));
    }
    *reinterpret_cast<float4*>(sq + 4 * hl) = q4;
    __syncwarp();
    const float s_new = half_sum((q4.x * k4.x + q4.y * k4.y) + (q4.z * k4.z + q4.w * k4.w)) * 0.125f;
    float mx = s_new;
    for (int j = hl; j < past; j += 16) {
      const int phys = p.slot != nullptr ? p.slot[seq_base + j] : seq;
      const uint4* kp = reinterpret_cast<const uint4*>(kbase + (static_cast<size_t>(phys) * heads + head) * p.s_max * 64 + static_cast<size_t>(j) * 64);
      uint4 u[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) u[c] = __ldcg(kp + c);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 qa = *reinterpret_cast<const float4*>(sq + c * 8), qb = *reinterpret_cast<const float4*>(sq + c * 8 + 4);
        const float2 a = unpack_bf16(u[c].x), b = unpack_bf16(u[c].y), cc = unpack_bf16(u[c].z), d = unpack_bf16(u[c].w);
        acc = fmaf(qa.x, a.x, acc); acc = fmaf(qa.y, a.y, acc); acc = fmaf(qa.z, b.x, acc); acc = fmaf(qa.w, b.y, acc);
        acc = fmaf(qb.x, cc.x, acc); acc = fmaf(qb.y, cc.y, acc); acc = fmaf(qb.z, d.x, acc); acc = fmaf(qb.w, d.y, acc);
      }
      acc *= 0.125f;
      sp[j] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = half_max(mx);
    float sum = 0.f;
    for (int j = hl; j < past; j += 16) {
      const float e = __expf(sp[j] - mx);
      sp[j] = e;
      sum += e;
    }
    const float p_new = __expf(s_new - mx);
    sum = half_sum(sum) + p_new;
    __syncwarp();
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < past; j0 += 16) {        // 16 independent V-row loads in flight per lane
      uint2 u[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j = j0 + i;
        if (j < past) {
          const int phys = p.slot != nullptr ? p.slot[seq_base + j] : seq;
          u[i] = __ldcg(reinterpret_cast<const uint2*>(vbase + (static_cast<size_t>(phys) * heads + head) * p.s_max * 64 + static_cast<size_t>(j) * 64) + hl);
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j = j0 + i;
        if (j < past) {
          const float2 va = unpack_bf16(u[i].x), vb = unpack_bf16(u[i].y);
          const float pj = sp[j];
          o.x = fmaf(pj, va.x, o.x); o.y = fmaf(pj, va.y, o.y); o.z = fmaf(pj, vb.x, o.z); o.w = fmaf(pj, vb.y, o.w);
        }
      }
    }
    o.x = fmaf(p_new, v4.x, o.x); o.y = fmaf(p_new, v4.y, o.y); o.z = fmaf(p_new, v4.z, o.z); o.w = fmaf(p_new, v4.w, o.w);
    const float inv = 1.f / sum;
    if (valid)
      *reinterpret_cast<uint2*>(p.att + static_cast<size_t>(seq) * H + head * 64 + 4 * hl) =
          make_uint2(pack_bf16(o.x * inv, o.y * inv), pack_bf16(o.z * inv, o.w * inv));
    __syncwarp();
  }
}

__device__ __forceinline__ void argmax_combine2(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// final argmax over the per-CTA candidates of row r, greedy bookkeeping (benchmark_baseline.py:210-227), then the
// next step's input row: h = wte[feed] + wpe[pos_next], xn = LN1_0(h).
__device__ __forceinline__ void row_select(const DkParams& p, int s, int past, float* s_red, int* s_tok, int te) {
  const int r = blockIdx.x;
  if (r >= p.n_seq) return;
  const int G = gridDim.x;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = te; c < G; c += 128) {
    const float v = __ldcg(p.cand_v + static_cast<size_t>(c) * 128 + r);
    const int i = __ldcg(p.cand_i + static_cast<size_t>(c) * 128 + r);
    argmax_combine2(bv, bi, v, i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    argmax_combine2(bv, bi, ov, oi);
  }
  float* s_v = s_red;                       // [4] values, then ints
  int* s_i = reinterpret_cast<int*>(s_red + 4);
  if ((te & 31) == 0) { s_v[te >> 5] = bv; s_i[te >> 5] = bi; }
  bar_sync_epi();
  if (te == 0) {
    for (int w = 1; w < 4; ++w) argmax_combine2(bv, bi, s_v[w], s_i[w]);
    int tok = (bi == 0x7fffffff) ? 0 : bi;
    const int step = p.step0 + s;
    const bool was_finished = p.finished[r] != 0;
    if (was_finished) tok = p.eos;
    if (!was_finished) {
      p.ids_out[static_cast<size_t>(r) * p.max_new + step] = tok;
      p.len_out[r] += 1;
      if (tok == p.eos) p.finished[r] = 1;
    }
    if (p.next_ids != nullptr) p.next_ids[r] = tok;
    *s_tok = (p.forced != nullptr) ? p.forced[static_cast<size_t>(r) * p.max_new + step] : tok;
  }
  bar_sync_epi();
  const int feed = *s_tok;
  bar_sync_epi();
  if (s + 1 < p.n_steps)
    row_embed(RowCtx{p.h, p.xn, p.part, p.n_seq, p.H}, p.wpe, p.layer[0].ln1_g, p.layer[0].ln1_b, nullptr, p.wte + static_cast<size_t>(feed) * p.H,
              past + 1, s_red, te);
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(DK_THREADS, 1) gpt2_decode_kernel(const __grid_constant__ DkParams p) {
  extern __shared__ uint8_t dk_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dk_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* w_ring = smem;
  uint8_t* a_ring = smem + DK_W_STAGES * DK_STAGE_BYTES;
  float* s_p = reinterpret_cast<float*>(a_ring + DK_A_RING_BYTES);                   // [DK_GROUPS][DK_MAX_S]; also absorbs the
                                                                                     // (ignored) rows 64..127 the MMA reads past the last 8 KB slab
  uint8_t* misc = reinterpret_cast<uint8_t*>(s_p + DK_GROUPS * DK_MAX_S);
  uint64_t* fullW = reinterpret_cast<uint64_t*>(misc);
  uint64_t* emptyW = fullW + DK_W_STAGES;
  uint64_t* fullA = emptyW + DK_W_STAGES;
  uint64_t* emptyA = fullA + DK_A_MAX_STAGES;
  uint64_t* accfull = emptyA + DK_A_MAX_STAGES;     // [2]
  uint64_t* accempty = accfull + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
  int* s_tok = reinterpret_cast<int*>(tmem_slot + 1);
  float* s_red = reinterpret_cast<float*>(misc + 512);      // [8]
  float* s_q = reinterpret_cast<float*>(misc + 2048);       // [DK_GROUPS][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int L = p.layers;
  // activation slabs: [64 k x 64 rows] = 8 KB when n_seq <= 64 (the MMA's M=128 read then runs 8 KB into the next
  // slab: those accumulator rows are never used), else [64 k x 128 rows] = 16 KB
  const int a_stage_bytes = p.n_seq <= 64 ? 8192 : 16384;
  const int a_stages = DK_A_RING_BYTES / a_stage_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < DK_W_STAGES; ++s) { mbar_init(&fullW[s], 1); mbar_init(&emptyW[s], 1); }
    for (int s = 0; s < DK_A_MAX_STAGES; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&accfull[a], 1); mbar_init(&accempty[a], 4); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================= weight producer: runs ahead of everything, bounded by the ring
    int stage = 0; uint32_t phase = 0;
    for (int s = 0; s < p.n_steps; ++s) {
#pragma unroll 1
      for (int l = 0; l <= L; ++l) {
#pragma unroll 1
        for (int g = (l < L ? 0 : 4); g < (l < L ? 4 : 5); ++g) {
          const Gd d = get_gd(p, l, g);
          if (gemm_phase_index(p, s, l, g) >= p.max_phases) goto roles_done;
          for (int item = cta; item < d.n_items; item += G) {
            const int nt = item / d.ks, ksl = item - nt * d.ks;
            const int n0 = nt * d.NT, kc0 = ksl * d.chunks;
            for (int c0 = 0; c0 < d.chunks; c0 += d.cps) {
              const int nch = min(d.cps, d.chunks - c0);
              mbar_wait(&emptyW[stage], phase ^ 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(&fullW[stage], nch * d.NT * 128);
                for (int c = 0; c < nch; ++c)
                  tma_load_2d(&p.wmap[d.wmap], &fullW[stage], w_ring + stage * DK_STAGE_BYTES + c * d.NT * 128, (kc0 + c0 + c) * 64, n0);
              }
              __syncwarp();
              if (++stage == DK_W_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================= activation producer: one grid-barrier wait per GEMM phase, then all
    // chunks of the phase are issued back to back (TMA keeps every free ring stage in flight at once).
    int stage = 0; uint32_t phase = 0;
    for (int s = 0; s < p.n_steps; ++s) {
#pragma unroll 1
      for (int l = 0; l <= L; ++l) {
#pragma unroll 1
        for (int g = (l < L ? 0 : 4); g < (l < L ? 4 : 5); ++g) {
          const Gd d = get_gd(p, l, g);
          const int ph = gemm_phase_index(p, s, l, g);
          if (ph >= p.max_phases) goto roles_done;
          if (cta >= d.n_items) continue;
          if (lane == 0) grid_wait(p.bar, static_cast<unsigned int>(ph));   // one poller per warp; lane 0 is also the lane elect.sync picks
          __syncwarp();
          // the activations were written with generic-proxy stores by other CTAs earlier in this kernel: the
          // gpu-scope acquire fence inside grid_wait, then a cross-proxy fence before the async-proxy (TMA) reads
          fence_proxy_async();
          // "resident" activations: when the whole K extent fits the ring exactly (lm_head of GPT-2 small: 12 slabs),
          // they are loaded once and reused by every item of this CTA in the phase
          const bool resident = d.ks == 1 && d.chunks == a_stages;
          for (int item = cta; item < d.n_items; item += G) {
            const int nt = item / d.ks, ksl = item - nt * d.ks;
            const int kc0 = ksl * d.chunks;
            if (resident && item != cta) break;
            for (int c = 0; c < d.chunks; ++c) {
              mbar_wait(&emptyA[stage], phase ^ 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(&fullA[stage], a_stage_bytes);
                tma_load_2d(&p.amap[d.amap], &fullA[stage], a_ring + stage * a_stage_bytes, (kc0 + c) * 64, 0);
              }
              __syncwarp();
              if (++stage == a_stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================================= MMA issuer
    int ws = 0; uint32_t wph = 0;
    int as = 0; uint32_t aph = 0;
    int acc = 0; uint32_t accph = 0;
    const uint32_t w_lo = (smem_u32(w_ring) & 0x3FFFFu) >> 4, a_lo = (smem_u32(a_ring) & 0x3FFFFu) >> 4;
    constexpr uint64_t desc_hi = umma_desc_sw128_hi();
    for (int s = 0; s < p.n_steps; ++s) {
#pragma unroll 1
      for (int l = 0; l <= L; ++l) {
#pragma unroll 1
        for (int g = (l < L ? 0 : 4); g < (l < L ? 4 : 5); ++g) {
          const Gd d = get_gd(p, l, g);
          if (gemm_phase_index(p, s, l, g) >= p.max_phases) goto roles_done;
          const uint32_t idesc = umma_idesc_bf16(128, d.NT);
          const bool resident = d.ks == 1 && d.chunks == a_stages;
          const int as0 = as;
          for (int item = cta; item < d.n_items; item += G) {
            const bool first_item = item == cta, last_item = item + G >= d.n_items;
            mbar_wait(&accempty[acc], accph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int c = 0; c < d.chunks; ++c) {
              const int wsub = c % d.cps;
              if (wsub == 0) mbar_wait(&fullW[ws], wph);
              int ast = as;
              if (resident) { ast = as0 + c; if (ast >= a_stages) ast -= a_stages; }
              if (!resident || first_item) mbar_wait(&fullA[ast], aph);
              tc_fence_after();
              const bool w_done = (wsub == d.cps - 1) || (c == d.chunks - 1);
              if (elect_one()) {
                const uint64_t da = desc_hi | (a_lo + ast * (a_stage_bytes >> 4));
                const uint64_t db = desc_hi | (w_lo + ws * (DK_STAGE_BYTES >> 4) + wsub * (d.NT * 128 >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (c | k) != 0);
                if (!resident || last_item) tc_commit(&emptyA[ast]);
                if (w_done) tc_commit(&emptyW[ws]);
                if (c == d.chunks - 1) tc_commit(&accfull[acc]);
              }
              __syncwarp();
              if (!resident || first_item) { if (++as == a_stages) { as = 0; aph ^= 1; } }
              if (w_done) { if (++ws == DK_W_STAGES) { ws = 0; wph ^= 1; } }
            }
            if (++acc == 2) { acc = 0; accph ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================= epilogue + row phases (128 threads)
    const int wq = warp - 4;                 // TMEM lane quarter
    const int te = threadIdx.x - 128;
    const int row = wq * 32 + lane;          // sequence row of this thread in GEMM epilogues
    const bool row_ok = row < p.n_seq;
    const bool warp_has_rows = wq * 32 < p.n_seq;
    int acc = 0; uint32_t accph = 0;
    unsigned int ph = 0;                     // phases completed by this CTA

    // End of a phase for this CTA.  stores -> (cross-proxy fence for TMA readers) -> CTA barrier -> one red.release by
    // thread 0 (cumulative: covers the other threads' stores ordered before it by the barrier).
    // The barrier is ONE monotonic counter: an arrive for phase `ph` may only be posted once every CTA has finished
    // phase ph-1, otherwise early arrivals would be counted towards the previous phase's target.  CTAs that did work
    // in this phase have already waited for that (their loads were gated on it); idle ones must wait here.
    auto phase_done = [&](bool waited) {
      fence_proxy_async();
      bar_sync_epi();
      if (te == 0) {
        if (!waited) grid_wait(p.bar, ph);
        if (p.prof != nullptr && cta == 0 && ph < 4000) {
          unsigned long long t;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
          p.prof[ph] = t;
        }
        grid_arrive(p.bar);
      }
      ++ph;
    };
    auto wait_prev = [&]() {                 // all CTAs finished every earlier phase
      if (te == 0) grid_wait(p.bar, ph);
      bar_sync_epi();
    };

#define DK_STOP_CHECK() do { if (static_cast<int>(ph) >= p.max_phases) goto roles_done; } while (0)
    // ---- phase 0: first input row
    const RowCtx rc{p.h, p.xn, p.part, p.n_seq, p.H};
    if (cta < p.n_seq) row_embed(rc, p.wpe, p.layer[0].ln1_g, p.layer[0].ln1_b, p.emb + static_cast<size_t>(cta) * p.H, nullptr, p.past0, s_red, te);
    phase_done(true);

    float best_v = -INFINITY; int best_i = 0x7fffffff;

    auto gemm_epilogue = [&](const Gd& d, int s) -> bool {
      const int n_groups = d.NT / 16;
      for (int item = cta; item < d.n_items; item += G) {
        const int nt = item / d.ks, ksl = item - nt * d.ks;
        const int n0 = nt * d.NT;
        mbar_wait(&accfull[acc], accph);
        tc_fence_after();
        if (warp_has_rows) {
          for (int cg = 0; cg < n_groups; ++cg) {
            uint32_t r[16];
            tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * 256 + cg * 16, r);
            tmem_ld_wait();
            const int col = n0 + cg * 16;
            if (d.g == G_FC1) {
              const float4* b4 = reinterpret_cast<const float4*>(p.layer[d.layer].fc_b + col);
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b = __ldg(b4 + j);
                o[2 * j] = pack_bf16(gelu_tanh_fast(__uint_as_float(r[4 * j]) + b.x), gelu_tanh_fast(__uint_as_float(r[4 * j + 1]) + b.y));
                o[2 * j + 1] = pack_bf16(gelu_tanh_fast(__uint_as_float(r[4 * j + 2]) + b.z), gelu_tanh_fast(__uint_as_float(r[4 * j + 3]) + b.w));
              }
              if (row_ok) {
                uint4* dst = reinterpret_cast<uint4*>(p.hid + static_cast<size_t>(row) * d.N + col);
                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
              }
            } else if (d.g == G_LMHEAD) {
              if (row_ok) {
                if (p.logits != nullptr) {
                  float4* dst = reinterpret_cast<float4*>(p.logits + s * p.logits_step_stride + static_cast<size_t>(row) * p.vocab_pad + col);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float v = __uint_as_float(r[j]);
                  if (col + j < p.vocab && v > best_v) { best_v = v; best_i = col + j; }
                }
              }
            } else if (row_ok) {   // fp32 partial [ksl][row][col]
              float4* dst = reinterpret_cast<float4*>(p.part + (static_cast<size_t>(ksl) * p.n_seq + row) * d.N + col);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&accempty[acc]);
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
      return cta < d.n_items;
    };

    // One compact loop over the phases of a step with a single call site per phase kind: the code of a phase runs once per
    // phase, i.e. always from a cold instruction cache, so code size is latency here (the first version of this loop,
    // with every phase inlined at its own call site, was 64 KB of SASS and spent microseconds per phase on fetch).
#pragma unroll 1
    for (int s = 0; s < p.n_steps; ++s) {
      const int past = p.past0 + s;
#pragma unroll 1
      for (int k = 0; k <= 7 * L; ++k) {
        DK_STOP_CHECK();
        const bool is_lm = k == 7 * L;
        const int l = is_lm ? L : k / 7;
        const int kk = is_lm ? 0 : k - 7 * l;
        bool waited = true;
        if (kk == 0 || kk == 2 || kk == 4 || kk == 5) {
          const int g = is_lm ? G_LMHEAD : (kk == 0 ? G_QKV : kk == 2 ? G_APROJ : kk == 4 ? G_FC1 : G_FC2);
          if (is_lm) { best_v = -INFINITY; best_i = 0x7fffffff; }
          waited = gemm_epilogue(get_gd(p, l, g), s);
          if (is_lm && p.greedy && row_ok) {
            p.cand_v[static_cast<size_t>(cta) * 128 + row] = best_v;
            p.cand_i[static_cast<size_t>(cta) * 128 + row] = best_i;
          }
        } else {
          wait_prev();
          if (kk == 1) {
            row_attention(AttCtx{p.part, p.layer[l].attn_b, p.kv, p.slot, p.att, p.n_seq, p.H, p.heads, p.cache_n_seq, p.s_max, p.ks[G_QKV]}, l,
                          past, s_p, s_q, wq, lane);
          } else {
            const bool post_attn = kk == 3, last = l + 1 == L;
            const float* gamma = post_attn ? p.layer[l].ln2_g : (last ? p.lnf_g : p.layer[l + 1].ln1_g);
            const float* beta = post_attn ? p.layer[l].ln2_b : (last ? p.lnf_b : p.layer[l + 1].ln1_b);
            row_resid_ln(rc, p.ks[post_attn ? G_APROJ : G_FC2], post_attn ? p.layer[l].aproj_b : p.layer[l].mproj_b, gamma, beta, s_red, te);
          }
        }
        phase_done(waited);
      }
      if (p.greedy) {
        DK_STOP_CHECK();
        wait_prev();
        row_select(p, s, past, s_red, s_tok, te);
        phase_done(true);
      }
    }
  }

roles_done:
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_dk_sms = 0;
bool g_dk_attr = false;

// (NT, K-split) per GEMM: ~one wave of work items, then the least L2 traffic per item.
void plan_gemm(int N, int K, int n_seq, bool allow_split, int G, int* nt_out, int* ks_out) {
  double best = 1e30;
  int bnt = 16, bks = 1;
  const int kc = K / 64;
  for (int nt = 16; nt <= 128; nt *= 2) {
    if (N % nt) continue;
    for (int ks = 1; ks <= kc; ++ks) {
      if (kc % ks || ks > 12) continue;          // consumers keep at most DK_MAX_KS = 12 partial loads in flight
      if (!allow_split && ks != 1) continue;
      const int items = (N / nt) * ks;
      const int waves = (items + G - 1) / G;
      const double KS = static_cast<double>(K) / ks;
      const double bytes = KS * (n_seq + nt) * 2.0 + (ks > 1 || allow_split ? static_cast<double>(n_seq) * nt * 8.0 : static_cast<double>(n_seq) * nt * 2.0);
      const double cost = waves * (bytes + 4096.0);
      if (cost < best) { best = cost; bnt = nt; bks = ks; }
    }
  }
  *nt_out = bnt; *ks_out = bks;
}

}  // namespace

void decode_plan(const VcGptWeights* w, int* nt, int* ks) {
  const int H = w->dim, G = 148, nominal = 64;
  plan_gemm(3 * H, H, nominal, true, G, &nt[G_QKV], &ks[G_QKV]);
  plan_gemm(H, H, nominal, true, G, &nt[G_APROJ], &ks[G_APROJ]);
  plan_gemm(4 * H, H, nominal, false, G, &nt[G_FC1], &ks[G_FC1]);
  plan_gemm(H, 4 * H, nominal, true, G, &nt[G_FC2], &ks[G_FC2]);
  nt[G_LMHEAD] = 128; ks[G_LMHEAD] = 1;
}

size_t decode_partial_floats_per_row(const VcGptWeights* w) {
  int nt[5], ks[5];
  decode_plan(w, nt, ks);
  const size_t H = w->dim;
  return std::max({static_cast<size_t>(ks[G_QKV]) * 3 * H, static_cast<size_t>(ks[G_APROJ]) * H, static_cast<size_t>(ks[G_FC2]) * H});
}

bool decode_supported(const VcGptWeights* w, int n_seq, const VcKvCache* c) {
  // (the plan keeps every K-split <= DK_MAX_KS: K/64 <= 64 and items <= ~148)
  static const bool disabled = getenv("VC_DECODE_FALLBACK") != nullptr;   // A/B aid: per-op kernels instead of the persistent one
  if (disabled) return false;
  return w->layers <= DK_MAX_LAYERS && w->dim % 256 == 0 && w->dim <= 1024 && w->dim == w->heads * 64 && n_seq >= 1 && n_seq <= 128 &&
         w->vocab_pad % 128 == 0 && c->head_dim == 64 && c->s_max <= DK_MAX_S;
}

int decode_steps(const VcGptWeights* w, const DecodeBuffers& b, const VcKvCache* cache, int n_seq, int past0, int n_steps, const float* emb,
                 const DecodeGreedy* greedy, float* logits, long long logits_step_stride, cudaStream_t stream) {
  VC_REQUIRE(decode_supported(w, n_seq, cache), "decode_steps: unsupported shape (layers=%d dim=%d n_seq=%d)", w->layers, w->dim, n_seq);
  VC_REQUIRE(past0 >= 1 && past0 + n_steps <= cache->s_max, "decode_steps: positions %d..%d exceed cache s_max=%d", past0, past0 + n_steps, cache->s_max);
  VC_REQUIRE(greedy != nullptr || (n_steps == 1 && logits != nullptr), "decode_steps: a plain forward is one step with a logits buffer");
  if (g_dk_sms == 0) {
    int dev = 0;
    VC_CUDA_OK(cudaGetDevice(&dev));
    VC_CUDA_OK(cudaDeviceGetAttribute(&g_dk_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (!g_dk_attr) {
    VC_CUDA_OK(cudaFuncSetAttribute(gpt2_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DK_SMEM));
    g_dk_attr = true;
  }
  static DkParams p;   // 15 KB: keep it off the stack; filled and consumed by the launch below (by-value copy at launch)
  const int H = w->dim, L = w->layers;
  decode_plan(w, p.nt, p.ks);
  for (int l = 0; l < L; ++l) {
    const VcGptLayer& y = w->layer[l];
    int e;
    if ((e = make_tmap_bf16_kmajor(&p.wmap[l * 4 + 0], y.attn_w, 3 * H, H, p.nt[G_QKV]))) return e;
    if ((e = make_tmap_bf16_kmajor(&p.wmap[l * 4 + 1], y.aproj_w, H, H, p.nt[G_APROJ]))) return e;
    if ((e = make_tmap_bf16_kmajor(&p.wmap[l * 4 + 2], y.fc_w, 4 * H, H, p.nt[G_FC1]))) return e;
    if ((e = make_tmap_bf16_kmajor(&p.wmap[l * 4 + 3], y.mproj_w, H, 4 * H, p.nt[G_FC2]))) return e;
    p.layer[l] = DkLayer{y.ln1_g, y.ln1_b, y.attn_b, y.aproj_b, y.ln2_g, y.ln2_b, y.fc_b, y.mproj_b};
  }
  int e;
  if ((e = make_tmap_bf16_kmajor(&p.wmap[L * 4], w->wte, w->vocab_pad, H, p.nt[G_LMHEAD]))) return e;
  const int a_rows = n_seq <= 64 ? 64 : 128;
  if ((e = make_tmap_bf16_kmajor(&p.amap[0], b.xn, n_seq, H, a_rows))) return e;
  if ((e = make_tmap_bf16_kmajor(&p.amap[1], b.att, n_seq, H, a_rows))) return e;
  if ((e = make_tmap_bf16_kmajor(&p.amap[2], b.hid, n_seq, 4 * H, a_rows))) return e;
  p.H = H; p.heads = w->heads; p.layers = L; p.vocab = w->vocab; p.vocab_pad = w->vocab_pad; p.n_seq = n_seq;
  p.wpe = w->wpe; p.lnf_g = w->lnf_g; p.lnf_b = w->lnf_b; p.wte = static_cast<const __nv_bfloat16*>(w->wte);
  p.h = b.h; p.xn = static_cast<__nv_bfloat16*>(b.xn); p.att = static_cast<__nv_bfloat16*>(b.att); p.hid = static_cast<__nv_bfloat16*>(b.hid);
  p.part = b.part; p.cand_v = b.cand_v; p.cand_i = b.cand_i;
  p.kv = static_cast<__nv_bfloat16*>(cache->kv); p.slot = cache->slot; p.cache_n_seq = cache->n_seq; p.s_max = cache->s_max;
  p.bar = b.bar; p.emb = emb; p.past0 = past0; p.n_steps = n_steps;
  p.greedy = greedy != nullptr ? 1 : 0;
  if (greedy != nullptr) {
    p.step0 = greedy->step0; p.max_new = greedy->max_new; p.eos = greedy->eos; p.finished = greedy->finished; p.ids_out = greedy->ids_out;
    p.len_out = greedy->len_out; p.forced = greedy->forced; p.next_ids = greedy->next_ids;
  } else {
    p.step0 = 0; p.max_new = 0; p.eos = 0; p.finished = nullptr; p.ids_out = nullptr; p.len_out = nullptr; p.forced = nullptr; p.next_ids = nullptr;
  }
  p.logits = logits; p.logits_step_stride = logits_step_stride;
  {
    p.prof = getenv("VC_DK_PROF") != nullptr ? reinterpret_cast<unsigned long long*>(b.cand_i + 148 * 128) : nullptr;
    const char* stop = getenv("VC_DK_STOP");
    p.max_phases = stop != nullptr ? atoi(stop) : 0x7fffffff;
  }
  VC_CUDA_OK(cudaMemsetAsync(b.bar, 0, 128, stream));
  {
    KernelScope ks("gpt2_decode_steps", 0.0, stream);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g_dk_sms);
    cfg.blockDim = dim3(DK_THREADS);
    cfg.dynamicSmemBytes = DK_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the phases are separated by a device-wide barrier
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    VC_CUDA_OK(cudaLaunchKernelEx(&cfg, gpt2_decode_kernel, p));
  }
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vc
