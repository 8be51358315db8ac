// Lean persistent GPT-2 decode kernel: all decode steps of a greedy loop (or one forward step) in ONE cooperative
// launch of one 12-warp CTA per SM, every warp symmetric.  Second take on decode_step.cu after measuring it: a decode
// phase moves ~20 KB per CTA, so its cost is the number of DEPENDENT L2 round trips (~0.8 us each on B200) plus the
// device-wide barrier (~1.3 us), not bandwidth.  Each phase is therefore written as "issue every load at once, then
// compute":
//   * GEMM phase: a CTA owns FG 16-feature tiles x one K slice; its warps split the slice KG ways.  The weight rows of a
//     warp (<= 2 batches of 64 k) are loaded into registers BEFORE the preceding barrier (weights are constants), the
//     activation slice is staged once per CTA in shared memory, mma.sync m16n8k16 runs out of registers/smem, the KG
//     partial tiles are summed through shared memory and written with a fused epilogue (bias / gelu_new / fp32 partial).
//   * attention: one warp per (sequence, head); q/k/v of the new token, the K rows and the V rows are all requested
//     before the first use (online softmax over chunks of 32 cached positions).
//   * residual + LayerNorm rows: one CTA per row, residual, partial sums, bias, gamma and beta loaded together.
// Phases per layer (7): QKV -> attention -> proj -> +LN2 -> fc1(+bias, gelu_new) -> fc2 -> +LN1(next); then lm_head and
// token selection.  n_seq <= 64.  Arithmetic: transformers GPT2Model, benchmark_baseline.py:210-227.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <cstdio>
#include <cstdlib>

namespace vc {

namespace {

constexpr int LK_THREADS = 384;
constexpr int LK_WARPS = 12;
constexpr int LK_MAX_LAYERS = 48;
constexpr int LK_MAX_KS = 6;
constexpr int LK_MAX_S = 1024;
constexpr int LK_MAX_KSLICE = 1024;                         // staged activation slice: 64 rows x <= 1024 k
constexpr int LK_XS_BYTES = 64 * (LK_MAX_KSLICE * 2 + 64);  // row pitch = slice bytes + 64 (conflict-free 128-bit reads)
constexpr int LK_RED_PITCH = 64 * 17;                       // one warp's 16 x 64 tile, m-major with a pad
constexpr int LK_RED_BYTES = LK_WARPS * LK_RED_PITCH * 4;
constexpr int LK_SMEM = LK_XS_BYTES + LK_RED_BYTES + 1024;

struct LkLayer {
  const __nv_bfloat16 *attn_w, *aproj_w, *fc_w, *mproj_w;
  const float *ln1_g, *ln1_b, *attn_b, *aproj_b, *ln2_g, *ln2_b, *fc_b, *mproj_b;
};
struct LkPlan { int FG, KG, ks; };     // feature tiles per CTA, warps per tile (K split inside the CTA), K slices across CTAs

struct LkParams {
  LkLayer layer[LK_MAX_LAYERS];
  LkPlan plan[5];                      // qkv, proj, fc1, fc2, lm_head
  int H, heads, layers, vocab, vocab_pad, n_seq;
  const float* wpe; const float* lnf_g; const float* lnf_b; const __nv_bfloat16* wte;
  float* h; __nv_bfloat16* xn; __nv_bfloat16* att; __nv_bfloat16* hid; float* part; float* logits_ws;
  __nv_bfloat16* kv; const int32_t* slot; int cache_n_seq, s_max;
  unsigned int* bar;
  const float* emb; int past0, n_steps;
  int greedy, step0, max_new, eos;
  int32_t* finished; int32_t* ids_out; int32_t* len_out; const int32_t* forced; int32_t* next_ids;
  float* logits; long long logits_step_stride;
  unsigned long long* prof;
};

struct GemmDesc {
  const __nv_bfloat16* W; const __nv_bfloat16* x;
  int N, K, FG, KG, ks, mode;          // mode 0: fp32 partial [ksl][row][N]; 1: bf16 acc+bias; 2: bf16 gelu_new(acc+bias); 3: fp32 [row][N]
  const float* bias; void* out;
};
struct WarpJob {
  const __nv_bfloat16* w_lo; int nb; int xoff; int tile; int kbase; bool active;
};

__device__ __forceinline__ uint4 ldg_w(const void* p) {      // weights: immutable, streamed once per step
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// debug timeline (VC_DK_PROF): CTA 0 thread 0 records (tag, globaltimer) pairs while s_st[0] != 0
__shared__ int s_st[2];
__device__ __forceinline__ void stamp(const LkParams& p, int tag) {
  if (p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && s_st[0] != 0 && s_st[1] < 400) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.prof[3000 + 2 * s_st[1]] = tag;
    p.prof[3001 + 2 * s_st[1]] = t;
    s_st[1] += 1;
  }
}

// ---------------------------------------------------------------- device-wide barrier: every CTA arrives and waits each phase
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const unsigned int target = epoch * gridDim.x;
    unsigned int v;
    unsigned int spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++spins > (1u << 26)) __trap();
    } while (v < target);
  }
  __syncthreads();
}

// ---------------------------------------------------------------- GEMM phase pieces
__device__ __forceinline__ WarpJob job_for(const GemmDesc& d, int grp, int ksl, int warp, int lane) {
  WarpJob j;
  const int fg = warp / d.KG, kg = warp - fg * d.KG;
  const int tile = grp * d.FG + fg;
  const int kslice = d.K / d.ks;
  j.nb = (kslice >> 6) / d.KG;
  j.xoff = kg * j.nb * 64;
  j.active = fg < d.FG && tile < (d.N >> 4);
  j.tile = j.active ? tile : 0;
  j.kbase = ksl * kslice + j.xoff;
  const int g = lane >> 2, t = lane & 3;
  j.w_lo = d.W + static_cast<size_t>(j.tile * 16 + g) * d.K + j.kbase + 8 * t;
  return j;
}
__device__ __forceinline__ void load_batch(uint4 (&w)[4], const WarpJob& j, int K, int b) {
  const __nv_bfloat16* lo = j.w_lo + b * 64;
  const __nv_bfloat16* hi = lo + static_cast<size_t>(8) * K;
  w[0] = ldg_w(lo);      w[1] = ldg_w(hi);
  w[2] = ldg_w(lo + 32); w[3] = ldg_w(hi + 32);
}
// Ask L2 for the weight rows of this CTA's first unit of a later GEMM phase (issued a phase ahead: the rows arrive from
// HBM while the current phase and its barrier run, so the GEMM phase's own loads are L2 hits that overlap the x staging).
// Nothing is held in registers across the barrier: a pending load that gets spilled stalls the warp for the full latency.
__device__ __forceinline__ void prefetch_w(const GemmDesc& d, int warp, int lane) {
  const int groups = ((d.N >> 4) + d.FG - 1) / d.FG, units = groups * d.ks;
  const int u = blockIdx.x;
  if (u >= units) return;
  const int grp = u / d.ks, ksl = u - grp * d.ks;
  const WarpJob j = job_for(d, grp, ksl, warp, lane);
  if (!j.active) return;
  // 16 rows x nb 128-byte lines; lane -> (row = lane >> 1, first line = lane & 1)
  const __nv_bfloat16* row = d.W + static_cast<size_t>(j.tile * 16 + (lane >> 1)) * d.K + j.kbase;
  for (int c = lane & 1; c < j.nb; c += 2) prefetch_l2(row + c * 64);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void gemm_phase(const LkParams& p, const GemmDesc& d, uint8_t* xs, float* red, int M, int warp, int lane) {
  const int tid = threadIdx.x, g = lane >> 2, t = lane & 3;
  const int tiles = d.N >> 4, groups = (tiles + d.FG - 1) / d.FG, units = groups * d.ks;
  const int kslice = d.K / d.ks, pitch = kslice * 2 + 64, cpr = kslice >> 3;
  int staged = -1;
  bool first = true;
  uint4 wq[2][4];
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int grp = u / d.ks, ksl = u - grp * d.ks;
    const WarpJob j = job_for(d, grp, ksl, warp, lane);
    if (j.active) {
      load_batch(wq[0], j, d.K, 0);
      if (j.nb > 1) load_batch(wq[1], j, d.K, 1);
      if (!first || j.nb > 2) {                        // not prefetched a phase ahead: at least get the lines moving
        const __nv_bfloat16* row = d.W + static_cast<size_t>(j.tile * 16 + (lane >> 1)) * d.K + j.kbase;
        for (int c = 2 + (lane & 1); c < j.nb; c += 2) prefetch_l2(row + c * 64);
      }
    }
    if (ksl != staged) {
      if (!first) __syncthreads();
      // stage x[0..64) x [ksl*kslice, +kslice): asynchronous 16-byte copies, warp w takes rows w, w+12, ...
      const __nv_bfloat16* xsrc = d.x + ksl * kslice;
      for (int r = warp; r < 64; r += LK_WARPS) {
        const int rr = r < M ? r : M - 1;
        const uint4* src = reinterpret_cast<const uint4*>(xsrc + static_cast<size_t>(rr) * d.K);
        for (int col = lane; col < cpr; col += 32) cp_async16(xs + r * pitch + col * 16, src + col);
      }
      staged = ksl;
    }
    // the bias values of the outputs this thread will write (modes 1, 2), requested before the wait
    float bias_r[6];
    if (d.mode == 1 || d.mode == 2) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int idx = tid + i * LK_THREADS;
        const int tile = grp * d.FG + (idx >> 10);
        bias_r[i] = (idx < d.FG * 1024 && tile < tiles) ? __ldg(d.bias + tile * 16 + (idx & 15)) : 0.f;
      }
    }
    stamp(p, 13);
    cp_async_wait_all();
    stamp(p, 14);
    __syncthreads();
    stamp(p, 11);
    if (j.active) {
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      const uint8_t* xw = xs + g * pitch + (j.xoff + 8 * t) * 2;
      auto compute = [&](const uint4 (&w)[4], int b) {
#pragma unroll
        for (int st = 0; st < 2; ++st) {
          const uint4 a = w[2 * st], c = w[2 * st + 1];
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            const uint4 xb = *reinterpret_cast<const uint4*>(xw + nt * 8 * pitch + (b * 64 + st * 32) * 2);
            mma16816(acc[nt], a.x, c.x, a.y, c.y, xb.x, xb.y);
            mma16816(acc[nt], a.z, c.z, a.w, c.w, xb.z, xb.w);
          }
        }
      };
      for (int b = 0; b < j.nb; b += 2) {
        compute(wq[0], b);
        if (b + 2 < j.nb) load_batch(wq[0], j, d.K, b + 2);
        if (b + 1 < j.nb) {
          compute(wq[1], b + 1);
          if (b + 3 < j.nb) load_batch(wq[1], j, d.K, b + 3);
        }
      }
      stamp(p, 15);
      float* rw = red + warp * LK_RED_PITCH;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int m = nt * 8 + 2 * t;
        rw[m * 17 + g] = acc[nt][0];
        rw[(m + 1) * 17 + g] = acc[nt][1];
        rw[m * 17 + g + 8] = acc[nt][2];
        rw[(m + 1) * 17 + g + 8] = acc[nt][3];
      }
      stamp(p, 16);
    }
    __syncthreads();
    stamp(p, 12);
    // epilogue: sum the KG partial tiles, fused output
    if (d.mode == 1 || d.mode == 2) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int idx = tid + i * LK_THREADS;
        const int fg = idx >> 10, m = (idx >> 4) & 63, f = idx & 15;
        const int tile = grp * d.FG + fg;
        if (idx < d.FG * 1024 && tile < tiles && m < M) {
          const float* rr = red + (fg * d.KG) * LK_RED_PITCH + m * 17 + f;
          float a = rr[0];
          for (int kg = 1; kg < d.KG; ++kg) a += rr[kg * LK_RED_PITCH];
          a += bias_r[i];
          if (d.mode == 2) a = gelu_tanh(a);
          static_cast<__nv_bfloat16*>(d.out)[static_cast<size_t>(m) * d.N + tile * 16 + f] = __float2bfloat16_rn(a);
        }
      }
    } else {
      for (int idx = tid; idx < d.FG * 1024; idx += LK_THREADS) {
        const int fg = idx >> 10, m = (idx >> 4) & 63, f = idx & 15;
        const int tile = grp * d.FG + fg;
        if (tile >= tiles || m >= M) continue;
        const float* rr = red + (fg * d.KG) * LK_RED_PITCH + m * 17 + f;
        float a = rr[0];
        for (int kg = 1; kg < d.KG; ++kg) a += rr[kg * LK_RED_PITCH];
        const int col = tile * 16 + f;
        if (d.mode == 0) static_cast<float*>(d.out)[(static_cast<size_t>(ksl) * M + m) * d.N + col] = a;
        else static_cast<float*>(d.out)[static_cast<size_t>(m) * d.N + col] = a;
      }
    }
    first = false;
  }
}

// ---------------------------------------------------------------- row phases (one CTA per sequence row)
__device__ __forceinline__ float cta_sum(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();                                   // s_red free
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < LK_WARPS; ++w) t += s_red[w];
  return t;
}

// thread owns float2 columns c0 = 2*tid and (H > 768) c1 = 768 + 2*tid.  a0/a1 = new residual values; writes h and xn
__device__ __forceinline__ void ln_finish(const LkParams& p, int r, float2 a0, float2 a1, float2 g0, float2 b0, float2 g1, float2 b1, bool v0, bool v1,
                                          float* s_red) {
  const int H = p.H, c0 = 2 * threadIdx.x, c1 = 768 + 2 * threadIdx.x;
  float sum = 0.f;
  if (v0) { *reinterpret_cast<float2*>(p.h + static_cast<size_t>(r) * H + c0) = a0; sum += a0.x + a0.y; }
  if (v1) { *reinterpret_cast<float2*>(p.h + static_cast<size_t>(r) * H + c1) = a1; sum += a1.x + a1.y; }
  const float mean = cta_sum(sum, s_red) / static_cast<float>(H);
  float sq = 0.f;
  if (v0) { const float x = a0.x - mean, y = a0.y - mean; sq += x * x + y * y; }
  if (v1) { const float x = a1.x - mean, y = a1.y - mean; sq += x * x + y * y; }
  const float rstd = rsqrtf(cta_sum(sq, s_red) / static_cast<float>(H) + 1e-5f);
  if (v0) *reinterpret_cast<uint32_t*>(p.xn + static_cast<size_t>(r) * H + c0) = pack_bf16((a0.x - mean) * rstd * g0.x + b0.x, (a0.y - mean) * rstd * g0.y + b0.y);
  if (v1) *reinterpret_cast<uint32_t*>(p.xn + static_cast<size_t>(r) * H + c1) = pack_bf16((a1.x - mean) * rstd * g1.x + b1.x, (a1.y - mean) * rstd * g1.y + b1.y);
}

// h[r] += bias + sum_s P[s][r]; xn[r] = LN(h[r])
__device__ __forceinline__ void phase_resid_ln(const LkParams& p, int ks, const float* bias, const float* gamma, const float* beta, float* s_red) {
  const int H = p.H, c0 = 2 * threadIdx.x, c1 = 768 + 2 * threadIdx.x;
  const bool v0 = c0 < H && c0 < 768, v1 = c1 < H;
  const size_t plane = static_cast<size_t>(p.n_seq) * H;
  const float2 z = make_float2(0.f, 0.f);
  for (int r = blockIdx.x; r < p.n_seq; r += gridDim.x) {
    const size_t o0 = static_cast<size_t>(r) * H + c0, o1 = static_cast<size_t>(r) * H + c1;
    float2 q0[LK_MAX_KS], q1[LK_MAX_KS];
    float2 a0 = z, a1 = z, bb0 = z, bb1 = z, g0 = z, g1 = z, e0 = z, e1 = z;
    if (v0) {
      a0 = __ldcg(reinterpret_cast<const float2*>(p.h + o0));
#pragma unroll
      for (int s = 0; s < LK_MAX_KS; ++s)
        if (s < ks) q0[s] = __ldcg(reinterpret_cast<const float2*>(p.part + s * plane + o0));
      bb0 = __ldg(reinterpret_cast<const float2*>(bias + c0));
      g0 = __ldg(reinterpret_cast<const float2*>(gamma + c0));
      e0 = __ldg(reinterpret_cast<const float2*>(beta + c0));
    }
    if (v1) {
      a1 = __ldcg(reinterpret_cast<const float2*>(p.h + o1));
#pragma unroll
      for (int s = 0; s < LK_MAX_KS; ++s)
        if (s < ks) q1[s] = __ldcg(reinterpret_cast<const float2*>(p.part + s * plane + o1));
      bb1 = __ldg(reinterpret_cast<const float2*>(bias + c1));
      g1 = __ldg(reinterpret_cast<const float2*>(gamma + c1));
      e1 = __ldg(reinterpret_cast<const float2*>(beta + c1));
    }
    if (v0) {
      a0.x += bb0.x; a0.y += bb0.y;
#pragma unroll
      for (int s = 0; s < LK_MAX_KS; ++s)
        if (s < ks) { a0.x += q0[s].x; a0.y += q0[s].y; }
    }
    if (v1) {
      a1.x += bb1.x; a1.y += bb1.y;
#pragma unroll
      for (int s = 0; s < LK_MAX_KS; ++s)
        if (s < ks) { a1.x += q1[s].x; a1.y += q1[s].y; }
    }
    ln_finish(p, r, a0, a1, g0, e0, g1, e1, v0, v1, s_red);
  }
}

// first input row of a step: h = embedding + wpe[pos]; xn = LN1_0(h).  emb_row (fp32) or wte_row (bf16)
__device__ __forceinline__ void row_embed(const LkParams& p, int r, const float* emb_row, const __nv_bfloat16* wte_row, int pos, float* s_red) {
  const int H = p.H, c0 = 2 * threadIdx.x, c1 = 768 + 2 * threadIdx.x;
  const bool v0 = c0 < H && c0 < 768, v1 = c1 < H;
  const float2 z = make_float2(0.f, 0.f);
  float2 a0 = z, a1 = z, w0 = z, w1 = z, g0 = z, g1 = z, e0 = z, e1 = z;
  const float* gamma = p.layer[0].ln1_g;
  const float* beta = p.layer[0].ln1_b;
  if (v0) {
    a0 = emb_row != nullptr ? __ldcg(reinterpret_cast<const float2*>(emb_row + c0)) : unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(wte_row + c0)));
    w0 = __ldg(reinterpret_cast<const float2*>(p.wpe + static_cast<size_t>(pos) * H + c0));
    g0 = __ldg(reinterpret_cast<const float2*>(gamma + c0));
    e0 = __ldg(reinterpret_cast<const float2*>(beta + c0));
  }
  if (v1) {
    a1 = emb_row != nullptr ? __ldcg(reinterpret_cast<const float2*>(emb_row + c1)) : unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(wte_row + c1)));
    w1 = __ldg(reinterpret_cast<const float2*>(p.wpe + static_cast<size_t>(pos) * H + c1));
    g1 = __ldg(reinterpret_cast<const float2*>(gamma + c1));
    e1 = __ldg(reinterpret_cast<const float2*>(beta + c1));
  }
  a0.x += w0.x; a0.y += w0.y; a1.x += w1.x; a1.y += w1.y;
  ln_finish(p, r, a0, a1, g0, e0, g1, e1, v0, v1, s_red);
}

// ---------------------------------------------------------------- attention: one warp per (sequence, head), online softmax over chunks of 32
__device__ __forceinline__ void phase_attention(const LkParams& p, const __nv_bfloat16* qkv, int layer, int past, float* sq_all, int warp, int lane) {
  const int heads = p.heads, H = p.H;
  const int n_items = p.n_seq * heads, n_warps = gridDim.x * LK_WARPS;
  const size_t plane = static_cast<size_t>(p.cache_n_seq) * heads * p.s_max * 64;
  __nv_bfloat16* kbase = p.kv + (static_cast<size_t>(layer) * 2 + 0) * plane;
  __nv_bfloat16* vbase = p.kv + (static_cast<size_t>(layer) * 2 + 1) * plane;
  float* sq = sq_all + warp * 64;
  for (int item = warp * gridDim.x + blockIdx.x; item < n_items; item += n_warps) {
    const int seq = item / heads, head = item - seq * heads;
    const __nv_bfloat16* row = qkv + static_cast<size_t>(seq) * 3 * H + head * 64 + 2 * lane;
    const uint32_t qu = __ldcg(reinterpret_cast<const uint32_t*>(row));
    const uint32_t ku = __ldcg(reinterpret_cast<const uint32_t*>(row + H));
    const uint32_t vu = __ldcg(reinterpret_cast<const uint32_t*>(row + 2 * H));
    const size_t seq_base = static_cast<size_t>(seq) * p.s_max;
    const size_t own = (static_cast<size_t>(seq) * heads + head) * p.s_max * 64;
    float m_run = -INFINITY, l_run = 0.f, o0 = 0.f, o1 = 0.f;
    bool q_ready = false;
    float2 q2 = make_float2(0.f, 0.f), k2 = q2, v2 = q2;
    for (int j0 = 0; j0 < past; j0 += 32) {
      const int n = min(32, past - j0);
      // every K row (lane j owns position j0+j) and every V pair (lane owns dims 2*lane, 2*lane+1) requested up front
      int phys_l = seq;
      if (p.slot != nullptr && lane < n) phys_l = p.slot[seq_base + j0 + lane];
      uint4 kr[8];
      if (lane < n) {
        const uint4* kp = reinterpret_cast<const uint4*>(kbase + (static_cast<size_t>(phys_l) * heads + head) * p.s_max * 64 + static_cast<size_t>(j0 + lane) * 64);
#pragma unroll
        for (int c = 0; c < 8; ++c) kr[c] = __ldcg(kp + c);
      }
      uint32_t vr[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i < n) {
          const int phys = p.slot != nullptr ? __shfl_sync(0xffffffffu, phys_l, i) : seq;
          vr[i] = __ldcg(reinterpret_cast<const uint32_t*>(vbase + (static_cast<size_t>(phys) * heads + head) * p.s_max * 64 + static_cast<size_t>(j0 + i) * 64) + lane);
        }
      }
      if (!q_ready) {
        q2 = unpack_bf16(qu); k2 = unpack_bf16(ku); v2 = unpack_bf16(vu);
        sq[2 * lane] = q2.x; sq[2 * lane + 1] = q2.y;
        __syncwarp();
        q_ready = true;
      }
      float s = -INFINITY;
      if (lane < n) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 qa = *reinterpret_cast<const float4*>(sq + c * 8), qb = *reinterpret_cast<const float4*>(sq + c * 8 + 4);
          const float2 a = unpack_bf16(kr[c].x), b = unpack_bf16(kr[c].y), cc = unpack_bf16(kr[c].z), d = unpack_bf16(kr[c].w);
          acc = fmaf(qa.x, a.x, acc); acc = fmaf(qa.y, a.y, acc); acc = fmaf(qa.z, b.x, acc); acc = fmaf(qa.w, b.y, acc);
          acc = fmaf(qb.x, cc.x, acc); acc = fmaf(qb.y, cc.y, acc); acc = fmaf(qb.z, d.x, acc); acc = fmaf(qb.w, d.y, acc);
        }
        s = acc * 0.125f;
      }
      const float m_new = fmaxf(m_run, warp_max(s));
      const float scale = __expf(m_run - m_new);           // 0 on the first chunk (m_run = -inf)
      const float pj = lane < n ? __expf(s - m_new) : 0.f;
      l_run = l_run * scale + warp_sum(pj);
      o0 *= scale; o1 *= scale;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i < n) {
          const float pi = __shfl_sync(0xffffffffu, pj, i);
          const float2 vv = unpack_bf16(vr[i]);
          o0 = fmaf(pi, vv.x, o0);
          o1 = fmaf(pi, vv.y, o1);
        }
      }
      m_run = m_new;
    }
    if (!q_ready) { q2 = unpack_bf16(qu); k2 = unpack_bf16(ku); v2 = unpack_bf16(vu); }
    // the new token itself, and its K/V row appended to the cache
    *reinterpret_cast<uint32_t*>(kbase + own + static_cast<size_t>(past) * 64 + 2 * lane) = ku;
    *reinterpret_cast<uint32_t*>(vbase + own + static_cast<size_t>(past) * 64 + 2 * lane) = vu;
    const float s_new = warp_sum(q2.x * k2.x + q2.y * k2.y) * 0.125f;
    const float m_new = fmaxf(m_run, s_new);
    const float scale = __expf(m_run - m_new), p_new = __expf(s_new - m_new);
    l_run = l_run * scale + p_new;
    o0 = fmaf(p_new, v2.x, o0 * scale);
    o1 = fmaf(p_new, v2.y, o1 * scale);
    const float inv = 1.f / l_run;
    *reinterpret_cast<uint32_t*>(p.att + static_cast<size_t>(seq) * H + head * 64 + 2 * lane) = pack_bf16(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

__device__ __forceinline__ void argmax_cmb(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// argmax of the logits row (ties -> lowest index), greedy bookkeeping, next input row
__device__ __forceinline__ void phase_select(const LkParams& p, const float* lg, int s, int past, float* s_red, int* s_tok) {
  const int tid = threadIdx.x;
  for (int r = blockIdx.x; r < p.n_seq; r += gridDim.x) {
    const float* row = lg + static_cast<size_t>(r) * p.vocab_pad;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j0 = tid * 4; j0 < p.vocab; j0 += LK_THREADS * 4 * 8) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = j0 + i * LK_THREADS * 4;
        if (j < p.vocab) v[i] = __ldcg(reinterpret_cast<const float4*>(row + j));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = j0 + i * LK_THREADS * 4;
        if (j < p.vocab) {
          if (v[i].x > bv) { bv = v[i].x; bi = j; }
          if (j + 1 < p.vocab && v[i].y > bv) { bv = v[i].y; bi = j + 1; }
          if (j + 2 < p.vocab && v[i].z > bv) { bv = v[i].z; bi = j + 2; }
          if (j + 3 < p.vocab && v[i].w > bv) { bv = v[i].w; bi = j + 3; }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      argmax_cmb(bv, bi, ov, oi);
    }
    int* s_i = reinterpret_cast<int*>(s_red + 16);
    __syncthreads();
    if ((tid & 31) == 0) { s_red[tid >> 5] = bv; s_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < LK_WARPS; ++w) argmax_cmb(bv, bi, s_red[w], s_i[w]);
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
    __syncthreads();
    const int feed = *s_tok;
    if (s + 1 < p.n_steps) row_embed(p, r, nullptr, p.wte + static_cast<size_t>(feed) * p.H, past + 1, s_red);
    __syncthreads();
  }
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(LK_THREADS, 1) gpt2_decode_lean_kernel(const __grid_constant__ LkParams p) {
  extern __shared__ __align__(128) uint8_t lk_smem[];
  uint8_t* xs = lk_smem;
  float* red = reinterpret_cast<float*>(lk_smem + LK_XS_BYTES);
  float* s_red = reinterpret_cast<float*>(lk_smem + LK_XS_BYTES + LK_RED_BYTES);     // 16 floats + 16 ints
  int* s_tok = reinterpret_cast<int*>(s_red + 64);
  float* sq_all = red;                                                               // attention: q rows (12 x 64 floats)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, M = p.n_seq, L = p.layers;
  unsigned int epoch = 0;
  unsigned int ph = 0;
  if (threadIdx.x == 0) { s_st[0] = 0; s_st[1] = 0; }
  __syncthreads();
  auto sync = [&]() {
    stamp(p, 1);
    grid_sync(p.bar, epoch);
    stamp(p, 2);
    if (p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && ph < 3000) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.prof[ph] = t;
    }
    ++ph;
  };
  __nv_bfloat16* qkv = p.hid;                      // the fc1 activation buffer is free between fc2 and the next fc1
  auto desc = [&](int l, int k) -> GemmDesc {
    if (l >= L) return GemmDesc{p.wte, p.xn, p.vocab_pad, H, p.plan[4].FG, p.plan[4].KG, 1, 3, nullptr, nullptr};
    const LkLayer& y = p.layer[l];
    if (k == 0) return GemmDesc{y.attn_w, p.xn, 3 * H, H, p.plan[0].FG, p.plan[0].KG, 1, 1, y.attn_b, qkv};
    if (k == 1) return GemmDesc{y.aproj_w, p.att, H, H, p.plan[1].FG, p.plan[1].KG, p.plan[1].ks, 0, nullptr, p.part};
    if (k == 2) return GemmDesc{y.fc_w, p.xn, 4 * H, H, p.plan[2].FG, p.plan[2].KG, 1, 2, y.fc_b, p.hid};
    return GemmDesc{y.mproj_w, p.hid, H, 4 * H, p.plan[3].FG, p.plan[3].KG, p.plan[3].ks, 0, nullptr, p.part};
  };
  if (p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.prof[2990] = clock64(); p.prof[2991] = t;
  }
  // ---- phase 0: first input rows
  prefetch_w(desc(0, 0), warp, lane);
  for (int r = blockIdx.x; r < M; r += gridDim.x) row_embed(p, r, p.emb + static_cast<size_t>(r) * H, nullptr, p.past0, s_red);
  sync();

#pragma unroll 1
  for (int s = 0; s < p.n_steps; ++s) {
    const int past = p.past0 + s;
    float* lg = p.logits != nullptr ? p.logits + s * p.logits_step_stride : p.logits_ws;
#pragma unroll 1
    for (int k = 0; k <= 7 * L; ++k) {
      const bool is_lm = k == 7 * L;
      const int l = is_lm ? L : k / 7;
      const int kk = is_lm ? 0 : k - 7 * l;
      if (threadIdx.x == 0) s_st[0] = (s == 2 && l == 5) ? 1 : 0;
      if (kk == 1 || kk == 3 || kk == 6) {
        // row phases: first request the weights of the GEMM phase that follows (they arrive during this phase and its barrier)
        stamp(p, kk == 1 ? 30 : 20);
        prefetch_w(kk == 1 ? desc(l, 1) : kk == 3 ? desc(l, 2) : desc(l + 1, 0), warp, lane);
        if (kk == 1) {
          phase_attention(p, qkv, l, past, sq_all, warp, lane);
        } else {
          const bool last = l + 1 == L;
          const LkLayer& y = p.layer[l];
          if (kk == 3) phase_resid_ln(p, p.plan[1].ks, y.aproj_b, y.ln2_g, y.ln2_b, s_red);
          else phase_resid_ln(p, p.plan[3].ks, y.mproj_b, last ? p.lnf_g : p.layer[l + 1].ln1_g, last ? p.lnf_b : p.layer[l + 1].ln1_b, s_red);
        }
      } else {
        // QKV (+bias) -> bf16 | attention projection -> partials | fc1 (+bias, gelu_new) -> bf16 | fc2 -> partials | lm_head -> fp32 logits
        stamp(p, 10);
        GemmDesc d = desc(l, kk == 0 ? 0 : kk == 2 ? 1 : kk == 4 ? 2 : 3);
        if (is_lm) d.out = lg;
        gemm_phase(p, d, xs, red, M, warp, lane);
        if (kk == 4) prefetch_w(desc(l, 3), warp, lane);
      }
      sync();
    }
    if (p.greedy) {
      if (s + 1 < p.n_steps) prefetch_w(desc(0, 0), warp, lane);
      phase_select(p, lg, s, past, s_red, s_tok);
      sync();
    }
  }
  if (p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.prof[2992] = clock64(); p.prof[2993] = t;
  }
}

bool g_lk_attr = false;
int g_lk_grid = 0;

// Choose (FG, KG, ks) for a product: FG*KG <= 12 warps, the K batches of a slice split evenly over KG warps, the staged
// slice fits shared memory; cost model in units of one L2 round trip.
LkPlan plan_gemm(int N, int K, bool whole_k, int max_fg, int n_ctas) {
  const int tiles = N / 16, nbt = K / 64;
  LkPlan best{1, 1, 1};
  double best_cost = 1e30;
  for (int ks = 1; ks <= (whole_k ? 1 : LK_MAX_KS); ++ks) {
    if (nbt % ks) continue;
    const int sb = nbt / ks;                        // batches per slice
    if (sb * 64 > LK_MAX_KSLICE) continue;
    for (int KG = 1; KG <= LK_WARPS; ++KG) {
      if (sb % KG) continue;
      for (int FG = 1; FG * KG <= LK_WARPS && FG <= max_fg; ++FG) {
        const int units = ((tiles + FG - 1) / FG) * ks;
        const int rounds = (units + n_ctas - 1) / n_ctas;
        const int nb = sb / KG;
        const double stage = (1.0 + sb * 64 / 1024.0) * (ks == 1 ? 1 : rounds);
        const double cost = stage + rounds * (0.6 + 0.4 * nb) + 0.05 * ks;
        if (cost < best_cost) { best_cost = cost; best = LkPlan{FG, KG, ks}; }
      }
    }
  }
  return best;
}

}  // namespace

bool decode_lean_supported(const VcGptWeights* w, int n_seq, const VcKvCache* c) {
  return w->layers <= LK_MAX_LAYERS && w->dim % 256 == 0 && w->dim <= 1024 && w->dim == w->heads * 64 && n_seq >= 1 && n_seq <= 64 &&
         w->vocab_pad % 16 == 0 && c->head_dim == 64 && c->s_max <= LK_MAX_S;
}

size_t decode_lean_partial_floats_per_row(const VcGptWeights* w) {
  const size_t H = w->dim;
  return static_cast<size_t>(LK_MAX_KS) * H;
}

int decode_lean_steps(const VcGptWeights* w, const DecodeBuffers& b, float* logits_ws, const VcKvCache* cache, int n_seq, int past0, int n_steps,
                      const float* emb, const DecodeGreedy* greedy, float* logits, long long logits_step_stride, cudaStream_t stream) {
  VC_REQUIRE(decode_lean_supported(w, n_seq, cache), "decode_lean: unsupported shape (layers=%d dim=%d n_seq=%d)", w->layers, w->dim, n_seq);
  VC_REQUIRE(past0 >= 1 && past0 + n_steps <= cache->s_max, "decode_lean: positions %d..%d exceed cache s_max=%d", past0, past0 + n_steps, cache->s_max);
  VC_REQUIRE(greedy != nullptr || (n_steps == 1 && logits != nullptr), "decode_lean: a plain forward is one step with a logits buffer");
  if (!g_lk_attr) {
    int dev = 0, sms = 0, per_sm = 0;
    VC_CUDA_OK(cudaGetDevice(&dev));
    VC_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VC_CUDA_OK(cudaFuncSetAttribute(gpt2_decode_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LK_SMEM));
    VC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gpt2_decode_lean_kernel, LK_THREADS, LK_SMEM));
    VC_REQUIRE(per_sm >= 1, "decode_lean: kernel does not fit an SM");
    g_lk_grid = sms;
    g_lk_attr = true;
  }
  static LkParams p;
  const int H = w->dim, L = w->layers;
  for (int l = 0; l < L; ++l) {
    const VcGptLayer& y = w->layer[l];
    p.layer[l] = LkLayer{static_cast<const __nv_bfloat16*>(y.attn_w), static_cast<const __nv_bfloat16*>(y.aproj_w),
                         static_cast<const __nv_bfloat16*>(y.fc_w), static_cast<const __nv_bfloat16*>(y.mproj_w),
                         y.ln1_g, y.ln1_b, y.attn_b, y.aproj_b, y.ln2_g, y.ln2_b, y.fc_b, y.mproj_b};
  }
  p.H = H; p.heads = w->heads; p.layers = L; p.vocab = w->vocab; p.vocab_pad = w->vocab_pad; p.n_seq = n_seq;
  p.plan[0] = plan_gemm(3 * H, H, true, 2, g_lk_grid);
  p.plan[1] = plan_gemm(H, H, false, LK_WARPS, g_lk_grid);
  p.plan[2] = plan_gemm(4 * H, H, true, 2, g_lk_grid);
  p.plan[3] = plan_gemm(H, 4 * H, false, LK_WARPS, g_lk_grid);
  p.plan[4] = plan_gemm(w->vocab_pad, H, true, LK_WARPS, g_lk_grid);
  p.wpe = w->wpe; p.lnf_g = w->lnf_g; p.lnf_b = w->lnf_b; p.wte = static_cast<const __nv_bfloat16*>(w->wte);
  p.h = b.h; p.xn = static_cast<__nv_bfloat16*>(b.xn); p.att = static_cast<__nv_bfloat16*>(b.att); p.hid = static_cast<__nv_bfloat16*>(b.hid);
  p.part = b.part; p.logits_ws = logits_ws;
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
  p.prof = getenv("VC_DK_PROF") != nullptr ? reinterpret_cast<unsigned long long*>(b.cand_i + 148 * 128) : nullptr;
  if (p.prof != nullptr) {
    static bool once = false;
    if (!once) {
      once = true;
      const char* nm[5] = {"qkv", "proj", "fc1", "fc2", "lm_head"};
      for (int i = 0; i < 5; ++i) fprintf(stderr, "decode_lean plan %s: FG=%d KG=%d ks=%d\n", nm[i], p.plan[i].FG, p.plan[i].KG, p.plan[i].ks);
    }
  }
  VC_CUDA_OK(cudaMemsetAsync(b.bar, 0, 128, stream));
  {
    KernelScope ks("gpt2_decode_lean", 0.0, stream);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g_lk_grid);
    cfg.blockDim = dim3(LK_THREADS);
    cfg.dynamicSmemBytes = LK_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    VC_CUDA_OK(cudaLaunchKernelEx(&cfg, gpt2_decode_lean_kernel, p));
  }
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vc
