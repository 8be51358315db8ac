// GPT-2 few-row forward (a decode step, or the prefill of a caption batch) as FIVE dependent kernels per layer:
//
//     QKV (full-K, LayerNorm folded) -> attention (gpt2_kernels.cu) -> proj (cluster split-K) ->
//     fc1 (full-K, LayerNorm folded, gelu_new) -> fc2 (cluster split-K)
//
// plus one lm_head kernel that keeps only (max, argmax) candidates per CTA and one selection kernel that also builds the
// next step's input row (transformers GPT2LMHeadModel.forward; loop: core/scripts/benchmark_baseline.py:160-240).
//
// Why this shape.  A step is ~100 dependent operations on 64 rows; each dependent kernel costs 3-6 us whatever it computes
// (DESIGN.md 4.5), so the step time is (number of dependent kernels) x (latency of one).  This file removes every kernel that
// only re-reads what the previous one wrote:
//   * no split-K partial buffers in global memory: products whose output is wide (N = 3H, 4H) keep the whole K inside a CTA
//     (8 warps x K/8, reduced through shared memory); products whose output is the residual stream (N = H) split K across
//     the 8 CTAs of a thread-block cluster and reduce through distributed shared memory, then the CTA that owns a row
//     adds bias + residual and writes h (fp32), bf16(h) and that row's partial LayerNorm statistics;
//   * no LayerNorm kernel: LN(h) W^T + b = rstd * (h W'^T - mean * colsum(W')) + b'  with W' = gamma (.) W and
//     b' = b + W beta, both folded once at pack time (packing.py), so the consumer product runs on bf16(h) directly and
//     applies (mean, rstd) in its epilogue;
//   * no logits round trip for greedy decoding: the lm_head CTAs keep a running (max, index) per row and write one
//     candidate per CTA; selection reads n_ctas candidates per row instead of 50,257 logits.
// Weights are read with 128-bit streaming loads (registers for the full-K products, bulk copies to shared memory for the
// cluster products) BEFORE griddepcontrol.wait, and every kernel asks for the NEXT kernel's weights to be brought into L2
// (cp.async.bulk.prefetch.L2), so the HBM stream runs one kernel ahead of the dependency chain.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace vc {

int make_tmap_bf16_kmajor(CUtensorMap* tm, const void* ptr, int rows, int K, int box_rows);
int make_tmap_bf16_kmajor_ld(CUtensorMap* tm, const void* ptr, int rows, int K, long long ld, int box_rows);

#define VC_LAUNCH(name, work, stream, ...)        \
  do {                                            \
    vc::KernelScope _ks(name, work, stream);      \
    __VA_ARGS__;                                  \
  } while (0)

namespace {

constexpr int DC_THREADS = 256;          // 8 warps
constexpr int DC_WARPS = 8;
constexpr int DC_MAX_PARTS = 16;         // partial LayerNorm statistics per row (one per cluster of the N = H products)
constexpr int DC_CLUSTER = 8;            // CTAs per cluster in the N = H products (portable maximum)
constexpr int DC_NCLUSTERS = 16;         // clusters per N = H product: 8 GPCs x 2 clusters of 8 SMs
constexpr int DC_MAX_ROWS = 1024;         // what the kernels handle (row tiles loop inside the CTAs)
constexpr int DC_BEST_ROWS = 64;          // where this chain beats the split-K chain (measured: 333 vs 463 us at 64 rows, 1028 vs 759 us at 256)

// Optional in-kernel timeline (tools/trace_decode.py): when a buffer is installed, thread 0 of the first and last CTA of every
// kernel of this file records %globaltimer at fixed points.  buf[0] = record counter, records of 8 x u64 from buf[8].
__device__ unsigned long long* dc_trace = nullptr;
__device__ int dc_trace_max = 0;
struct Trace {
  unsigned long long t[8];
  bool on;
  __device__ __forceinline__ Trace(int kernel_id) {
    on = dc_trace != nullptr && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && blockIdx.y == 0;
    if (on) {
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = 0;
      t[0] = (static_cast<unsigned long long>(blockIdx.x) << 8) | static_cast<unsigned>(kernel_id);
      mark(1);
    }
  }
  __device__ __forceinline__ void mark(int i) {
    if (on) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t[i]));
  }
  __device__ __forceinline__ void flush() {
    if (!on) return;
    mark(7);
    const unsigned long long idx = atomicAdd(dc_trace, 1ULL);
    if (idx < static_cast<unsigned long long>(dc_trace_max))
      for (int i = 0; i < 8; ++i) dc_trace[8 + idx * 8 + i] = t[i];
  }
};

__device__ __forceinline__ uint4 ldg_nc(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void l2_prefetch(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// This CTA's share of the next kernel's weight matrix -> L2 (16-byte granules, chunks of <= 32 KB, one thread issues).
__device__ __forceinline__ void prefetch_share(const void* next_w, long long next_bytes, int cta, int n_ctas) {
  if (next_w == nullptr || next_bytes <= 0) return;
  const long long per = ((next_bytes / n_ctas) + 15) & ~15LL;
  long long lo = per * cta, hi = lo + per;
  if (hi > next_bytes) hi = next_bytes & ~15LL;
  const uint8_t* p = static_cast<const uint8_t*>(next_w);
  for (; lo < hi; lo += 32768) l2_prefetch(p + lo, static_cast<uint32_t>(hi - lo < 32768 ? hi - lo : 32768));
}
__device__ __forceinline__ float gelu_new(float x) { return gelu_tanh(x); }

// LayerNorm statistics of one row from its partial (sum, M2) pairs over equal column groups (Chan's parallel variance).
__device__ __forceinline__ float2 row_mean_rstd(const float2* __restrict__ stat, int n_part, long long plane, long long row, int dim, float eps) {
  float2 v[DC_MAX_PARTS];
#pragma unroll
  for (int p = 0; p < DC_MAX_PARTS; ++p)
    if (p < n_part) v[p] = stat[p * plane + row];
  float s = 0.f;
#pragma unroll
  for (int p = 0; p < DC_MAX_PARTS; ++p)
    if (p < n_part) s += v[p].x;
  const float cols = static_cast<float>(dim / n_part);
  const float mean = s / static_cast<float>(dim);
  float m2 = 0.f;
#pragma unroll
  for (int p = 0; p < DC_MAX_PARTS; ++p)
    if (p < n_part) {
      const float d = v[p].x / cols - mean;
      m2 += v[p].y + cols * d * d;
    }
  return make_float2(mean, rsqrtf(m2 / static_cast<float>(dim) + eps));
}

// same, from partials staged in shared memory as [part][stride] (cp.async copies that travel with the activations)
__device__ __forceinline__ float2 row_mean_rstd_smem(const float2* sp, int n_part, int stride, int r, int dim, float eps) {
  float2 v[DC_MAX_PARTS];
#pragma unroll
  for (int p = 0; p < DC_MAX_PARTS; ++p) v[p] = p < n_part ? sp[p * stride + r] : make_float2(0.f, 0.f);   // all loads in flight together
  float s = 0.f;
#pragma unroll
  for (int p = 0; p < DC_MAX_PARTS; ++p) s += v[p].x;
  const float cols = static_cast<float>(dim / n_part), inv_cols = 1.0f / cols;
  const float mean = s / static_cast<float>(dim);
  float m2 = 0.f;
#pragma unroll
  for (int p = 0; p < DC_MAX_PARTS; ++p) {
    const float d = v[p].x * inv_cols - mean;
    m2 += p < n_part ? v[p].y + cols * d * d : 0.f;
  }
  return make_float2(mean, rsqrtf(m2 / static_cast<float>(dim) + eps));
}
__device__ __forceinline__ void cp_async8(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

// ------------------------------------------------------------------------------------------------ first rows of a forward
// h = embeds + wpe[past_len + l]; hb = bf16(h); stat[0][row] = (sum, M2) of the row.  One warp per row.
__global__ void __launch_bounds__(256) dc_add_pos_stats_kernel(const float* __restrict__ e, const float* __restrict__ wpe, float* __restrict__ h,
                                                               __nv_bfloat16* __restrict__ hb, float2* __restrict__ stat, int rows, int L,
                                                               int past_len, int dim, const void* next_w, long long next_bytes) {
  Trace tr(1);
  pdl_launch_dependents();       // at entry (see dc_fullk_kernel)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int nv = dim >> 7;
  float4 pe[8];
  const int l = row < rows ? row % L : 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) pe[i] = __ldg(reinterpret_cast<const float4*>(wpe + static_cast<long long>(past_len + l) * dim) + lane + 32 * i);
  if (threadIdx.x == 0) prefetch_share(next_w, next_bytes, blockIdx.x, gridDim.x);
  pdl_wait();
  tr.mark(2);
  if (row >= rows) { tr.flush(); return; }
  float4 v[8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      const float4 a = reinterpret_cast<const float4*>(e + static_cast<long long>(row) * dim)[lane + 32 * i];
      v[i] = make_float4(a.x + pe[i].x, a.y + pe[i].y, a.z + pe[i].z, a.w + pe[i].w);
      reinterpret_cast<float4*>(h + static_cast<long long>(row) * dim)[lane + 32 * i] = v[i];
      uint2 w;
      w.x = pack_bf16(v[i].x, v[i].y);
      w.y = pack_bf16(v[i].z, v[i].w);
      reinterpret_cast<uint2*>(hb + static_cast<long long>(row) * dim)[lane + 32 * i] = w;
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  sum = warp_sum(sum);
  const float mean = sum / static_cast<float>(dim);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  sq = warp_sum(sq);
  if (lane == 0) stat[row] = make_float2(sum, sq);
  tr.flush();
}

// ------------------------------------------------------------------------------------------------ full-K products (QKV, fc1)
// out[m][n] = epi( rstd_m * (sum_k hb[m][k] W'[n][k] - mean_m * cs[n]) + bf[n] )        M rows, N features, K = 256 * KB
// grid.x = N / (8 NF): a CTA owns 8 NF features for ALL rows.  Warp w owns the K range [w K/8, (w+1) K/8): its weights
// (NF x KB 16-byte loads per lane) sit in registers from before the dependency wait; activations arrive 32 rows at a time
// by bulk copy (two buffers), are the A operand of mma.sync m16n8k16 (weights: B operand, 8 features per tile, so feature
// tiles are 8 wide and every product of GPT-2 small / medium divides into <= 148 equal CTAs).  Both operands use the same
// lane-permuted K order inside a 32-wide block (one 16-byte load per lane covers slots {2t,2t+1,2t+8,2t+9} of two MMAs).
__host__ __device__ constexpr int red_pitch(int nf) { return (nf * 8) % 32 == 8 || (nf * 8) % 32 == 24 ? nf * 8 : nf * 8 + 8; }   // conflict-free float2 stores

// Register caps: a full-K CTA (<= 144 registers x 256 threads) and a cluster CTA (<= 112) fit one SM's register file together,
// and 135 KB + 65 KB of shared memory fit too, so each kernel's CTAs become resident and fetch their weights while the
// previous kernel is still running (without this the sixteenth cluster of fc2 waited for a second wave: 4 us per layer).
template <int NF, int KB, int EPI>
__global__ void __launch_bounds__(DC_THREADS, 1) __maxnreg__(KB <= 3 ? 144 : 192)
dc_fullk_kernel(const __nv_bfloat16* __restrict__ hb, const float2* __restrict__ stat, int n_part, const __nv_bfloat16* __restrict__ W,
                const float* __restrict__ cs, const float* __restrict__ bf, __nv_bfloat16* __restrict__ out, int M, int N, float eps,
                const void* next_w, long long next_bytes) {
  constexpr int K = DC_THREADS * KB;                 // 8 warps x KB blocks x 32
  constexpr int PITCH = K * 2 + 64;                  // bytes; == 64 (mod 128): 16-byte fragment loads of 8 rows never collide
  constexpr int RP = red_pitch(NF);
  constexpr int XBUF = 32 * PITCH;
  extern __shared__ __align__(128) uint8_t dc_smem[];
  uint8_t* xs = dc_smem;                                                       // [2][32][PITCH]
  float* red = reinterpret_cast<float*>(dc_smem + 2 * XBUF);                   // [8][32][RP]
  float2* s_part = reinterpret_cast<float2*>(red + DC_WARPS * 32 * RP);        // [2][16 parts][32 rows] partial row statistics
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * NF * 8;
  Trace tr(EPI ? 4 : 2);
  // Dependents are released at ENTRY: launching ~140 CTAs with >100 KB of shared memory takes the hardware 2-4 us, which is
  // only hidden if it starts this early.  Consequence for every kernel of the chain: before its own griddepcontrol.wait a
  // kernel may read constants only (weights, biases, tables) — its predecessor's predecessor may still be running.
  pdl_launch_dependents();

  uint4 w[NF][KB];
#pragma unroll
  for (int nf = 0; nf < NF; ++nf)
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
      w[nf][kb] = ldg_nc(W + static_cast<size_t>(n0 + nf * 8 + g) * K + (warp * KB + kb) * 32 + 8 * t);
  const int e_row = tid >> 3, e_f = tid & 7;          // epilogue: 32 rows x 8 feature lanes, NF features each
  float cs_r[NF], bf_r[NF];
#pragma unroll
  for (int j = 0; j < NF; ++j) {
    cs_r[j] = __ldg(cs + n0 + e_f + 8 * j);
    bf_r[j] = __ldg(bf + n0 + e_f + 8 * j);
  }
  if (tid == 0) prefetch_share(next_w, next_bytes, blockIdx.x, gridDim.x);
  pdl_wait();
  tr.mark(2);

  const int n_chunks = (M + 31) >> 5;
  // 32 rows x K*2 bytes per chunk as 16-byte asynchronous copies spread over all threads (one commit group per chunk).
  // Measured: 64 per-row bulk copies took 3.5 us to land 96 KB; the same bytes as cp.async arrive in about one L2 round trip
  // plus the SM's fill time.
  auto issue = [&](int c) {
    constexpr int CPR = K * 2 / 16;                    // 16-byte pieces per row
    const uint32_t dst0 = smem_u32(xs + (c & 1) * XBUF);
    for (int i = tid; i < 32 * CPR; i += DC_THREADS) {
      const int r = i / CPR, col = i - r * CPR;
      int row = c * 32 + r;
      row = row < M ? row : M - 1;
      cp_async16(dst0 + r * PITCH + col * 16, reinterpret_cast<const uint8_t*>(hb + static_cast<size_t>(row) * K) + col * 16);
    }
    // the rows' partial LayerNorm statistics ride in the same group: consumed in the epilogue, never waited for on their own
    // (computing mean / rstd from global memory up front held the whole CTA back by ~2 us: measured with tools/trace_decode.py)
    for (int i = tid; i < n_part * 32; i += DC_THREADS) {
      const int p = i >> 5, r = i & 31;
      int row = c * 32 + r;
      row = row < M ? row : M - 1;
      cp_async8(smem_u32(s_part + ((c & 1) * DC_MAX_PARTS + p) * 32 + r), stat + static_cast<size_t>(p) * M + row);
    }
    cp_async_commit();
  };
  issue(0);
  if (n_chunks > 1) issue(1); else cp_async_commit();

  const uint32_t kbyte = static_cast<uint32_t>((warp * KB * 32 + 8 * t) * 2);
  for (int c = 0; c < n_chunks; ++c) {
    cp_async_wait<1>();                                // this thread's pieces of chunk c (the newest group may still fly)
    __syncthreads();                                   // ... and everybody else's
    if (c == 0) tr.mark(3);
    // this thread's epilogue row: (mean, rstd) from the partials that arrived with the chunk, computed under the MMAs
    const float2 mr = row_mean_rstd_smem(s_part + (c & 1) * DC_MAX_PARTS * 32, n_part, 32, e_row, K, eps);
    const uint32_t xb0 = smem_u32(xs + (c & 1) * XBUF) + kbyte;
    float acc[2][NF][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) acc[mt][nf][0] = acc[mt][nf][1] = acc[mt][nf][2] = acc[mt][nf][3] = 0.f;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      uint4 xa[2], xb[2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        xa[mt] = lds128(xb0 + (mt * 16 + g) * PITCH + kb * 64);
        xb[mt] = lds128(xb0 + (mt * 16 + g + 8) * PITCH + kb * 64);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) mma16816(acc[mt][nf], xa[mt].x, xb[mt].x, xa[mt].y, xb[mt].y, w[nf][kb].x, w[nf][kb].y);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) mma16816(acc[mt][nf], xa[mt].z, xb[mt].z, xa[mt].w, xb[mt].w, w[nf][kb].z, w[nf][kb].w);
    }
    float* rw = red + warp * 32 * RP;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) {
        *reinterpret_cast<float2*>(rw + (mt * 16 + g) * RP + nf * 8 + 2 * t) = make_float2(acc[mt][nf][0], acc[mt][nf][1]);
        *reinterpret_cast<float2*>(rw + (mt * 16 + g + 8) * RP + nf * 8 + 2 * t) = make_float2(acc[mt][nf][2], acc[mt][nf][3]);
      }
    if (c == 0) tr.mark(4);
    __syncthreads();                                   // partials complete; every warp is done with this activation buffer
    if (c == 0) tr.mark(5);
    const int row = c * 32 + e_row;
    if (row < M) {
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const float* rr = red + e_row * RP + e_f + 8 * j;
        float s = rr[0];
#pragma unroll
        for (int ww = 1; ww < DC_WARPS; ++ww) s += rr[ww * 32 * RP];         // fixed order
        float y = mr.y * (s - mr.x * cs_r[j]) + bf_r[j];
        if (EPI == 1) y = gelu_new(y);
        out[static_cast<size_t>(row) * N + n0 + e_f + 8 * j] = __float2bfloat16_rn(y);
      }
    }
    __syncthreads();                                   // the partial buffer is reused by the next chunk
    if (c + 2 < n_chunks) issue(c + 2); else cp_async_commit();       // always one group per iteration: wait_group<1> stays exact
  }
  tr.flush();
}

// ------------------------------------------------------------------------------------------------ N = H products (proj, fc2)
// h[m][n] += bias[n] + sum_k x[m][k] W[n][k];  hb = bf16(h);  stat[cluster][m] = (sum, M2) over the cluster's features.
// 16 clusters x 8 CTAs.  Cluster c owns features [c FT, (c+1) FT), FT = H/16 = 8 NT; CTA r of a cluster owns the K slice
// [r K/8, (r+1) K/8).  Warp w = (feature group w / KQ, K quarter w % KQ) keeps ITS weights (NPW x KBW 16-byte loads per lane)
// in registers from before the dependency wait and runs all 64 rows of a row tile against them; the activation slice
// arrives by cp.async after the wait.  With KQ > 1 the K quarters are first summed through shared memory (the buffer the
// activations occupied).  Every CTA then sends its [64 x FT] partial straight into the shared memory of the CTAs that own the
// rows (rows 8 q .. 8 q + 7 of a row tile belong to CTA q); after one cluster barrier each owner adds the 8 partials in a
// fixed order, then bias and residual, and writes h, bf16(h) and the row statistics.
template <int NT, int NSPLIT, int KQ, int KBW>
__global__ void __cluster_dims__(DC_CLUSTER, 1, 1) __launch_bounds__(DC_THREADS, 1) __maxnreg__(KBW <= 3 ? 112 : 224)
dc_nk_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ W, const float* __restrict__ bias, float* __restrict__ h,
             __nv_bfloat16* __restrict__ hb, float2* __restrict__ stat, int M, int K, const void* next_w, long long next_bytes) {
  constexpr int FT = NT * 8;
  constexpr int H = DC_NCLUSTERS * FT;
  constexpr int NPW = NT / NSPLIT;                     // n8 tiles per warp
  constexpr int KSLICE = KQ * KBW * 32;
  constexpr int PITCH = (KSLICE * 2) % 128 == 64 ? KSLICE * 2 : KSLICE * 2 + 64;
  constexpr int LP = FT + 8;                           // fp32 pitch of the local K-quarter partials: == 8 or 24 (mod 32), conflict-free float2
  constexpr int XS_BYTES = 64 * PITCH > KQ * 64 * LP * 4 ? 64 * PITCH : KQ * 64 * LP * 4;
  static_assert(NSPLIT * KQ <= DC_WARPS && NT % NSPLIT == 0, "warp layout");
  extern __shared__ __align__(128) uint8_t dc_smem[];
  uint8_t* xs = dc_smem;                                                       // [64][PITCH]; reused as [KQ][64][FT] fp32
  float* lred = reinterpret_cast<float*>(dc_smem);
  float* recv = reinterpret_cast<float*>(dc_smem + XS_BYTES);                  // [2][8 src][8 rows][FT]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / DC_CLUSTER;
  const int n0 = cid * FT;
  const int k0 = static_cast<int>(rank) * KSLICE;
  const int n_tiles = (M + 63) >> 6;
  Trace tr(K == H ? 3 : 5);
  pdl_launch_dependents();       // at entry (see dc_fullk_kernel)
  const bool mma_warp = warp < NSPLIT * KQ;
  const int ng = warp / KQ, kq = warp % KQ;

  uint4 w[NPW][KBW];
  if (mma_warp) {
#pragma unroll
    for (int j = 0; j < NPW; ++j)
#pragma unroll
      for (int kb = 0; kb < KBW; ++kb)
        w[j][kb] = ldg_nc(W + static_cast<size_t>(n0 + (ng * NPW + j) * 8 + g) * K + k0 + (kq * KBW + kb) * 32 + 8 * t);
  }
  if (tid == 0) prefetch_share(next_w, next_bytes, blockIdx.x, gridDim.x);
  // owner role: warp = local row, lanes = features lane, lane + 32 (< FT)
  const bool c1 = lane + 32 < FT;
  const float b0 = __ldg(bias + n0 + lane), b1 = c1 ? __ldg(bias + n0 + lane + 32) : 0.f;
  auto own_row = [&](int tile) { return tile * 64 + 8 * static_cast<int>(rank) + warp; };
  cluster_sync_all();                                  // every CTA of the cluster is running before any remote store
  pdl_wait();
  tr.mark(2);
  // residual of the row this warp owns: requested now, consumed after the reduction
  float h0 = 0.f, h1 = 0.f;
  {
    const int row = own_row(0);
    if (row < M) {
      h0 = h[static_cast<size_t>(row) * H + n0 + lane];
      if (c1) h1 = h[static_cast<size_t>(row) * H + n0 + lane + 32];
    }
  }

  auto issue_x = [&](int tile) {                       // 64 rows x KSLICE*2 bytes as 16-byte asynchronous copies
    constexpr int CPR = KSLICE * 2 / 16;
    const uint32_t dst0 = smem_u32(xs);
    for (int i = tid; i < 64 * CPR; i += DC_THREADS) {
      const int r = i / CPR, col = i - r * CPR;
      int row = tile * 64 + r;
      row = row < M ? row : M - 1;
      cp_async16(dst0 + r * PITCH + col * 16, reinterpret_cast<const uint8_t*>(x + static_cast<size_t>(row) * K + k0) + col * 16);
    }
    cp_async_commit();
  };
  issue_x(0);
  const uint32_t xa_addr = smem_u32(xs) + g * PITCH + static_cast<uint32_t>((kq * KBW * 32 + 8 * t) * 2);
  for (int tile = 0; tile < n_tiles; ++tile) {
    cp_async_wait<0>();
    __syncthreads();
    if (tile == 0) tr.mark(3);
    float acc[4][NPW][4];
    if (mma_warp) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int j = 0; j < NPW; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;
#pragma unroll
      for (int kb = 0; kb < KBW; ++kb) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const uint4 xa = lds128(xa_addr + mt * 16 * PITCH + kb * 64), xb = lds128(xa_addr + (mt * 16 + 8) * PITCH + kb * 64);
#pragma unroll
          for (int j = 0; j < NPW; ++j) mma16816(acc[mt][j], xa.x, xb.x, xa.y, xb.y, w[j][kb].x, w[j][kb].y);
#pragma unroll
          for (int j = 0; j < NPW; ++j) mma16816(acc[mt][j], xa.z, xb.z, xa.w, xb.w, w[j][kb].z, w[j][kb].w);
        }
      }
    }
    float* rbase = recv + (tile & 1) * DC_CLUSTER * 8 * FT + static_cast<int>(rank) * 8 * FT;     // [8 rows][FT] slot of this source
    if (KQ == 1) {
      // a warp holds complete K-slice sums for its features: rows 16 mt + g -> CTA 2 mt, rows 16 mt + 8 + g -> CTA 2 mt + 1
      if (mma_warp) {
        const uint32_t slot = smem_u32(rbase + g * FT + ng * NPW * 8 + 2 * t);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const uint32_t dst_lo = mapa_shared(slot, 2 * mt), dst_hi = mapa_shared(slot, 2 * mt + 1);
#pragma unroll
          for (int j = 0; j < NPW; ++j) {
            asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(dst_lo + j * 32), "f"(acc[mt][j][0]), "f"(acc[mt][j][1]) : "memory");
            asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(dst_hi + j * 32), "f"(acc[mt][j][2]), "f"(acc[mt][j][3]) : "memory");
          }
        }
      }
    } else {
      __syncthreads();                                 // every warp is done reading the activation slice
      if (mma_warp) {
        float* lw = lred + kq * 64 * LP + ng * NPW * 8 + 2 * t;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int j = 0; j < NPW; ++j) {
            *reinterpret_cast<float2*>(lw + (mt * 16 + g) * LP + j * 8) = make_float2(acc[mt][j][0], acc[mt][j][1]);
            *reinterpret_cast<float2*>(lw + (mt * 16 + g + 8) * LP + j * 8) = make_float2(acc[mt][j][2], acc[mt][j][3]);
          }
      }
      __syncthreads();
      for (int i = tid; i < 64 * FT / 2; i += DC_THREADS) {
        const int row = i / (FT / 2), c2 = (i - row * (FT / 2)) * 2;
        float2 a = *reinterpret_cast<const float2*>(lred + row * LP + c2);
#pragma unroll
        for (int q = 1; q < KQ; ++q) {                 // fixed order
          const float2 b = *reinterpret_cast<const float2*>(lred + (q * 64 + row) * LP + c2);
          a.x += b.x; a.y += b.y;
        }
        const uint32_t dst = mapa_shared(smem_u32(rbase + (row & 7) * FT + c2), row >> 3);
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(dst), "f"(a.x), "f"(a.y) : "memory");
      }
    }
    if (tile == 0) tr.mark(4);
    cluster_sync_all();                                // partials of all 8 K slices have landed; xs is free in every CTA
    if (tile == 0) tr.mark(5);
    if (tile + 1 < n_tiles) issue_x(tile + 1);
    const int row = own_row(tile);
    const float* rr = recv + ((tile & 1) * DC_CLUSTER * 8 + warp) * FT + lane;
    float v0 = rr[0], v1 = c1 ? rr[32] : 0.f;
#pragma unroll
    for (int s2 = 1; s2 < DC_CLUSTER; ++s2) {          // fixed order
      v0 += rr[s2 * 8 * FT];
      if (c1) v1 += rr[s2 * 8 * FT + 32];
    }
    v0 += b0 + h0;
    v1 = c1 ? v1 + b1 + h1 : 0.f;
    if (row < M) {
      h[static_cast<size_t>(row) * H + n0 + lane] = v0;
      hb[static_cast<size_t>(row) * H + n0 + lane] = __float2bfloat16_rn(v0);
      if (c1) {
        h[static_cast<size_t>(row) * H + n0 + lane + 32] = v1;
        hb[static_cast<size_t>(row) * H + n0 + lane + 32] = __float2bfloat16_rn(v1);
      }
    }
    const float sum = warp_sum(v0 + v1);
    const float mloc = sum / static_cast<float>(FT);
    const float d0 = v0 - mloc, d1 = c1 ? v1 - mloc : 0.f;
    const float m2 = warp_sum(d0 * d0 + d1 * d1);
    if (lane == 0 && row < M) stat[static_cast<size_t>(cid) * M + row] = make_float2(sum, m2);
    if (tile + 1 < n_tiles) {                          // residual of the next tile's row
      const int nrow = own_row(tile + 1);
      h0 = h1 = 0.f;
      if (nrow < M) {
        h0 = h[static_cast<size_t>(nrow) * H + n0 + lane];
        if (c1) h1 = h[static_cast<size_t>(nrow) * H + n0 + lane + 32];
      }
    }
  }
  tr.flush();
}

// ------------------------------------------------------------------------------------------------ lm_head + argmax candidates
// logits[r][n] = rstd_r * (sum_k hb[row(r)][k] Wf[n][k] - mean_r cs[n]) + bf[n],  row(r) = (m0 + r) * row_stride + row_offset
// (last position of every sequence).  grid = (G, row tiles of 64).  A CTA owns a contiguous range of 16-feature tiles; every WARP
// takes whole tiles (w, w + 8, ... of the range) and runs all 64 rows and the full K of a tile by itself, streaming the tile's
// weights through registers in K chunks of 192 (double-buffered, the next chunk — of this or of the warp's next tile — always in
// flight), so there is no cross-warp reduction and no __syncthreads after the activations have landed.  (The first version split K
// over the 8 warps like dc_fullk_kernel: its shared-memory reduction per tile cost as much as the MMAs — 33 us per step.)  Each
// lane keeps a running (max, lowest index) for the 8 rows it sees; one candidate per row and CTA leaves at the end; logits are
// stored only when a buffer is given (teacher-forced tests, beam search).
template <int KB>   // K = 256 * KB = 8 chunks... K / 192 or K / 256 chunks of CK k32-blocks
__global__ void __launch_bounds__(DC_THREADS, 1)
dc_lmhead_kernel(const __nv_bfloat16* __restrict__ hb, const float2* __restrict__ stat, int n_part, long long stat_plane, long long row_stride,
                 long long row_offset, const __nv_bfloat16* __restrict__ W, const float* __restrict__ cs, const float* __restrict__ bf, int vocab,
                 int vocab_pad, int n_rows, float eps, float* __restrict__ logits, long long ld, float* __restrict__ cand_v,
                 int* __restrict__ cand_i) {
  constexpr int K = DC_THREADS * KB;
  constexpr int PITCH = K * 2 + 64;
  constexpr int CK = KB == 3 ? 6 : 8;                  // k32 blocks per register chunk: 192 (K = 768) or 256 (K = 1024) columns
  constexpr int NCH = (K / 32) / CK;                   // chunks per tile: 4
  extern __shared__ __align__(128) uint8_t dc_smem[];
  uint8_t* xs = dc_smem;                                                       // [64][PITCH]
  float2* s_mr = reinterpret_cast<float2*>(dc_smem + 64 * PITCH);              // [64] (mean, rstd)
  float2* s_part = s_mr + 64;                                                  // [16 parts][64 rows]
  float* s_bv = reinterpret_cast<float*>(s_part + DC_MAX_PARTS * 64);          // [8 warps][64 rows] best value
  int* s_bi = reinterpret_cast<int*>(s_bv + DC_WARPS * 64);                    // [8 warps][64 rows] best index
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.y * 64;
  const int tiles = vocab_pad >> 4;
  const int t_begin = static_cast<int>(static_cast<long long>(tiles) * blockIdx.x / gridDim.x);
  const int t_end = static_cast<int>(static_cast<long long>(tiles) * (blockIdx.x + 1) / gridDim.x);
  Trace tr(6);
  pdl_launch_dependents();       // at entry (see dc_fullk_kernel)

  // this warp's chunk stream: tile t_begin + warp + 8 i, chunk c of it
  uint4 wA[2][CK], wB[2][CK];
  auto load_w = [&](uint4 (&w)[2][CK], int tile, int ch) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int kb = 0; kb < CK; ++kb)
        w[nt][kb] = ldg_nc(W + static_cast<size_t>(tile * 16 + nt * 8 + g) * K + (ch * CK + kb) * 32 + 8 * t);
  };
  int tile = t_begin + warp;
  if (tile < t_end) {
    load_w(wA, tile, 0);
    load_w(wB, tile, 1);
  }
  pdl_wait();
  tr.mark(2);
  {
    constexpr int CPR = K * 2 / 16;
    const uint32_t dst0 = smem_u32(xs);
    for (int i = tid; i < 64 * CPR; i += DC_THREADS) {
      const int rr = i / CPR, col = i - rr * CPR;
      int r = m0 + rr;
      r = r < n_rows ? r : n_rows - 1;
      cp_async16(dst0 + rr * PITCH + col * 16, reinterpret_cast<const uint8_t*>(hb + static_cast<size_t>(r * row_stride + row_offset) * K) + col * 16);
    }
    for (int i = tid; i < n_part * 64; i += DC_THREADS) {
      const int p = i >> 6, rr = i & 63;
      int r = m0 + rr;
      r = r < n_rows ? r : n_rows - 1;
      cp_async8(smem_u32(s_part + p * 64 + rr), stat + static_cast<size_t>(p) * stat_plane + r * row_stride + row_offset);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  tr.mark(3);
  __syncthreads();                                     // activations and partial statistics visible
  if (tid < 64) s_mr[tid] = row_mean_rstd_smem(s_part, n_part, 64, tid, K, eps);
  __syncthreads();
  // rows of this lane: mt * 16 + g and + 8  ->  index 2 mt, 2 mt + 1
  float mean[8], rstd[8], best[8];
  int best_i[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float2 mr = s_mr[(q >> 1) * 16 + g + (q & 1) * 8];
    mean[q] = mr.x; rstd[q] = mr.y;
    best[q] = -INFINITY; best_i[q] = 0x7fffffff;
  }
  const uint32_t xrow = smem_u32(xs) + g * PITCH + t * 16;

  float acc[4][2][4];
  auto mma_chunk = [&](const uint4 (&w)[2][CK], int ch) {
#pragma unroll
    for (int kb = 0; kb < CK; ++kb) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const uint4 xa = lds128(xrow + mt * 16 * PITCH + (ch * CK + kb) * 64), xb = lds128(xrow + (mt * 16 + 8) * PITCH + (ch * CK + kb) * 64);
        mma16816(acc[mt][0], xa.x, xb.x, xa.y, xb.y, w[0][kb].x, w[0][kb].y);
        mma16816(acc[mt][1], xa.x, xb.x, xa.y, xb.y, w[1][kb].x, w[1][kb].y);
        mma16816(acc[mt][0], xa.z, xb.z, xa.w, xb.w, w[0][kb].z, w[0][kb].w);
        mma16816(acc[mt][1], xa.z, xb.z, xa.w, xb.w, w[1][kb].z, w[1][kb].w);
      }
    }
  };
  for (; tile < t_end; tile += DC_WARPS) {
    const int f0 = tile * 16;
    // epilogue constants of this tile (columns nt * 8 + 2 t, + 1): requested before the MMAs
    float2 c_s[2], c_b[2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      c_s[nt] = __ldg(reinterpret_cast<const float2*>(cs + f0 + nt * 8 + 2 * t));
      c_b[nt] = __ldg(reinterpret_cast<const float2*>(bf + f0 + nt * 8 + 2 * t));
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
    const int next = tile + DC_WARPS;
#pragma unroll
    for (int ch = 0; ch < NCH; ch += 2) {
      mma_chunk(wA, ch);
      if (ch + 2 < NCH) load_w(wA, tile, ch + 2);
      else if (next < t_end) load_w(wA, next, 0);
      mma_chunk(wB, ch + 1);
      if (ch + 3 < NCH) load_w(wB, tile, ch + 3);
      else if (next < t_end) load_w(wB, next, 1);
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = f0 + nt * 8 + 2 * t;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {               // rows mt * 16 + g (h2 = 0) and + 8 (h2 = 1)
          const int q = 2 * mt + h2;
          const float y0 = rstd[q] * (acc[mt][nt][2 * h2] - mean[q] * c_s[nt].x) + c_b[nt].x;
          const float y1 = rstd[q] * (acc[mt][nt][2 * h2 + 1] - mean[q] * c_s[nt].y) + c_b[nt].y;
          if (col < vocab && y0 > best[q]) { best[q] = y0; best_i[q] = col; }            // increasing index, strict '>'
          if (col + 1 < vocab && y1 > best[q]) { best[q] = y1; best_i[q] = col + 1; }
          const int r = m0 + mt * 16 + g + 8 * h2;
          if (logits != nullptr && r < n_rows) *reinterpret_cast<float2*>(logits + static_cast<size_t>(r) * ld + col) = make_float2(y0, y1);
        }
      }
  }
  // lanes t = 0..3 of a row group hold different columns of the same rows; then the 8 warps; ties -> lowest index
#pragma unroll
  for (int q = 0; q < 8; ++q) {
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best[q], o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i[q], o);
      if (ov > best[q] || (ov == best[q] && oi < best_i[q])) { best[q] = ov; best_i[q] = oi; }
    }
    if (t == 0) {
      const int row = (q >> 1) * 16 + g + (q & 1) * 8;
      s_bv[warp * 64 + row] = best[q];
      s_bi[warp * 64 + row] = best_i[q];
    }
  }
  __syncthreads();
  if (tid < 64 && m0 + tid < n_rows) {
    float bv = s_bv[tid];
    int bi = s_bi[tid];
#pragma unroll
    for (int w2 = 1; w2 < DC_WARPS; ++w2) {
      const float ov = s_bv[w2 * 64 + tid];
      const int oi = s_bi[w2 * 64 + tid];
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    cand_v[static_cast<size_t>(m0 + tid) * gridDim.x + blockIdx.x] = bv;
    cand_i[static_cast<size_t>(m0 + tid) * gridDim.x + blockIdx.x] = bi;
  }
  tr.flush();
}

// ------------------------------------------------------------------------------------------------ lm_head on tcgen05
// Same product and the same outputs as dc_lmhead_kernel, for up to 64 rows per row tile.  ncu on the mma.sync kernel: 8 warps of
// dependent LDS -> HMMA chains keep the tensor pipe 26 % busy and issue on 22 % of the cycles — neither HBM nor L2 is the limit
// (requesting the CTA's whole weight range into L2 ahead of time changes nothing): 23-27 us per step for 77 MB.  Here the
// activations sit in shared memory ONCE as the A operand (64 rows x K, K-major, 128-byte swizzle: K/64 TMA boxes of 8 KB) and the
// CTA's ~340 vocabulary rows stream through a 5-stage ring of [64 rows x 64 K] TMA boxes as the B operand of
// tcgen05.mma (M = 128, N = 64, K = 16; accumulators of the CTA's up to six 64-column sub-tiles side by side in TMEM).  M = 128
// with 64 real rows: the instruction reads 16 KB from each A box, i.e. it runs on into the next box (for the last one: into the B
// ring), and accumulator lanes 64-127 hold garbage nobody reads — the M = 128 accumulator layout (row = TMEM lane) is the one the
// rest of this code base uses, and a single-CTA M = 64 instruction takes the same time.  Four epilogue warps (two per lane
// quarter, one 32-column chunk of every sub-tile each) apply the folded ln_f + bias to one row per thread and keep the running
// (max, lowest index) while the next sub-tile's MMAs run; logits are stored only when a buffer is given.  Shared memory stays at
// 141 KB so that these CTAs become resident beside the last fc2's (80 KB) — and request their first weight boxes and the L2
// prefetch of the rest — while that kernel is still running.  Measured: 15-16 us from the dependency to the last candidate
// (mma.sync kernel: 23-27), step p50 323 -> 314 us at 64 sequences; the kernel now runs at what HBM delivers (77 MB at ~5 TB/s).
// Tried, none faster: a 12-stage ring without the co-residency (316 us); K blocks in the outer loop with A streamed through a
// 4-deep ring and 13 weight boxes in flight (325 us); keeping the matrix in L2 across steps — evict_last / evict_first cache
// hints on the weight loads, or the last layer's four kernels each prefetching a quarter of it — ncu (--cache-control none)
// still shows 78 MB of DRAM reads inside this kernel: 77 MB does not stay in the 126 MB (two-partition) L2 under a 247 MB step.
constexpr int LT_NT = 64;                      // vocabulary rows per sub-tile (UMMA N)
constexpr int LT_SUB = 6;                      // sub-tiles per CTA: up to 384 rows (148 CTAs: 340)
constexpr int LT_STAGES = 5;
constexpr int LT_B_BYTES = LT_NT * 128;        // 8 KB per stage
constexpr int LT_THREADS = 192;                // warps 0, 1, 4, 5: epilogue (lane quarters 0 and 1), warp 2: TMA, warp 3: MMA
__host__ __device__ constexpr int lt_smem(int K) { return (K / 64) * 8192 + LT_STAGES * LT_B_BYTES + 2 * LT_SUB * LT_NT * 4 + 2 * 2 * 64 * 4 + 256 + 1024; }

__global__ void __launch_bounds__(LT_THREADS, 1)
dc_lmhead_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, int K, const float2* __restrict__ stat,
                    int n_part, long long stat_plane, long long row_stride, long long row_offset, const __nv_bfloat16* __restrict__ W,
                    const float* __restrict__ cs, const float* __restrict__ bf, int vocab, int vocab_pad, int n_rows, float eps,
                    float* __restrict__ logits, long long ld, float* __restrict__ cand_v, int* __restrict__ cand_i) {
  extern __shared__ uint8_t lt_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lt_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int k_blocks = K >> 6;
  uint8_t* sA = smem;                                            // [k_blocks][64 rows][128 B]
  uint8_t* sB = smem + k_blocks * 8192;                          // [LT_STAGES][64 rows][128 B]
  float* s_cs = reinterpret_cast<float*>(sB + LT_STAGES * LT_B_BYTES);   // [384]
  float* s_bf = s_cs + LT_SUB * LT_NT;                           // [384]
  float* s_bv = s_bf + LT_SUB * LT_NT;                           // [2 halves][64 rows]
  int* s_bi = reinterpret_cast<int*>(s_bv + 2 * 64);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bi + 2 * 64);
  uint64_t* empty_bar = full_bar + LT_STAGES;
  uint64_t* a_full = empty_bar + LT_STAGES;
  uint64_t* tfull = a_full + 1;                                  // [LT_SUB]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + LT_SUB);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * 64;
  const int tiles = vocab_pad >> 4;
  const int t_begin = static_cast<int>(static_cast<long long>(tiles) * blockIdx.x / gridDim.x);
  const int t_end = static_cast<int>(static_cast<long long>(tiles) * (blockIdx.x + 1) / gridDim.x);
  const int v0 = t_begin * 16, v1 = t_end * 16;                  // this CTA's vocabulary rows (host: v1 - v0 <= LT_SUB * LT_NT)
  const int n_sub = (v1 - v0 + LT_NT - 1) / LT_NT;
  const int total = n_sub * k_blocks;                            // ring stages this CTA walks
  Trace tr(6);
  pdl_launch_dependents();       // at entry (see dc_fullk_kernel)

  if (tid == 0) {
    for (int i = 0; i < LT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(a_full, 1);
    for (int i = 0; i < LT_SUB; ++i) mbar_init(&tfull[i], 1);
    fence_mbar_init();
  }
  if (warp == 2 && lane == 0) { tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_w); }
  if (warp == 3) tmem_alloc(tmem_slot, 512);
  // folded ln_f constants of the CTA's columns (constants: before the dependency wait)
  for (int i = tid; i < LT_SUB * LT_NT; i += LT_THREADS) {
    const int col = v0 + i;
    s_cs[i] = col < vocab_pad ? __ldg(cs + col) : 0.f;
    s_bf[i] = col < vocab_pad ? __ldg(bf + col) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 2) {
    // ------------------------------------------------ TMA producer: weights are constants, so the ring is filled and the rest of
    // the CTA's weight range is requested into L2 before the dependency wait; the activations follow it
    int it = 0;
    for (; it < LT_STAGES && it < total; ++it) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[it], LT_B_BYTES);
        tma_load_2d(&tm_w, &full_bar[it], sB + it * LT_B_BYTES, (it % k_blocks) * 64, v0 + (it / k_blocks) * LT_NT);
      }
      __syncwarp();
    }
    {
      const uint8_t* wb = reinterpret_cast<const uint8_t*>(W) + static_cast<size_t>(v0) * K * 2;
      const long long bytes = static_cast<long long>(v1 - v0) * K * 2;
      for (long long off = static_cast<long long>(lane) * 16384; off < bytes; off += 32LL * 16384)
        l2_prefetch(wb + off, static_cast<uint32_t>(bytes - off < 16384 ? bytes - off : 16384));
    }
    pdl_wait();
    if (elect_one()) {
      mbar_arrive_expect_tx(a_full, static_cast<uint32_t>(k_blocks) * 8192);
      for (int kb = 0; kb < k_blocks; ++kb) tma_load_2d(&tm_a, a_full, sA + kb * 8192, kb * 64, m0);
    }
    __syncwarp();
    int stage = it % LT_STAGES;
    uint32_t phase = 0;                                          // parity of the ring pass whose slots are being refilled
    for (; it < total; ++it) {
      mbar_wait(&empty_bar[stage], phase);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], LT_B_BYTES);
        tma_load_2d(&tm_w, &full_bar[stage], sB + stage * LT_B_BYTES, (it % k_blocks) * 64, v0 + (it / k_blocks) * LT_NT);
      }
      __syncwarp();
      if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, LT_NT);
    constexpr uint64_t desc_hi = umma_desc_sw128_hi();
    const uint32_t a_lo = (smem_u32(sA) & 0x3FFFFu) >> 4, b_lo = (smem_u32(sB) & 0x3FFFFu) >> 4;
    mbar_wait(a_full, 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    for (int sub = 0; sub < n_sub; ++sub) {
      const uint32_t d_tmem = tmem_base + sub * LT_NT;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = desc_hi | (a_lo + kb * (8192 >> 4));
          const uint64_t db = desc_hi | (b_lo + stage * (LT_B_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          tc_commit(&empty_bar[stage]);
          if (kb == k_blocks - 1) tc_commit(&tfull[sub]);
        }
        __syncwarp();
        if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue: one row per thread
    pdl_wait();
    const int quarter = warp & 3;          // 0 / 1: accumulator lanes (= rows) 32 quarter .. + 31
    const int half = warp >> 2;            // which 32-column chunk of every 64-column sub-tile
    const int rr = quarter * 32 + lane, row = m0 + rr;
    const int rowc = row < n_rows ? row : n_rows - 1;
    const float2 mr = row_mean_rstd(stat, n_part, stat_plane, rowc * row_stride + row_offset, K, eps);
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    tr.mark(2);
    for (int sub = 0; sub < n_sub; ++sub) {
      mbar_wait(&tfull[sub], 0);
      tc_fence_after();
      if (sub == 0) tr.mark(3);
      const int ci = sub * LT_NT + half * 32;             // column index inside the CTA's range
      if (v0 + ci >= v1) continue;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ci, r);
      tmem_ld_wait();
      float y[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        y[j] = mr.y * (__uint_as_float(r[j]) - mr.x * s_cs[ci + j]) + s_bf[ci + j];
        const int col = v0 + ci + j;
        if (col < vocab && col < v1 && y[j] > best) { best = y[j]; best_i = col; }       // increasing index, strict '>'
      }
      if (logits != nullptr && row < n_rows) {
        float* o = logits + static_cast<size_t>(row) * ld + v0 + ci;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (v0 + ci + j < v1) *reinterpret_cast<float4*>(o + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
      }
    }
    tc_fence_before();
    s_bv[half * 64 + rr] = best;
    s_bi[half * 64 + rr] = best_i;
    asm volatile("bar.sync 1, 128;" ::: "memory");                 // the four epilogue warps
    if (half == 0 && row < n_rows) {
      const float ov = s_bv[64 + rr];
      const int oi = s_bi[64 + rr];
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
      cand_v[static_cast<size_t>(row) * gridDim.x + blockIdx.x] = best;
      cand_i[static_cast<size_t>(row) * gridDim.x + blockIdx.x] = best_i;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  tr.flush();
}

// ------------------------------------------------------------------------------------------------ selection
__device__ __forceinline__ int block_argmax_cands(const float* __restrict__ v, const int* __restrict__ idx, int n) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float a = v[j];
    const int i = idx[j];
    if (a > best || (a == best && i < bi)) { best = a; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { s_v[warp] = best; s_i[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < nw ? s_v[lane] : -INFINITY;
    bi = lane < nw ? s_i[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) s_i[0] = (bi == 0x7fffffff) ? 0 : bi;
  }
  __syncthreads();
  return s_i[0];
}

// argmax over the lm_head candidates (ties -> lowest index, torch.argmax) + the bookkeeping of benchmark_baseline.py:210-227
// + the NEXT step's input row: h = wte[token] + wpe[next_pos], hb = bf16(h), stat[0][row] — one CTA (dim/4 threads) per row.
__global__ void __launch_bounds__(256) dc_select_kernel(const float* __restrict__ cand_v, const int* __restrict__ cand_i, int n_cand, int step,
                                                        int max_new, int eos, int32_t* __restrict__ finished, int32_t* __restrict__ ids_out,
                                                        int32_t* __restrict__ len_out, const int32_t* __restrict__ forced,
                                                        const __nv_bfloat16* __restrict__ wte, const float* __restrict__ wpe, int next_pos, int dim,
                                                        float* __restrict__ h, __nv_bfloat16* __restrict__ hb, float2* __restrict__ stat,
                                                        int32_t* __restrict__ next_ids, const void* next_w, long long next_bytes) {
  __shared__ float s_red[8];
  Trace tr(7);
  pdl_launch_dependents();       // at entry (see dc_fullk_kernel)
  const int r = blockIdx.x, c4 = threadIdx.x;
  float4 pe = make_float4(0.f, 0.f, 0.f, 0.f);
  if (h != nullptr) pe = __ldg(reinterpret_cast<const float4*>(wpe + static_cast<long long>(next_pos) * dim) + c4);
  if (threadIdx.x == 0) prefetch_share(next_w, next_bytes, blockIdx.x, gridDim.x);
  pdl_wait();
  tr.mark(2);
  int tok = block_argmax_cands(cand_v + static_cast<size_t>(r) * n_cand, cand_i + static_cast<size_t>(r) * n_cand, n_cand);
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
  if (h == nullptr) { tr.flush(); return; }
  const long long feed = (forced != nullptr) ? forced[static_cast<long long>(r) * max_new + step] : tok;
  const uint2 e = *(reinterpret_cast<const uint2*>(wte + feed * dim) + c4);
  const float2 a = unpack_bf16(e.x), b = unpack_bf16(e.y);
  const float4 v = make_float4(a.x + pe.x, a.y + pe.y, b.x + pe.z, b.y + pe.w);
  reinterpret_cast<float4*>(h + static_cast<long long>(r) * dim)[c4] = v;
  uint2 o;
  o.x = pack_bf16(v.x, v.y);
  o.y = pack_bf16(v.z, v.w);
  reinterpret_cast<uint2*>(hb + static_cast<long long>(r) * dim)[c4] = o;
  auto block_sum = [&](float x) {
    x = warp_sum(x);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    float t2 = lane < nw ? s_red[lane] : 0.f;
    t2 = warp_sum(t2);
    __syncthreads();
    return t2;
  };
  const float sum = block_sum((v.x + v.y) + (v.z + v.w));
  const float mean = sum / static_cast<float>(dim);
  const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  const float m2 = block_sum((dx * dx + dy * dy) + (dz * dz + dw * dw));
  if (threadIdx.x == 0) stat[r] = make_float2(sum, m2);
  tr.flush();
}

// next_ids[r] = argmax over candidates (vc_gpt2_forward's next_ids on the candidate path)
__global__ void __launch_bounds__(256) dc_argmax_cands_kernel(const float* __restrict__ cand_v, const int* __restrict__ cand_i, int n_cand,
                                                              int32_t* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int tok = block_argmax_cands(cand_v + static_cast<size_t>(blockIdx.x) * n_cand, cand_i + static_cast<size_t>(blockIdx.x) * n_cand, n_cand);
  if (threadIdx.x == 0) out[blockIdx.x] = tok;
}

// opt-in to > 48 KB of dynamic shared memory, once per (device, kernel): a second GPU in the same process needs its own
template <typename Kern>
int set_smem(Kern k, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;
  int dev = 0;
  VC_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  size_t& d = done[std::make_pair(dev, reinterpret_cast<const void*>(k))];
  if (d < bytes + 1) {
    if (bytes > 48 * 1024) VC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    // every kernel of the chain asks for the largest shared-memory carve-out: an SM only takes a CTA of the next kernel beside a
    // running one if its current L1/shared split already has the room (a different split waits for the SM to drain)
    VC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    d = bytes + 1;
  }
  return 0;
}

template <int NF, int KB, int EPI>
int launch_fullk(const __nv_bfloat16* hb, const float2* stat, int n_part, const __nv_bfloat16* W, const float* cs, const float* bf,
                 __nv_bfloat16* out, int M, int N, float eps, const void* next_w, long long next_bytes, cudaStream_t s) {
  constexpr int K = DC_THREADS * KB;
  constexpr size_t smem = 2 * 32 * (K * 2 + 64) + DC_WARPS * 32 * red_pitch(NF) * 4 + 2 * DC_MAX_PARTS * 32 * 8;
  if (int e = set_smem(dc_fullk_kernel<NF, KB, EPI>, smem)) return e;
  VC_LAUNCH(EPI ? "dc_fc1_gelu" : "dc_qkv", static_cast<double>(N) * K * 2.0, s,
            VC_CUDA_OK(launch_pdl(dc_fullk_kernel<NF, KB, EPI>, dim3(N / (8 * NF)), dim3(DC_THREADS), smem, s, hb, stat, n_part, W, cs, bf, out, M, N, eps,
                                  next_w, next_bytes)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int NT, int NSPLIT, int KQ, int KBW>
int launch_nk(const __nv_bfloat16* x, const __nv_bfloat16* W, const float* bias, float* h, __nv_bfloat16* hb, float2* stat, int M, int K,
              const void* next_w, long long next_bytes, const char* name, cudaStream_t s) {
  constexpr int FT = NT * 8, KSLICE = KQ * KBW * 32;
  constexpr int PITCH = (KSLICE * 2) % 128 == 64 ? KSLICE * 2 : KSLICE * 2 + 64;
  constexpr size_t xs_bytes = 64 * PITCH > KQ * 64 * (FT + 8) * 4 ? 64 * PITCH : KQ * 64 * (FT + 8) * 4;
  constexpr size_t smem = xs_bytes + 2 * DC_CLUSTER * 8 * FT * 4;
  VC_REQUIRE(K == DC_CLUSTER * KSLICE, "dc_nk: K=%d does not match the compiled slice %d x 8", K, KSLICE);
  if (int e = set_smem(dc_nk_kernel<NT, NSPLIT, KQ, KBW>, smem)) return e;
  VC_LAUNCH(name, static_cast<double>(DC_NCLUSTERS * FT) * K * 2.0, s,
            VC_CUDA_OK(launch_pdl(dc_nk_kernel<NT, NSPLIT, KQ, KBW>, dim3(DC_NCLUSTERS * DC_CLUSTER), dim3(DC_THREADS), smem, s, x, W, bias, h, hb, stat, M, K,
                                  next_w, next_bytes)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
int chain_set_trace(void* buf, int max_records) {
  unsigned long long* p = static_cast<unsigned long long*>(buf);
  VC_CUDA_OK(cudaMemcpyToSymbol(dc_trace, &p, sizeof(p)));
  VC_CUDA_OK(cudaMemcpyToSymbol(dc_trace_max, &max_records, sizeof(int)));
  return 0;
}

bool chain_supported(const VcGptWeights* w, int rows) {
  if (w == nullptr || w->layer == nullptr || w->lmh_w == nullptr || w->lmh_cs == nullptr || w->lmh_b == nullptr) return false;
  if (!(w->dim == 768 || w->dim == 1024) || w->heads * 64 != w->dim || w->vocab_pad % 16 != 0) return false;
  if (rows <= 0 || rows > (getenv("VC_DECODE_CHAIN") != nullptr && atoi(getenv("VC_DECODE_CHAIN")) == 2 ? DC_MAX_ROWS : DC_BEST_ROWS)) return false;
  for (int l = 0; l < w->layers; ++l)
    if (w->layer[l].attn_wf == nullptr || w->layer[l].fc_wf == nullptr) return false;
  return true;
}

int chain_lmhead_ctas() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    n = sms;
  }
  return n;
}

int chain_add_pos_stats(const VcGptWeights* w, const float* embeds, const ChainBuffers& b, int n_seq, int L, int past_len, cudaStream_t s) {
  const int rows = n_seq * L, H = w->dim;
  if (rows <= 0) return 0;
  if (int e = set_smem(dc_add_pos_stats_kernel, 0)) return e;
  VC_LAUNCH("dc_add_pos_stats", rows * H * 14.0, s,
            VC_CUDA_OK(launch_pdl(dc_add_pos_stats_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, embeds, w->wpe, b.h, static_cast<__nv_bfloat16*>(b.hb),
                                  reinterpret_cast<float2*>(b.stat), rows, L, past_len, H, static_cast<const void*>(w->layer[0].attn_wf),
                                  static_cast<long long>(3) * H * H * 2)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// All layers + lm_head over rows = n_seq * L; expects h / hb / stat[0] of the input rows (chain_add_pos_stats or the
// previous step's selection kernel).  first_parts: 1.
int chain_layers(const VcGptWeights* w, const ChainBuffers& b, int n_seq, int L, int past_len, VcKvCache* cache, float* logits, long long ld,
                 cudaStream_t s) {
  const int H = w->dim, M = n_seq * L;
  const float eps = 1e-5f;
  auto* hb = static_cast<__nv_bfloat16*>(b.hb);
  auto* qkv = static_cast<__nv_bfloat16*>(b.qkv);
  auto* att = static_cast<__nv_bfloat16*>(b.att);
  auto* hid = static_cast<__nv_bfloat16*>(b.hid);
  auto* stat = reinterpret_cast<float2*>(b.stat);
  const long long bytes_hh = static_cast<long long>(H) * H * 2;
  int e;
  int parts = 1;
  for (int l = 0; l < w->layers; ++l) {
    const VcGptLayer& Ly = w->layer[l];
    const bool last = l + 1 == w->layers;
    const auto* wq = static_cast<const __nv_bfloat16*>(Ly.attn_wf);
    const auto* wf = static_cast<const __nv_bfloat16*>(Ly.fc_wf);
    if (H == 768) e = launch_fullk<2, 3, 0>(hb, stat, parts, wq, Ly.attn_cs, Ly.attn_bf, qkv, M, 3 * H, eps, Ly.aproj_w, bytes_hh, s);
    else e = launch_fullk<3, 4, 0>(hb, stat, parts, wq, Ly.attn_cs, Ly.attn_bf, qkv, M, 3 * H, eps, Ly.aproj_w, bytes_hh, s);
    if (e) return e;
    if ((e = gpt_attention(qkv, nullptr, 0, nullptr, att, cache, l, n_seq, L, past_len, s))) return e;
    if (H == 768) e = launch_nk<6, 6, 1, 3>(att, static_cast<const __nv_bfloat16*>(Ly.aproj_w), Ly.aproj_b, b.h, hb, stat, M, H, Ly.fc_wf, 4 * bytes_hh, "dc_proj", s);
    else e = launch_nk<8, 8, 1, 4>(att, static_cast<const __nv_bfloat16*>(Ly.aproj_w), Ly.aproj_b, b.h, hb, stat, M, H, Ly.fc_wf, 4 * bytes_hh, "dc_proj", s);
    if (e) return e;
    parts = DC_NCLUSTERS;
    if (H == 768) e = launch_fullk<3, 3, 1>(hb, stat, parts, wf, Ly.fc_cs, Ly.fc_bf, hid, M, 4 * H, eps, Ly.mproj_w, 4 * bytes_hh, s);
    else e = launch_fullk<4, 4, 1>(hb, stat, parts, wf, Ly.fc_cs, Ly.fc_bf, hid, M, 4 * H, eps, Ly.mproj_w, 4 * bytes_hh, s);
    if (e) return e;
    const void* nxt = last ? w->lmh_w : w->layer[l + 1].attn_wf;
    const long long nxt_bytes = last ? static_cast<long long>(8) * 1024 * 1024 : 3 * bytes_hh;   // head of the lm_head stream
    if (H == 768) e = launch_nk<6, 2, 4, 3>(hid, static_cast<const __nv_bfloat16*>(Ly.mproj_w), Ly.mproj_b, b.h, hb, stat, M, 4 * H, nxt, nxt_bytes, "dc_fc2", s);
    else e = launch_nk<8, 2, 4, 4>(hid, static_cast<const __nv_bfloat16*>(Ly.mproj_w), Ly.mproj_b, b.h, hb, stat, M, 4 * H, nxt, nxt_bytes, "dc_fc2", s);
    if (e) return e;
  }
  // ln_f + tied lm_head on the last position of every sequence (HF computes every position; only the last is read)
  const int G = chain_lmhead_ctas();
  const dim3 grid(G, (n_seq + 63) / 64);
  const auto* lw = static_cast<const __nv_bfloat16*>(w->lmh_w);
  // tcgen05 version (dc_lmhead_tc_kernel) when every CTA's vocabulary range fits its TMEM accumulators; VC_LMHEAD_TC=0: A/B switch
  const char* tc_env = getenv("VC_LMHEAD_TC");                 // read per call: the test flips it
  const bool tc_off = tc_env != nullptr && atoi(tc_env) == 0;
  const int max_range = (((w->vocab_pad >> 4) + G - 1) / G) * 16;
  if (!tc_off && H % 64 == 0 && max_range <= LT_SUB * LT_NT && lt_smem(H) <= 227 * 1024) {
    CUtensorMap ta, tw;
    // A: row r = last position of sequence r; rows beyond n_seq read as zeros
    if ((e = make_tmap_bf16_kmajor_ld(&ta, hb + static_cast<size_t>(L - 1) * H, n_seq, H, static_cast<long long>(L) * H, 64))) return e;
    if ((e = make_tmap_bf16_kmajor(&tw, lw, w->vocab_pad, H, LT_NT))) return e;
    if ((e = set_smem(dc_lmhead_tc_kernel, static_cast<size_t>(lt_smem(H))))) return e;
    VC_LAUNCH("dc_lm_head_tc", static_cast<double>(w->vocab_pad) * H * 2.0, s,
              VC_CUDA_OK(launch_pdl(dc_lmhead_tc_kernel, grid, dim3(LT_THREADS), static_cast<size_t>(lt_smem(H)), s, ta, tw, H, stat, parts,
                                    static_cast<long long>(M), static_cast<long long>(L), static_cast<long long>(L - 1), lw, w->lmh_cs, w->lmh_b,
                                    w->vocab, w->vocab_pad, n_seq, eps, logits, ld, b.cand_v, b.cand_i)));
    VC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (H == 768) {
    constexpr size_t smem = 64 * (768 * 2 + 64) + 64 * 8 + DC_MAX_PARTS * 64 * 8 + 2 * DC_WARPS * 64 * 4;
    if ((e = set_smem(dc_lmhead_kernel<3>, smem))) return e;
    VC_LAUNCH("dc_lm_head", static_cast<double>(w->vocab_pad) * H * 2.0, s,
              VC_CUDA_OK(launch_pdl(dc_lmhead_kernel<3>, grid, dim3(DC_THREADS), smem, s, hb, stat, parts, static_cast<long long>(M), static_cast<long long>(L),
                                    static_cast<long long>(L - 1), lw, w->lmh_cs, w->lmh_b, w->vocab, w->vocab_pad, n_seq, eps, logits, ld, b.cand_v, b.cand_i)));
  } else {
    constexpr size_t smem = 64 * (1024 * 2 + 64) + 64 * 8 + DC_MAX_PARTS * 64 * 8 + 2 * DC_WARPS * 64 * 4;
    if ((e = set_smem(dc_lmhead_kernel<4>, smem))) return e;
    VC_LAUNCH("dc_lm_head", static_cast<double>(w->vocab_pad) * H * 2.0, s,
              VC_CUDA_OK(launch_pdl(dc_lmhead_kernel<4>, grid, dim3(DC_THREADS), smem, s, hb, stat, parts, static_cast<long long>(M), static_cast<long long>(L),
                                    static_cast<long long>(L - 1), lw, w->lmh_cs, w->lmh_b, w->vocab, w->vocab_pad, n_seq, eps, logits, ld, b.cand_v, b.cand_i)));
  }
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int chain_select(const VcGptWeights* w, const ChainBuffers& b, int n_seq, int step, int max_new, int eos, int32_t* finished, int32_t* ids_out,
                 int32_t* len_out, const int32_t* forced, int next_pos, bool feed_next, int32_t* next_ids, cudaStream_t s) {
  const int H = w->dim;
  if (int e = set_smem(dc_select_kernel, 0)) return e;
  VC_LAUNCH("dc_select", static_cast<double>(n_seq) * chain_lmhead_ctas() * 8.0, s,
            VC_CUDA_OK(launch_pdl(dc_select_kernel, dim3(n_seq), dim3(H / 4), 0, s, static_cast<const float*>(b.cand_v), static_cast<const int*>(b.cand_i),
                                  chain_lmhead_ctas(), step, max_new, eos, finished, ids_out, len_out, forced, static_cast<const __nv_bfloat16*>(w->wte),
                                  w->wpe, next_pos, H, feed_next ? b.h : nullptr, static_cast<__nv_bfloat16*>(b.hb), reinterpret_cast<float2*>(b.stat),
                                  next_ids, static_cast<const void*>(w->layer[0].attn_wf), static_cast<long long>(3) * H * H * 2)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int chain_argmax(const ChainBuffers& b, int n_seq, int32_t* out, cudaStream_t s) {
  VC_LAUNCH("dc_argmax_cands", static_cast<double>(n_seq) * chain_lmhead_ctas() * 8.0, s,
            VC_CUDA_OK(launch_pdl(dc_argmax_cands_kernel, dim3(n_seq), dim3(256), 0, s, static_cast<const float*>(b.cand_v), static_cast<const int*>(b.cand_i),
                                  chain_lmhead_ctas(), out)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vc
