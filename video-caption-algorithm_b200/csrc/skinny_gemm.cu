// Weight-streaming "skinny" GEMM for the GPT-2 decode step (M = live sequences <= 128):
//     P[s][m][n] = sum_{k in K-slice s} x[m][k] * W[n][k]          (fp32 partials, deterministic)
//
// The decode step is HBM-bound: every step streams all 247 MB of bf16 weights once
// (SURVEY.md §8a9).  A 128x256 tcgen05 tile gives only N/256 CTAs for M=64 (3..12 CTAs), far
// too few to pull HBM bandwidth, so this kernel instead spreads (64-feature block) x (K slice)
// over >= 144 CTAs.  Each warp owns 16 output features and streams their weight rows straight
// from global memory with 128-bit loads (double-buffered in registers, no shared memory), feeds
// them to mma.sync m16n8k16 as the A operand, and takes the (tiny, L1-resident) activations as
// the B operand.  K is consumed in a lane-permuted order (each lane's 16-byte load covers slots
// {2t,2t+1,2t+8,2t+9} of two MMAs) — a dot product does not care, and it lets both operands
// use plain 16-byte loads.  Split-K partials go to a small fp32 buffer and are summed in a fixed
// order by the consumer (residual+LayerNorm, bias+GELU, or the attention kernel), so results do
// not depend on scheduling.
//
// Consumers in this file: resid_ln (h += bias + sum_s P; xn = LN(h)) and bias_act.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <algorithm>

namespace vc {

#define VC_LAUNCH(name, work, stream, ...)        \
  do {                                            \
    vc::KernelScope _ks(name, work, stream);      \
    __VA_ARGS__;                                  \
  } while (0)

namespace {

constexpr int SK_WARPS = 4;             // 4 warps x 16 features = 64 features per CTA
constexpr int SK_BN = SK_WARPS * 16;
constexpr int SK_MT = 64;               // sequences per CTA pass (8 n-tiles of 8)
constexpr int SK_KB = 64;               // k elements per register batch (2 MMA pairs)
constexpr int SK_STAGE_MAX = 384;       // widest K slice whose [64 x slice] activation block is staged in shared memory (53 KB)

// Weights are streamed once per step: keep them out of L1 and mark them first-to-evict in L2, so that 247 MB of decoder
// weights per step do not push the concurrently running encoder's working set out of the 126 MB L2 (CaptionPipeline).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ldg_stream(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// grid = (N/64, ksplit, ceil(M/64)), block = 128.  STAGE: the activation slice (64 rows x kslice) is copied to shared
// memory once per CTA (asynchronous 16-byte copies, all in flight together) instead of each warp pulling its B fragments
// through L1 batch by batch: one memory round trip for x instead of one per 64-k batch.
// The weight rows are constants, so the first two register batches are requested BEFORE griddepcontrol.wait: they stream
// in from HBM while the previous kernel of the chain is still running.
// Register budget: the staged variant must fit beside an encoder GEMM CTA (448 threads x 104 registers leave 19K per SM),
// so it is capped at 144 registers; the unstaged one (lm_head) only needs three CTAs per SM.
template <bool STAGE>
__global__ void __maxnreg__(STAGE ? 144 : 168) skinny_gemm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ W,
                                                                   float* __restrict__ P, int M, int N, int K, int kslice) {
  extern __shared__ __align__(16) uint8_t sk_xs[];       // STAGE: 64 rows, pitch kslice*2 + 64 bytes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int f0 = blockIdx.x * SK_BN + warp * 16;
  const int k0 = blockIdx.y * kslice;
  const int m0 = blockIdx.z * SK_MT;
  const __nv_bfloat16* w_lo = W + static_cast<size_t>(f0 + g) * K + k0 + 8 * t;        // feature row g
  const __nv_bfloat16* w_hi = w_lo + static_cast<size_t>(8) * K;                        // feature row g+8
  const int nb = kslice / SK_KB;
  // two register buffers with compile-time names (A/B) so nothing lands in local memory
  uint4 wA[4], wB[4];   // {lo step0, hi step0, lo step1, hi step1}
  const uint64_t pol = l2_evict_first_policy();
  auto load = [&](uint4 (&w)[4], int b) {
    const int o = b * SK_KB;
    w[0] = ldg_stream(w_lo + o, pol);      w[1] = ldg_stream(w_hi + o, pol);
    w[2] = ldg_stream(w_lo + o + 32, pol); w[3] = ldg_stream(w_hi + o + 32, pol);
  };
  load(wA, 0);
  if (nb > 1) load(wB, 1);
  pdl_wait();
  pdl_launch_dependents();
  const int pitch = kslice * 2 + 64;
  // activation rows for this lane's B fragments: sequence m0 + nt*8 + g (clamped; masked at the store)
  const __nv_bfloat16* xrow[8];
  if (STAGE) {
    const int cpr = kslice >> 3;
    for (int r = warp; r < SK_MT; r += SK_WARPS) {
      int m = m0 + r;
      m = m < M ? m : M - 1;
      const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<size_t>(m) * K + k0);
      for (int col = lane; col < cpr; col += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sk_xs + r * pitch + col * 16)), "l"(src + col) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  } else {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      int m = m0 + nt * 8 + g;
      m = m < M ? m : M - 1;
      xrow[nt] = x + static_cast<size_t>(m) * K + k0 + 8 * t;
    }
  }
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;

  auto compute = [&](const uint4 (&w)[4], int b) {
#pragma unroll
    for (int st = 0; st < 2; ++st) {
      const uint4 a = w[2 * st], c = w[2 * st + 1];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint4 xb;
        if (STAGE) xb = *reinterpret_cast<const uint4*>(sk_xs + (nt * 8 + g) * pitch + (b * SK_KB + st * 32 + 8 * t) * 2);
        else xb = __ldg(reinterpret_cast<const uint4*>(xrow[nt] + b * SK_KB + st * 32));
        mma16816(acc[nt], a.x, c.x, a.y, c.y, xb.x, xb.y);     // slots from elements 0..3 of each 16-byte load
        mma16816(acc[nt], a.z, c.z, a.w, c.w, xb.z, xb.w);     // slots from elements 4..7
      }
    }
  };
  for (int b = 0; b < nb; b += 2) {
    compute(wA, b);
    if (b + 2 < nb) load(wA, b + 2);
    if (b + 1 < nb) {
      compute(wB, b + 1);
      if (b + 3 < nb) load(wB, b + 3);
    }
  }
  // C fragment: (feature f0+g, seq 2t,2t+1) and (feature f0+g+8, ...)
  float* Pp = P + static_cast<size_t>(blockIdx.y) * M * N;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int m = m0 + nt * 8 + 2 * t;
    if (m < M) {
      Pp[static_cast<size_t>(m) * N + f0 + g] = acc[nt][0];
      Pp[static_cast<size_t>(m) * N + f0 + g + 8] = acc[nt][2];
    }
    if (m + 1 < M) {
      Pp[static_cast<size_t>(m + 1) * N + f0 + g] = acc[nt][1];
      Pp[static_cast<size_t>(m + 1) * N + f0 + g + 8] = acc[nt][3];
    }
  }
}

// out[m][n] = bf16(gelu_new(bias[n] + sum_k x[m][k] W[n][k])) — the fc1 product of a decode step with its activation fused,
// so the step has no separate bias+GELU kernel (one dependent launch less per layer).  GELU needs the complete sum, so K is
// split across the four warps of a CTA (not across CTAs) and reduced through shared memory: grid = (N/16, 1, ceil(M/64)),
// one 16-feature tile per CTA, warp kg owns K/4.  Small enough (128 threads, <= 144 registers, 17 KB) to run beside an
// encoder GEMM CTA; the price is that every CTA pulls the whole activation block through L1 (~2 us of SM fill).
constexpr int SKA_RED = 64 * 17;
__global__ void __maxnreg__(144) skinny_gemm_gelu_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ W,
                                                         const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int M, int N, int K) {
  __shared__ float s_red[SK_WARPS * SKA_RED];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int f0 = blockIdx.x * 16;
  const int kq = K / SK_WARPS, k0 = warp * kq;
  const int m0 = blockIdx.z * SK_MT;
  const __nv_bfloat16* w_lo = W + static_cast<size_t>(f0 + g) * K + k0 + 8 * t;
  const __nv_bfloat16* w_hi = w_lo + static_cast<size_t>(8) * K;
  const int nb = kq / SK_KB;
  uint4 wA[4], wB[4];
  const uint64_t pol = l2_evict_first_policy();
  auto load = [&](uint4 (&w)[4], int b) {
    const int o = b * SK_KB;
    w[0] = ldg_stream(w_lo + o, pol);      w[1] = ldg_stream(w_hi + o, pol);
    w[2] = ldg_stream(w_lo + o + 32, pol); w[3] = ldg_stream(w_hi + o + 32, pol);
  };
  load(wA, 0);
  if (nb > 1) load(wB, 1);
  const int tid = threadIdx.x;
  float bias_r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bias_r[i] = __ldg(bias + f0 + ((tid + i * 128) & 15));
  pdl_wait();
  pdl_launch_dependents();
  const __nv_bfloat16* xrow[8];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    int m = m0 + nt * 8 + g;
    m = m < M ? m : M - 1;
    xrow[nt] = x + static_cast<size_t>(m) * K + k0 + 8 * t;
  }
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  auto compute = [&](const uint4 (&w)[4], int b) {
#pragma unroll
    for (int st = 0; st < 2; ++st) {
      const uint4 a = w[2 * st], c = w[2 * st + 1];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint4 xb = __ldg(reinterpret_cast<const uint4*>(xrow[nt] + b * SK_KB + st * 32));
        mma16816(acc[nt], a.x, c.x, a.y, c.y, xb.x, xb.y);
        mma16816(acc[nt], a.z, c.z, a.w, c.w, xb.z, xb.w);
      }
    }
  };
  for (int b = 0; b < nb; b += 2) {
    compute(wA, b);
    if (b + 2 < nb) load(wA, b + 2);
    if (b + 1 < nb) {
      compute(wB, b + 1);
      if (b + 3 < nb) load(wB, b + 3);
    }
  }
  float* rw = s_red + warp * SKA_RED;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int m = nt * 8 + 2 * t;
    rw[m * 17 + g] = acc[nt][0];
    rw[(m + 1) * 17 + g] = acc[nt][1];
    rw[m * 17 + g + 8] = acc[nt][2];
    rw[(m + 1) * 17 + g + 8] = acc[nt][3];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = tid + i * 128, m = idx >> 4, f = idx & 15;
    if (m0 + m < M) {
      const float* rr = s_red + m * 17 + f;
      const float a = bias_r[i] + (((rr[0] + rr[SKA_RED]) + rr[2 * SKA_RED]) + rr[3 * SKA_RED]);      // fixed order
      out[static_cast<size_t>(m0 + m) * N + f0 + f] = __float2bfloat16_rn(gelu_tanh(a));
    }
  }
}

// h[row] += bias + sum_s P[s][row];  xn[row] = LayerNorm(h[row]) (bf16).
// One CTA per row, one float4 column per thread (dim/4 threads): the h load and all ksplit partial
// loads of a thread are independent, so the whole row costs about one memory round trip.
__device__ __forceinline__ float block_sum_small(float v, float* s_red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = lane < nw ? s_red[lane] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}
constexpr int RL_MAX_KS = 64;
constexpr int RL_BATCH = 12;            // partial loads in flight per thread
__global__ void __launch_bounds__(256) resid_ln_kernel(float* __restrict__ h, const float* __restrict__ P, int ksplit,
                                                       const float* __restrict__ bias, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, __nv_bfloat16* __restrict__ xn, int rows, int dim,
                                                       float eps) {
  __shared__ float s_red[8];
  const int row = blockIdx.x, c4 = threadIdx.x;          // blockDim.x == dim / 4
  // parameters are constants: requested before the previous kernel of the chain has finished
  const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
  const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  pdl_wait();
  pdl_launch_dependents();
  const size_t plane4 = static_cast<size_t>(rows) * dim / 4;
  const size_t o4 = static_cast<size_t>(row) * (dim / 4) + c4;
  float4 a = reinterpret_cast<const float4*>(h)[o4];
  const float4* P4 = reinterpret_cast<const float4*>(P) + o4;
  float4 acc = b;
  for (int s0 = 0; s0 < ksplit; s0 += RL_BATCH) {        // fixed summation order, RL_BATCH loads in flight
    float4 q[RL_BATCH];
#pragma unroll
    for (int i = 0; i < RL_BATCH; ++i)
      if (s0 + i < ksplit) q[i] = P4[(s0 + i) * plane4];
#pragma unroll
    for (int i = 0; i < RL_BATCH; ++i)
      if (s0 + i < ksplit) { acc.x += q[i].x; acc.y += q[i].y; acc.z += q[i].z; acc.w += q[i].w; }
  }
  a.x += acc.x; a.y += acc.y; a.z += acc.z; a.w += acc.w;
  reinterpret_cast<float4*>(h)[o4] = a;
  const float mean = block_sum_small((a.x + a.y) + (a.z + a.w), s_red) / static_cast<float>(dim);
  const float dx = a.x - mean, dy = a.y - mean, dz = a.z - mean, dw = a.w - mean;
  const float rstd = rsqrtf(block_sum_small((dx * dx + dy * dy) + (dz * dz + dw * dw), s_red) / static_cast<float>(dim) + eps);
  uint2 w;
  w.x = pack_bf16(dx * rstd * gm.x + bt.x, dy * rstd * gm.y + bt.y);
  w.y = pack_bf16(dz * rstd * gm.z + bt.z, dw * rstd * gm.w + bt.w);
  reinterpret_cast<uint2*>(xn)[o4] = w;
}

// out[m][n] = bf16(act(bias[n] + sum_s P[s][m][n]))
__global__ void __launch_bounds__(256) bias_act_kernel(const float* __restrict__ P, int ksplit, const float* __restrict__ bias,
                                                       __nv_bfloat16* __restrict__ out, int M, int N, int gelu) {
  const size_t i4 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t total4 = static_cast<size_t>(M) * N / 4;
  const int n4 = static_cast<int>(i4 % (N / 4));
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i4 < total4) a = __ldg(reinterpret_cast<const float4*>(bias) + n4);     // constant: before the wait
  pdl_wait();
  pdl_launch_dependents();
  if (i4 >= total4) return;
  for (int s0 = 0; s0 < ksplit; s0 += 8) {               // fixed order, 8 loads in flight
    float4 q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (s0 + i < ksplit) q[i] = *(reinterpret_cast<const float4*>(P) + (s0 + i) * total4 + i4);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (s0 + i < ksplit) { a.x += q[i].x; a.y += q[i].y; a.z += q[i].z; a.w += q[i].w; }
  }
  if (gelu) { a.x = gelu_tanh(a.x); a.y = gelu_tanh(a.y); a.z = gelu_tanh(a.z); a.w = gelu_tanh(a.w); }
  uint2 w;
  w.x = pack_bf16(a.x, a.y);
  w.y = pack_bf16(a.z, a.w);
  *(reinterpret_cast<uint2*>(out) + i4) = w;
}

}  // namespace

// K-slices of a multiple of 64; aim for >= ~200 CTAs (row tiles included) without exploding the partial buffer: the fp32
// partials are written and read back through L2 once per slice, which is the dominant traffic of a step beyond 64 rows.
// Slices stay <= 384 wide where possible so that the activation block is staged in shared memory; at four row tiles this
// keeps every product of GPT-2 small within one wave of CTAs (384-432 CTAs at three per SM).
int skinny_ksplit(int N, int K, int row_tiles) {
  const int blocks = (N / SK_BN) * (row_tiles > 0 ? row_tiles : 1);
  const int kb = K / SK_KB;
  if (N / SK_BN >= 296) return 1;
  int best = 1;
  for (int ks = 1; ks <= kb; ++ks) {
    if (kb % ks) continue;
    best = ks;
    if (blocks * ks >= 200 && K / ks <= SK_STAGE_MAX) break;
  }
  return best;
}
int skinny_ksplit(int N, int K) { return skinny_ksplit(N, K, 1); }

int skinny_gemm(const void* x, const void* W, float* P, int M, int N, int K, int ksplit, cudaStream_t s) {
  VC_REQUIRE(M > 0 && N % SK_BN == 0 && K % SK_KB == 0, "skinny_gemm: M=%d N=%d (%%64) K=%d (%%64)", M, N, K);
  VC_REQUIRE(ksplit >= 1 && (K / SK_KB) % ksplit == 0, "skinny_gemm: ksplit=%d does not divide K/64=%d", ksplit, K / SK_KB);
  dim3 grid(N / SK_BN, ksplit, (M + SK_MT - 1) / SK_MT);
  const int kslice = K / ksplit;
  if (kslice <= SK_STAGE_MAX) {
    const size_t smem = static_cast<size_t>(SK_MT) * (kslice * 2 + 64);
    static PerDeviceOnce attr;
    if (attr.first()) VC_CUDA_OK(cudaFuncSetAttribute(skinny_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_MT * (SK_STAGE_MAX * 2 + 64)));
    VC_LAUNCH("skinny_gemm", static_cast<double>(N) * K * 2.0, s,
              VC_CUDA_OK(launch_pdl(skinny_gemm_kernel<true>, grid, dim3(SK_WARPS * 32), smem, s, static_cast<const __nv_bfloat16*>(x),
                                    static_cast<const __nv_bfloat16*>(W), P, M, N, K, kslice)));
  } else {
    VC_LAUNCH("skinny_gemm", static_cast<double>(N) * K * 2.0, s,
              VC_CUDA_OK(launch_pdl(skinny_gemm_kernel<false>, grid, dim3(SK_WARPS * 32), 0, s, static_cast<const __nv_bfloat16*>(x),
                                    static_cast<const __nv_bfloat16*>(W), P, M, N, K, kslice)));
  }
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int skinny_gemm_gelu(const void* x, const void* W, const float* bias, void* out, int M, int N, int K, cudaStream_t s) {
  VC_REQUIRE(M > 0 && N % 16 == 0 && K % (SK_WARPS * SK_KB) == 0, "skinny_gemm_gelu: M=%d N=%d (%%16) K=%d (%%256)", M, N, K);
  dim3 grid(N / 16, 1, (M + SK_MT - 1) / SK_MT);
  VC_LAUNCH("skinny_gemm_gelu", static_cast<double>(N) * K * 2.0, s,
            VC_CUDA_OK(launch_pdl(skinny_gemm_gelu_kernel, grid, dim3(SK_WARPS * 32), 0, s, static_cast<const __nv_bfloat16*>(x),
                                  static_cast<const __nv_bfloat16*>(W), bias, static_cast<__nv_bfloat16*>(out), M, N, K)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int resid_ln(float* h, const float* P, int ksplit, const float* bias, const float* gamma, const float* beta, void* xn, int rows, int dim,
             float eps, cudaStream_t s) {
  VC_REQUIRE(dim % 128 == 0 && dim <= 1024 && ksplit <= RL_MAX_KS, "resid_ln: dim=%d ksplit=%d", dim, ksplit);
  if (rows <= 0) return 0;
  VC_LAUNCH("resid_ln", static_cast<double>(rows) * dim * (10.0 + 4.0 * ksplit), s,
            VC_CUDA_OK(launch_pdl(resid_ln_kernel, dim3(rows), dim3(dim / 4), 0, s, h, P, ksplit, bias, gamma, beta, static_cast<__nv_bfloat16*>(xn),
                                  rows, dim, eps)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int bias_act(const float* P, int ksplit, const float* bias, void* out, int M, int N, int gelu, cudaStream_t s) {
  VC_REQUIRE(N % 4 == 0, "bias_act: N=%d", N);
  const size_t total4 = static_cast<size_t>(M) * N / 4;
  if (total4 == 0) return 0;
  VC_LAUNCH("bias_act", static_cast<double>(M) * N * (2.0 + 4.0 * ksplit), s,
            VC_CUDA_OK(launch_pdl(bias_act_kernel, dim3(static_cast<unsigned>((total4 + 255) / 256)), dim3(256), 0, s, P, ksplit, bias,
                                  static_cast<__nv_bfloat16*>(out), M, N, gelu)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vc
