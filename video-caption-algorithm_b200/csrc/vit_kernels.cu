// Bandwidth-bound and attention kernels of the ViT side of the path:
//   preprocessing (core/preprocessing/frame_loader.py:34-47), LayerNorm, non-causal
//   attention over the 197/257 tokens of one frame, class-token rows, and the fused
//   cls-pool + temporal mean + proj + prefix-norm + mapper kernel that replaces the
//   reference's two CuPy operators (core/operators/cupy_vit_pool.py:23-104,
//   core/operators/cupy_linear_mapper.py:14-70) and core/engine.py:45-50.
#include "vc_common.cuh"
#include "vc_kernels.h"
#include <algorithm>
#include <cstdlib>

namespace vc {

#define VC_LAUNCH(name, work, stream, ...)        \
  do {                                            \
    vc::KernelScope _ks(name, work, stream);      \
    __VA_ARGS__;                                  \
  } while (0)

// =========================================================================== preprocessing
// layout 1, patch 16: one thread owns 16 pixels of one image row inside one patch:
// 48 contiguous input bytes (3 x 16 B loads), three 32-byte output runs (one per channel).
__global__ void __launch_bounds__(256) preprocess_patch16_kernel(const uint8_t* __restrict__ in, const float* __restrict__ lut,
                                                                 __nv_bfloat16* __restrict__ out, int n_frames, int H, int W,
                                                                 int k_pad) {
  __shared__ __nv_bfloat16 s_lut[3 * 256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = __float2bfloat16_rn(lut[i]);
  __syncthreads();
  const int gw = W / 16, gh = H / 16;
  const long long total = static_cast<long long>(n_frames) * H * gw;   // (frame, y, px)
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(idx % gw);
    const long long fy = idx / gw;
    const int y = static_cast<int>(fy % H);
    const long long f = fy / H;
    const uint4* src = reinterpret_cast<const uint4*>(in + (fy * W + px * 16) * 3);
    uint4 raw[3];
    raw[0] = __ldg(src); raw[1] = __ldg(src + 1); raw[2] = __ldg(src + 2);
    const uint8_t* b = reinterpret_cast<const uint8_t*>(raw);
    const int py = y >> 4, i = y & 15;
    __nv_bfloat16* dst = out + ((f * gh + py) * gw + px) * static_cast<long long>(k_pad) + i * 16;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      __align__(16) __nv_bfloat16 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = s_lut[c * 256 + b[j * 3 + c]];
      uint4* d = reinterpret_cast<uint4*>(dst + c * 256);
      d[0] = reinterpret_cast<uint4*>(v)[0];
      d[1] = reinterpret_cast<uint4*>(v)[1];
    }
  }
}

// generic: one thread per pixel.  layout 0 -> [n,3,H,W]; layout 1 -> patch-major with any patch size.
__global__ void __launch_bounds__(256) preprocess_generic_kernel(const uint8_t* __restrict__ in, const float* __restrict__ lut,
                                                                 __nv_bfloat16* __restrict__ out, int n_frames, int H, int W,
                                                                 int layout, int patch, int k_pad) {
  __shared__ __nv_bfloat16 s_lut[3 * 256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = __float2bfloat16_rn(lut[i]);
  __syncthreads();
  const long long total = static_cast<long long>(n_frames) * H * W;
  const int gw = (layout == 1) ? W / patch : 0, gh = (layout == 1) ? H / patch : 0;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W);
    const long long fy = idx / W;
    const int y = static_cast<int>(fy % H);
    const long long f = fy / H;
    const uint8_t* p = in + idx * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const __nv_bfloat16 v = s_lut[c * 256 + p[c]];
      if (layout == 0) {
        out[((f * 3 + c) * H + y) * W + x] = v;
      } else {
        const int py = y / patch, i = y - py * patch, px = x / patch, j = x - px * patch;
        if (py < gh && px < gw)
          out[((f * gh + py) * gw + px) * static_cast<long long>(k_pad) + (c * patch + i) * patch + j] = v;
      }
    }
  }
}

// fp32 [n,3,H,W] (the reference's normalised tensor, frame_loader.py:47) -> bf16 patch-major rows.
// Lets `model.encoder(video)` accept exactly what core/engine.py:43 passes.
__global__ void __launch_bounds__(256) patchify_f32_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int n_frames,
                                                           int H, int W, int patch, int k_pad) {
  const long long total = static_cast<long long>(n_frames) * 3 * H * W;
  const int gw = W / patch, gh = H / patch;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W);
    long long r = idx / W;
    const int y = static_cast<int>(r % H);
    r /= H;
    const int c = static_cast<int>(r % 3);
    const long long f = r / 3;
    const int py = y / patch, i = y - py * patch, px = x / patch, j = x - px * patch;
    out[((f * gh + py) * gw + px) * static_cast<long long>(k_pad) + (c * patch + i) * patch + j] = __float2bfloat16_rn(in[idx]);
  }
}

int patchify_f32(const float* video, void* out, int n, int H, int W, int patch, int k_pad, cudaStream_t s) {
  VC_REQUIRE(n >= 0 && patch > 0 && H % patch == 0 && W % patch == 0, "patchify: bad shape n=%d H=%d W=%d patch=%d", n, H, W, patch);
  VC_REQUIRE(k_pad >= 3 * patch * patch, "patchify: k_pad=%d too small", k_pad);
  if (n == 0) return 0;
  if (k_pad != 3 * patch * patch) {
    const size_t rows = static_cast<size_t>(n) * (H / patch) * (W / patch);
    VC_CUDA_OK(cudaMemsetAsync(out, 0, rows * k_pad * 2, s));
  }
  const long long total = static_cast<long long>(n) * 3 * H * W;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
  VC_LAUNCH("patchify_f32", static_cast<double>(total) * 6.0, s,
            (patchify_f32_kernel<<<grid, 256, 0, s>>>(video, static_cast<__nv_bfloat16*>(out), n, H, W, patch, k_pad)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int preprocess_u8(const uint8_t* frames, const float* lut, void* out, int n, int H, int W, int layout, int patch, int k_pad,
                  cudaStream_t s) {
  VC_REQUIRE(n >= 0 && H > 0 && W > 0, "preprocess: bad shape n=%d H=%d W=%d", n, H, W);
  if (n == 0) return 0;
  VC_REQUIRE(layout == 0 || layout == 1, "preprocess: layout must be 0 (CHW) or 1 (patch-major)");
  const double bytes = static_cast<double>(n) * H * W * 3 * 3;   // 1 B read + 2 B written per value
  if (layout == 1) {
    VC_REQUIRE(patch > 0 && H % patch == 0 && W % patch == 0, "preprocess: H,W must be multiples of patch=%d", patch);
    VC_REQUIRE(k_pad >= 3 * patch * patch && k_pad % 8 == 0, "preprocess: k_pad=%d too small / unaligned", k_pad);
    if (k_pad != 3 * patch * patch) {
      const size_t rows = static_cast<size_t>(n) * (H / patch) * (W / patch);
      VC_CUDA_OK(cudaMemsetAsync(out, 0, rows * k_pad * 2, s));
    }
    if (patch == 16 && (reinterpret_cast<uintptr_t>(frames) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      const long long total = static_cast<long long>(n) * H * (W / 16);
      const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
      VC_LAUNCH("preprocess_patch16", bytes, s,
                (preprocess_patch16_kernel<<<grid, 256, 0, s>>>(frames, lut, static_cast<__nv_bfloat16*>(out), n, H, W, k_pad)));
      VC_CUDA_OK(cudaGetLastError());
      return 0;
    }
  }
  const long long total = static_cast<long long>(n) * H * W;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
  VC_LAUNCH("preprocess_generic", bytes, s,
            (preprocess_generic_kernel<<<grid, 256, 0, s>>>(frames, lut, static_cast<__nv_bfloat16*>(out), n, H, W, layout,
                                                            patch, k_pad)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =========================================================================== (residual add +) LayerNorm
// One warp per row, fp32 residual stream, two-pass statistics in registers, bf16 and/or fp32 out.
// With `delta` (bf16 [rows_total, dim], the bias-added output of the preceding proj / fc2 GEMM) the kernel first
// folds it into the residual stream: x += delta, then normalises.  The residual is written back only once per
// transformer block: LN2 normalises x + d_proj without storing it, the next block's LN1 stores x + d_proj + d_fc2.  That is the
// reference's `x = x + attn(...)` / `x + mlp(...)` (torchvision EncoderBlock.forward, timm Block.forward;
// src/models/video_encoder.py:168-171) with the Linear output in bf16 exactly as under autocast, and it keeps
// the GEMM epilogue write-only: a read-modify-write of the fp32 stream inside the epilogue made the
// N=768,K=768 projection latency-bound (250-420 TFLOP/s).
constexpr int LN_MAX_V4 = 8;   // dim <= 1024

__global__ void __launch_bounds__(256) layernorm_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ delta,
                                                        const __nv_bfloat16* __restrict__ delta2, int write_x, long long row_stride,
                                                        long long row_offset, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                                                        int rows, int dim, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long in_row = static_cast<long long>(warp) * row_stride + row_offset;
  float4* src = reinterpret_cast<float4*>(x + in_row * dim);
  const uint2* dsrc = delta ? reinterpret_cast<const uint2*>(delta + in_row * dim) : nullptr;
  const uint2* dsrc2 = delta2 ? reinterpret_cast<const uint2*>(delta2 + in_row * dim) : nullptr;
  const int nv = dim >> 7;   // float4 per lane
  float4 v[LN_MAX_V4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      v[i] = src[i * 32 + lane];
      if (dsrc != nullptr) {
        const uint2 d = dsrc[i * 32 + lane];
        const float2 a = unpack_bf16(d.x), b = unpack_bf16(d.y);
        v[i].x += a.x; v[i].y += a.y; v[i].z += b.x; v[i].w += b.y;
        if (dsrc2 != nullptr) {                      // (x + d1) + d2: the order the two residual adds happen in the model
          const uint2 e = dsrc2[i * 32 + lane];
          const float2 c = unpack_bf16(e.x), f = unpack_bf16(e.y);
          v[i].x += c.x; v[i].y += c.y; v[i].z += f.x; v[i].w += f.y;
        }
        if (write_x) src[i * 32 + lane] = v[i];
      }
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  const float mean = warp_sum(sum) / static_cast<float>(dim);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(dim) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      const float4 g = __ldg(g4 + i * 32 + lane), b = __ldg(b4 + i * 32 + lane);
      float4 y;
      y.x = (v[i].x - mean) * rstd * g.x + b.x;
      y.y = (v[i].y - mean) * rstd * g.y + b.y;
      y.z = (v[i].z - mean) * rstd * g.z + b.z;
      y.w = (v[i].w - mean) * rstd * g.w + b.w;
      const long long o = static_cast<long long>(warp) * dim + (i * 32 + lane) * 4;
      if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + o) = y;
      if (out_bf16 != nullptr) {
        uint2 w;
        w.x = pack_bf16(y.x, y.y);
        w.y = pack_bf16(y.z, y.w);
        *reinterpret_cast<uint2*>(out_bf16 + o) = w;
      }
    }
}

int add_layernorm_rows(float* x, const void* delta_bf16, const void* delta2_bf16, int write_x, long long row_stride, long long row_offset,
                       const float* g, const float* b, float* out_f32, void* out_bf16, int rows, int dim, float eps, cudaStream_t s) {
  VC_REQUIRE(delta2_bf16 == nullptr || delta_bf16 != nullptr, "layernorm: delta2 without delta");
  VC_REQUIRE(dim % 128 == 0 && dim <= 128 * LN_MAX_V4, "layernorm: dim=%d must be a multiple of 128 and <= %d", dim,
             128 * LN_MAX_V4);
  if (rows <= 0) return 0;
  const int grid = (rows + 7) / 8;
  const double bytes_per_elem = 4.0 + (delta_bf16 ? 2.0 : 0.0) + (delta2_bf16 ? 2.0 : 0.0) + (delta_bf16 && write_x ? 4.0 : 0.0) + 2.0;
  VC_LAUNCH(delta_bf16 ? "add_layernorm" : "layernorm", static_cast<double>(rows) * dim * bytes_per_elem, s,
            (layernorm_kernel<<<grid, 256, 0, s>>>(x, static_cast<const __nv_bfloat16*>(delta_bf16), static_cast<const __nv_bfloat16*>(delta2_bf16),
                                                   write_x, row_stride, row_offset, g, b, out_f32,
                                                   static_cast<__nv_bfloat16*>(out_bf16), rows, dim, eps)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm_rows(const float* x, long long row_stride, long long row_offset, const float* g, const float* b, float* out_f32,
                   void* out_bf16, int rows, int dim, float eps, cudaStream_t s) {
  return add_layernorm_rows(const_cast<float*>(x), nullptr, nullptr, 0, row_stride, row_offset, g, b, out_f32, out_bf16, rows, dim, eps, s);
}

int layernorm_f32_bf16(const float* x, const float* g, const float* b, void* out, int rows, int dim, float eps, cudaStream_t s) {
  return layernorm_rows(x, 1, 0, g, b, nullptr, out, rows, dim, eps, s);
}

// =========================================================================== row statistics for the LayerNorm-folded GEMMs
// (mean, rstd) per row from the partial (sum, sum of squares) pairs the residual-updating GEMM epilogues leave behind
// (gemm_tcgen05.cu: VC_EPI_PROJ_STATS / VC_EPI_FC2_STATS).  Combined in double: the variance is a difference of two sums.
__global__ void __launch_bounds__(256) ln_stats_finalize_kernel(const float2* __restrict__ ps, int parts, int M, int N, float eps,
                                                                float2* __restrict__ stats) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  double s = 0.0, q = 0.0;
  for (int p0 = 0; p0 < parts; p0 += 4) {
    float2 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (p0 + i < parts) v[i] = ps[static_cast<size_t>(p0 + i) * M + row];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (p0 + i < parts) { s += v[i].x; q += v[i].y; }
  }
  const double mean = s / N;
  double var = q / N - mean * mean;
  var = var > 0.0 ? var : 0.0;
  stats[row] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
}
int ln_stats_finalize(const float* pstats, int parts, int M, int N, float eps, float* stats, cudaStream_t s) {
  if (M <= 0) return 0;
  VC_LAUNCH("ln_stats_finalize", static_cast<double>(M) * (parts + 1) * 8.0, s,
            (ln_stats_finalize_kernel<<<(M + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float2*>(pstats), parts, M, N, eps,
                                                                      reinterpret_cast<float2*>(stats))));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// xb = bf16(x) and (mean, rstd) of every fp32 row: the entry of the first encoder block (the rows come from the patch-embed
// GEMM and the class-token init, neither of which sees whole rows).  One warp per row, two-pass statistics in registers.
__global__ void __launch_bounds__(256) rowstats_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats,
                                                            int rows, int dim, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(x + static_cast<long long>(warp) * dim);
  const int nv = dim >> 7;
  float4 v[LN_MAX_V4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      v[i] = src[i * 32 + lane];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      uint2 w;
      w.x = pack_bf16(v[i].x, v[i].y);
      w.y = pack_bf16(v[i].z, v[i].w);
      reinterpret_cast<uint2*>(xb + static_cast<long long>(warp) * dim)[i * 32 + lane] = w;
    }
  const float mean = warp_sum(sum) / static_cast<float>(dim);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(dim) + eps);
  if (lane == 0) stats[warp] = make_float2(mean, rstd);
}
int rowstats_cast(const float* x, void* xb_bf16, float* stats, int M, int dim, float eps, cudaStream_t s) {
  VC_REQUIRE(dim % 128 == 0 && dim <= 128 * LN_MAX_V4, "rowstats_cast: dim=%d", dim);
  if (M <= 0) return 0;
  VC_LAUNCH("rowstats_cast", static_cast<double>(M) * dim * 6.0, s,
            (rowstats_cast_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, static_cast<__nv_bfloat16*>(xb_bf16), reinterpret_cast<float2*>(stats), M, dim, eps)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =========================================================================== class-token rows
__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls_pos0, int n_frames, int tokens, int dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frames * dim) return;
  const int f = i / dim, c = i - f * dim;
  x[static_cast<long long>(f) * tokens * dim + c] = cls_pos0[c];
}
int cls_rows_init(float* x, const float* cls_pos0, int n_frames, int tokens, int dim, cudaStream_t s) {
  const int total = n_frames * dim;
  VC_LAUNCH("cls_rows", total * 8.0, s, (cls_rows_kernel<<<(total + 255) / 256, 256, 0, s>>>(x, cls_pos0, n_frames, tokens, dim)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =========================================================================== ViT attention
// One CTA per (frame, head), head_dim 64, all S keys resident in shared memory (bf16,
// 128-byte rows, XOR-swizzled 16-byte chunks).  Each warp owns 16-query tiles and runs
// the flash-style loop over 64-key blocks with mma.sync m16n8k16 (bf16 in, fp32 acc),
// fp32 online softmax.  ~4 % of the encoder FLOPs; the GEMMs carry
// the tcgen05 path.  Reference: nn.MultiheadAttention / timm Attention -> SDPA, scale
// head_dim^-0.5, no mask (src/models/video_encoder.py:112-121, 266-286).
constexpr int HD = 64;

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int ATT_THREADS = 256;   // 8 warps: 13 query tiles of ViT-B/16 spread 2,2,2,2,2,1,1,1

__device__ __forceinline__ void cp_async_16_zfill(uint32_t smem_dst, const void* gsrc, bool valid) {
  const int src_bytes = valid ? 16 : 0;   // src-size 0 -> the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 2) vit_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                                      int tokens, int heads, int s_pad) {
  extern __shared__ __align__(128) uint8_t smem_att[];
  uint8_t* sQ = smem_att;
  uint8_t* sK = sQ + s_pad * 128;
  uint8_t* sV = sK + s_pad * 128;
  const int frame = blockIdx.x / heads, head = blockIdx.x - frame * heads;
  const int D = heads * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* base = qkv + static_cast<long long>(frame) * tokens * 3 * D + head * HD;

  // global -> shared with cp.async (all chunks in flight at once): 8 x 16-byte chunks per row per
  // matrix; rows >= tokens are zero-filled so padded keys/values stay finite
  {
    const uint32_t s0 = smem_u32(smem_att);
    const int per = s_pad * 8;
    for (int i = tid; i < per * 3; i += ATT_THREADS) {
      const int which = i / per;
      const int rem = i - which * per;
      const int row = rem >> 3, chunk = rem & 7;
      const bool ok = row < tokens;
      const __nv_bfloat16* src = base + static_cast<long long>(ok ? row : 0) * 3 * D + which * D + chunk * 8;
      cp_async_16_zfill(s0 + which * s_pad * 128 + swz(row, chunk), src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  const uint32_t q_base = smem_u32(sQ), k_base = smem_u32(sK), v_base = smem_u32(sV);
  const int g = lane >> 2, t4 = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  const float scale_log2 = 0.125f * 1.4426950408889634f;   // head_dim^-0.5 * log2(e)
  const int m_tiles = (tokens + 15) >> 4;
  const int k_blocks = (tokens + 63) >> 6;

  for (int mt = warp; mt < m_tiles; mt += ATT_THREADS / 32) {
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int row = mt * 16 + (mi & 1) * 8 + r8;
      ldmatrix_x4(q_base + swz(row, ks * 2 + (mi >> 1)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    for (int kb = 0; kb < k_blocks; ++kb) {
      float sc[8][4];
      const int key0 = kb * 64;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
        if (key0 + nt * 8 < tokens) {   // warp-uniform
          const int key = key0 + nt * 8 + r8;
          uint32_t b[8];
          ldmatrix_x4(k_base + swz(key, mi), b[0], b[1], b[2], b[3]);         // d chunks 0..3 -> k-steps 0,1
          ldmatrix_x4(k_base + swz(key, 4 + mi), b[4], b[5], b[6], b[7]);     // d chunks 4..7 -> k-steps 2,3
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_bf16_16816(sc[nt], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b[2 * ks], b[2 * ks + 1]);
        }
      }
      // mask + block row max (rows g and g+8)
      float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int kcol = key0 + nt * 8 + 2 * t4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = (kcol + (e & 1)) < tokens;
          sc[nt][e] = ok ? sc[nt][e] * scale_log2 : -INFINITY;
          bm[e >> 1] = fmaxf(bm[e >> 1], sc[nt][e]);
        }
      }
      float corr[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        bm[h] = fmaxf(bm[h], __shfl_xor_sync(0xffffffffu, bm[h], 1));
        bm[h] = fmaxf(bm[h], __shfl_xor_sync(0xffffffffu, bm[h], 2));
        const float m_new = fmaxf(m_run[h], bm[h]);   // finite: every block holds >= 1 valid key
        corr[h] = exp2f(m_run[h] - m_new);
        m_run[h] = m_new;
      }
      float bs[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = exp2f(sc[nt][e] - m_run[e >> 1]);
          sc[nt][e] = pv;
          bs[e >> 1] += pv;
        }
#pragma unroll
      for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + bs[h];   // quad-partial sums, reduced at the end
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        o[nd][0] *= corr[0]; o[nd][1] *= corr[0];
        o[nd][2] *= corr[1]; o[nd][3] *= corr[1];
      }
      // O += P V
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (key0 + j * 16 < tokens) {   // warp-uniform
          const uint32_t a0 = pack_bf16(sc[2 * j][0], sc[2 * j][1]);
          const uint32_t a1 = pack_bf16(sc[2 * j][2], sc[2 * j][3]);
          const uint32_t a2 = pack_bf16(sc[2 * j + 1][0], sc[2 * j + 1][1]);
          const uint32_t a3 = pack_bf16(sc[2 * j + 1][2], sc[2 * j + 1][3]);
          const int key = key0 + j * 16 + (mi & 1) * 8 + r8;
#pragma unroll
          for (int ndp = 0; ndp < 4; ++ndp) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(v_base + swz(key, ndp * 2 + (mi >> 1)), b0, b1, b2, b3);
            mma_bf16_16816(o[2 * ndp], a0, a1, a2, a3, b0, b1);
            mma_bf16_16816(o[2 * ndp + 1], a0, a1, a2, a3, b2, b3);
          }
        }
      }
    }
    // finalise: full row sums across the quad, normalise, store
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
      l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    const int row0 = mt * 16 + g, row1 = row0 + 8;
    __nv_bfloat16* ob = out + static_cast<long long>(frame) * tokens * D + head * HD;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      const int col = nd * 8 + 2 * t4;
      if (row0 < tokens) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(row0) * D + col) = pack_bf16(o[nd][0] * inv0, o[nd][1] * inv0);
      if (row1 < tokens) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(row1) * D + col) = pack_bf16(o[nd][2] * inv1, o[nd][3] * inv1);
    }
  }
}

int vit_attention_mma_sync(const void* qkv, void* out, int n_frames, int tokens, int heads, int head_dim, cudaStream_t s);

int vit_attention(const void* qkv, void* out, int n_frames, int tokens, int heads, int head_dim, cudaStream_t s) {
  VC_REQUIRE(head_dim == HD, "vit_attention: head_dim=%d (only 64 is built)", head_dim);
  VC_REQUIRE(tokens > 0 && tokens <= 576, "vit_attention: tokens=%d out of range", tokens);
  if (n_frames <= 0) return 0;
  // tokens <= 264 (ViT-B/16: 197, ViT-L/14: 257): the tcgen05 kernel; longer sequences stay on the mma.sync kernel below
  static const bool legacy = getenv("VC_VIT_ATTENTION_MMA_SYNC") != nullptr;
  if (!legacy && vit_attention_tc_supported(tokens, heads, head_dim) && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0)
    return vit_attention_tc(qkv, out, n_frames, tokens, heads, s);
  return vit_attention_mma_sync(qkv, out, n_frames, tokens, heads, head_dim, s);
}

int vit_attention_mma_sync(const void* qkv, void* out, int n_frames, int tokens, int heads, int head_dim, cudaStream_t s) {
  VC_REQUIRE(head_dim == HD, "vit_attention: head_dim=%d (only 64 is built)", head_dim);
  VC_REQUIRE(tokens > 0 && tokens <= 576, "vit_attention: tokens=%d out of range", tokens);
  if (n_frames <= 0) return 0;
  const int s_pad = ((tokens + 63) / 64) * 64;
  const int smem = 3 * s_pad * 128;
  static PerDeviceOnce attr;                    // opt in to the largest size this kernel ever needs (576 tokens), once per device
  if (attr.first()) VC_CUDA_OK(cudaFuncSetAttribute(vit_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 576 * 128));
  const double flops = 4.0 * n_frames * heads * static_cast<double>(tokens) * tokens * HD;
  VC_LAUNCH("vit_attention", flops, s,
            (vit_attention_kernel<<<n_frames * heads, ATT_THREADS, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                                     static_cast<__nv_bfloat16*>(out), tokens, heads, s_pad)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =========================================================================== class-token attention (last block)
// Only the class token of the LAST block is consumed downstream (video_encoder.py:256-258), so in that block the
// attention output is needed for one query per frame.  One warp per (frame, head): lanes split the keys for the
// scores (one 128-byte K row per lane), then split head_dim for the weighted V sum (coalesced 128-byte rows).
constexpr int CLS_MAX_TOKENS = 640;
// Generalised to any query row (row0 .. row0+n_rows-1 of every frame), with a compact ([n_frames*n_rows, D]) or an in-place
// ([n_frames*tokens, D]) output: also used for the query rows the tcgen05 kernel leaves out when tokens > 256.
__global__ void __launch_bounds__(128) vit_cls_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                                int n_items, int tokens, int heads, int row0, int n_rows, int compact) {
  __shared__ float s_p[4][CLS_MAX_TOKENS];
  __shared__ float s_q[4][HD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;
  if (item >= n_items) return;
  const int fr = item / heads, head = item - fr * heads;          // fr = frame * n_rows + row index
  const int frame = fr / n_rows, qrow = row0 + (fr - frame * n_rows);
  const int D = heads * HD;
  const __nv_bfloat16* base = qkv + static_cast<long long>(frame) * tokens * 3 * D + head * HD;
  {
    const float2 q2 = unpack_bf16(*reinterpret_cast<const uint32_t*>(base + static_cast<long long>(qrow) * 3 * D + 2 * lane));
    s_q[warp][2 * lane] = q2.x; s_q[warp][2 * lane + 1] = q2.y;
  }
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < tokens; j += 32) {
    const uint4* kp = reinterpret_cast<const uint4*>(base + static_cast<long long>(j) * 3 * D + D);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = __ldg(kp + c);
      const float4 qa = *reinterpret_cast<const float4*>(&s_q[warp][c * 8]), qb = *reinterpret_cast<const float4*>(&s_q[warp][c * 8 + 4]);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
      acc = fmaf(qa.x, a.x, acc); acc = fmaf(qa.y, a.y, acc); acc = fmaf(qa.z, b.x, acc); acc = fmaf(qa.w, b.y, acc);
      acc = fmaf(qb.x, cc.x, acc); acc = fmaf(qb.y, cc.y, acc); acc = fmaf(qb.z, d.x, acc); acc = fmaf(qb.w, d.y, acc);
    }
    acc *= 0.125f;
    s_p[warp][j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < tokens; j += 32) {
    const float e = __expf(s_p[warp][j] - mx);
    s_p[warp][j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  // weighted V sum: lane = (key group kg = lane >> 3, 16-byte chunk c = lane & 7): a warp instruction covers four whole V rows
  // (16-byte loads, 16 of them in flight per lane), each lane accumulates its 8 head dims over keys kg, kg+4, ...; the four
  // key groups are folded with two shuffles at the end
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int kg = lane >> 3, c8 = lane & 7;
  const __nv_bfloat16* vbase = base + 2 * D + c8 * 8;
  for (int j0 = 0; j0 < tokens; j0 += 64) {
    uint4 u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = j0 + 4 * i + kg;
      if (j < tokens) u[i] = __ldg(reinterpret_cast<const uint4*>(vbase + static_cast<long long>(j) * 3 * D));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = j0 + 4 * i + kg;
      if (j < tokens) {
        const float pj = s_p[warp][j];
        const float2 a = unpack_bf16(u[i].x), b = unpack_bf16(u[i].y), cc = unpack_bf16(u[i].z), d = unpack_bf16(u[i].w);
        o[0] = fmaf(pj, a.x, o[0]); o[1] = fmaf(pj, a.y, o[1]); o[2] = fmaf(pj, b.x, o[2]); o[3] = fmaf(pj, b.y, o[3]);
        o[4] = fmaf(pj, cc.x, o[4]); o[5] = fmaf(pj, cc.y, o[5]); o[6] = fmaf(pj, d.x, o[6]); o[7] = fmaf(pj, d.y, o[7]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
  }
  const float inv = 1.f / sum;
  const long long orow = compact ? fr : static_cast<long long>(frame) * tokens + qrow;
  if (kg == 0) {
    uint4 w;
    w.x = pack_bf16(o[0] * inv, o[1] * inv); w.y = pack_bf16(o[2] * inv, o[3] * inv);
    w.z = pack_bf16(o[4] * inv, o[5] * inv); w.w = pack_bf16(o[6] * inv, o[7] * inv);
    *(reinterpret_cast<uint4*>(out + orow * D + head * HD) + c8) = w;
  }
}

int vit_row_attention(const void* qkv, void* out, int n_frames, int tokens, int heads, int row0, int n_rows, cudaStream_t s) {
  VC_REQUIRE(tokens > 0 && tokens <= CLS_MAX_TOKENS && row0 >= 0 && n_rows > 0 && row0 + n_rows <= tokens, "vit_row_attention: bad rows");
  if (n_frames <= 0) return 0;
  const int items = n_frames * n_rows * heads;
  VC_LAUNCH("vit_row_attention", 4.0 * items * static_cast<double>(tokens) * HD, s,
            (vit_cls_attention_kernel<<<(items + 3) / 4, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), items,
                                                                    tokens, heads, row0, n_rows, 0)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

int vit_cls_attention(const void* qkv, void* out, int n_frames, int tokens, int heads, int head_dim, cudaStream_t s) {
  VC_REQUIRE(head_dim == HD, "vit_cls_attention: head_dim=%d (only 64 is built)", head_dim);
  VC_REQUIRE(tokens > 0 && tokens <= CLS_MAX_TOKENS, "vit_cls_attention: tokens=%d out of range", tokens);
  if (n_frames <= 0) return 0;
  const int items = n_frames * heads;
  VC_LAUNCH("vit_cls_attention", 4.0 * items * static_cast<double>(tokens) * HD, s,
            (vit_cls_attention_kernel<<<(items + 3) / 4, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), items,
                                                                    tokens, heads, 0, 1, 1)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =========================================================================== pool + proj + prefix
// One CTA per video: class-token temporal mean (video_encoder.py:256-258) -> encoder.proj
// Linear(dim, video_dim) (:316) -> F.layer_norm(no affine) * ln_scale, * in_weight
// (core/engine.py:45-50) -> decoder.mapper Linear(video_dim, P*H) (text_decoder.py:36-45,69).
__global__ void __launch_bounds__(1024) pool_prefix_kernel(const float* __restrict__ cls, int T, int dim,
                                                          const float* __restrict__ head_w, const float* __restrict__ head_b,
                                                          int video_dim, float ln_scale, float in_weight,
                                                          const float* __restrict__ mapper_w, const float* __restrict__ mapper_b,
                                                          int mapper_out, float* __restrict__ feat_out,
                                                          float* __restrict__ prefix_out) {
  extern __shared__ float s_pool[];          // [dim] pooled, then [video_dim] feat / emb
  float* s_feat = s_pool + dim;
  __shared__ float s_red[2];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  for (int c = tid; c < dim; c += blockDim.x) {
    float acc = 0.f;
#pragma unroll 8
    for (int t = 0; t < T; ++t) acc += cls[(static_cast<long long>(b) * T + t) * dim + c];     // same order, loads batched
    s_pool[c] = acc / static_cast<float>(T);
  }
  __syncthreads();
  // four output rows per warp iteration: their weight loads are independent, so one memory round trip serves four outputs
  // (each output keeps its own lane-strided accumulation order)
  for (int o0 = warp * 4; o0 < video_dim; o0 += nwarp * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < dim; c += 32) {
      const float x = s_pool[c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (o0 + i < video_dim) acc[i] = fmaf(x, __ldg(head_w + static_cast<long long>(o0 + i) * dim + c), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = warp_sum(acc[i]);
      if (lane == 0 && o0 + i < video_dim) {
        const float f = a + head_b[o0 + i];
        s_feat[o0 + i] = f;
        feat_out[static_cast<long long>(b) * video_dim + o0 + i] = f;
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    float sum = 0.f;
    for (int c = lane; c < video_dim; c += 32) sum += s_feat[c];
    const float mean = warp_sum(sum) / static_cast<float>(video_dim);
    float sq = 0.f;
    for (int c = lane; c < video_dim; c += 32) { const float d = s_feat[c] - mean; sq += d * d; }
    const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(video_dim) + 1e-5f);
    if (lane == 0) { s_red[0] = mean; s_red[1] = rstd; }
  }
  __syncthreads();
  const bool do_ln = ln_scale > 0.f, do_w = in_weight > 0.f;
  for (int c = tid; c < video_dim; c += blockDim.x) {
    float e = s_feat[c];
    if (do_ln) e = (e - s_red[0]) * s_red[1] * ln_scale;
    if (do_w) e = e * in_weight;
    s_pool[c] = e;   // reuse as emb (video_dim <= dim)
  }
  __syncthreads();
  for (int o0 = warp * 4; o0 < mapper_out; o0 += nwarp * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < video_dim; c += 32) {
      const float x = s_pool[c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (o0 + i < mapper_out) acc[i] = fmaf(x, __ldg(mapper_w + static_cast<long long>(o0 + i) * video_dim + c), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = warp_sum(acc[i]);
      if (lane == 0 && o0 + i < mapper_out) prefix_out[static_cast<long long>(b) * mapper_out + o0 + i] = a + mapper_b[o0 + i];
    }
  }
}

int pool_prefix(const float* cls, int B, int T, int dim, const float* head_w, const float* head_b, int video_dim, float ln_scale,
                float in_weight, const float* mapper_w, const float* mapper_b, int mapper_out, float* feat_out, float* prefix_out,
                cudaStream_t s) {
  VC_REQUIRE(T > 0 && dim > 0 && video_dim > 0 && video_dim <= dim, "pool_prefix: bad dims T=%d dim=%d video_dim=%d", T, dim, video_dim);
  if (B <= 0) return 0;
  const int smem = (dim + video_dim) * sizeof(float);
  VC_LAUNCH("pool_prefix", static_cast<double>(B) * (T * dim + (dim + mapper_out) * video_dim) * 4.0, s,
            (pool_prefix_kernel<<<B, 1024, smem, s>>>(cls, T, dim, head_w, head_b, video_dim, ln_scale, in_weight, mapper_w,
                                                     mapper_b, mapper_out, feat_out, prefix_out)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// vit_fused_pool_temporal (core/operators/cupy_vit_pool.py:127-186): feat [bsz*T, tokens, C] -> [bsz, C], the temporal mean of the
// class token (cls) or of the mean over the patch tokens (gap).  One CTA per (video, 32-channel group): lane = channel, so every
// row a warp touches is one 128-byte (fp32) / 64-byte (bf16) segment; the T x rows-per-frame rows are dealt round-robin to the
// 8 warps with four loads in flight each, and the warps' partial sums are added in warp order (fixed summation order).
template <typename T_in>
__global__ void __launch_bounds__(256) vit_pool_kernel(const T_in* __restrict__ x, float* __restrict__ y, int bsz, int T, int tokens,
                                                       int C, int gap) {
  __shared__ float s_part[8][32];
  const int groups = (C + 31) / 32;
  const int b = blockIdx.x / groups, c = (blockIdx.x - b * groups) * 32 + (threadIdx.x & 31);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = gap ? 1 : 0, per = gap ? tokens - 1 : 1, rows = T * per;
  float acc = 0.f;
  if (c < C) {
    for (int r0 = warp; r0 < rows; r0 += 32) {          // rows r0, r0 + 8, r0 + 16, r0 + 24 of this warp
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + 8 * u;
        v[u] = 0.f;
        if (r < rows) {
          const int t = r / per, p = p0 + (r - t * per);
          v[u] = static_cast<float>(x[((static_cast<long long>(b) * T + t) * tokens + p) * C + c]);
        }
      }
      acc += (v[0] + v[1]) + (v[2] + v[3]);
    }
  }
  s_part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < C) {
    float s = s_part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) s += s_part[w][lane];
    y[static_cast<long long>(b) * C + c] = s / static_cast<float>(rows);
  }
}
int vit_pool_temporal(const void* feat, int is_bf16, int bsz, int T, int tokens, int C, int gap, float* out, cudaStream_t s) {
  VC_REQUIRE(bsz >= 0 && T > 0 && tokens > (gap ? 1 : 0) && C > 0, "vit_pool: bad shape");
  if (bsz == 0) return 0;
  const int grid = bsz * ((C + 31) / 32);
  if (is_bf16)
    VC_LAUNCH("vit_pool", 0.0, s, (vit_pool_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(feat), out, bsz, T, tokens, C, gap)));
  else
    VC_LAUNCH("vit_pool", 0.0, s, (vit_pool_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(feat), out, bsz, T, tokens, C, gap)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

// cupy_linear_mapper.py:14-40 semantics (y = b + x w^T), but one warp per output so the
// weight row is read coalesced (the reference kernel strides threads by in_features).
__global__ void __launch_bounds__(256) linear_bias_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ y, int rows, int in_f,
                                                          int out_f) {
  const long long gw = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= static_cast<long long>(rows) * out_f) return;
  const int r = static_cast<int>(gw / out_f), o = static_cast<int>(gw - static_cast<long long>(r) * out_f);
  float acc = 0.f;
  for (int k = lane; k < in_f; k += 32) acc = fmaf(x[static_cast<long long>(r) * in_f + k], __ldg(w + static_cast<long long>(o) * in_f + k), acc);
  acc = warp_sum(acc);
  if (lane == 0) y[gw] = acc + (b ? b[o] : 0.f);
}
int linear_bias_f32(const float* x, const float* w, const float* b, float* y, int rows, int in_f, int out_f, cudaStream_t s) {
  VC_REQUIRE(rows >= 0 && in_f > 0 && out_f > 0, "linear_bias: bad shape");
  if (rows == 0) return 0;
  const long long warps = static_cast<long long>(rows) * out_f;
  VC_LAUNCH("linear_bias", 0.0, s, (linear_bias_kernel<<<static_cast<int>((warps + 7) / 8), 256, 0, s>>>(x, w, b, y, rows, in_f, out_f)));
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vc
