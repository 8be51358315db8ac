// Pillow-exact antialiased bilinear resize of uint8 RGB frames (the transform the reference applies to every frame before
// ToTensor/Normalize: core/preprocessing/frame_loader.py:34-45 -> torchvision Resize on a PIL image -> Image.resize(BILINEAR)
// -> Pillow ImagingResample, 8 bits per channel).  Pillow's rule, restated:
//   * separable, horizontal pass first, the intermediate image is uint8;
//   * per output index a window [xmin, xmin+count) of the input and fixed-point weights with 22 fractional bits
//     (PRECISION_BITS = 32 - 8 - 2), computed in double precision on the host (resample.py) exactly as Pillow does;
//   * pixel = clip8((2^21 + sum_i in[xmin+i] * w[i]) >> 22), 32-bit integer arithmetic.
// Byte-exact by construction; HBM-bound byte work: rows are staged through shared memory so that every global access is
// a full 16-byte vector, weights come from the read-only cache.
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

#define VC_LAUNCH(name, work, stream, ...)        \
  do {                                            \
    vc::KernelScope _ks(name, work, stream);      \
    __VA_ARGS__;                                  \
  } while (0)

namespace {

constexpr int RS_PREC = 22;
constexpr int RS_ROWS = 8;            // input rows per CTA in the horizontal pass
constexpr int RS_VROWS = 8;           // output rows per CTA in the vertical pass (consecutive rows share most of their taps: L1 hits)
constexpr int RS_KREG = 8;            // taps kept in registers per output column (scale <= 3.5); longer windows re-read the table

__device__ __forceinline__ uint32_t clip8(int v) {
  v >>= RS_PREC;
  return static_cast<uint32_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src [rows_total, W, 3] -> dst [rows_total, OW, 3]; one CTA per RS_ROWS rows.  The rows of a CTA are one
// contiguous byte range in global memory on both sides: it is moved as aligned 16-byte vectors (byte accesses only for the
// unaligned head and tail), the shared-memory copy is shifted so that vector k of global memory is vector k of shared memory.
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long rows_total, int W,
                                                       int OW, const int32_t* __restrict__ kk, const int32_t* __restrict__ bounds, int ksize) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  const int in_bytes = W * 3, out_bytes = OW * 3;
  const long long row0 = static_cast<long long>(blockIdx.x) * RS_ROWS;
  const int n_rows = static_cast<int>(rows_total - row0 < RS_ROWS ? rows_total - row0 : RS_ROWS);
  const uint8_t* gin = src + row0 * in_bytes;
  uint8_t* gout = dst + row0 * out_bytes;
  const int in_total = n_rows * in_bytes, out_total = n_rows * out_bytes;
  const int in_head = static_cast<int>((16 - (reinterpret_cast<uintptr_t>(gin) & 15)) & 15);     // bytes before the first aligned vector
  const int out_head = static_cast<int>((16 - (reinterpret_cast<uintptr_t>(gout) & 15)) & 15);
  uint8_t* s_in = rs_smem + ((16 - in_head) & 15);                                                // s_in + in_head is 16-byte aligned
  uint8_t* s_out = rs_smem + ((RS_ROWS * in_bytes + 47) & ~15) + ((16 - out_head) & 15);
  {
    const int body = in_total > in_head ? (in_total - in_head) >> 4 : 0;                          // whole vectors
    for (int v = threadIdx.x; v < body; v += blockDim.x)
      *reinterpret_cast<uint4*>(s_in + in_head + 16 * v) = __ldg(reinterpret_cast<const uint4*>(gin + in_head) + v);
    const int tail0 = in_head + 16 * body;
    for (int i = threadIdx.x; i < in_head && i < in_total; i += blockDim.x) s_in[i] = __ldg(gin + i);
    for (int i = tail0 + threadIdx.x; i < in_total; i += blockDim.x) s_in[i] = __ldg(gin + i);
  }
  __syncthreads();
  // one output column per thread (its window and weights live in registers), all rows of the CTA
  for (int xo = threadIdx.x; xo < OW; xo += blockDim.x) {
    const int xmin = __ldg(bounds + 2 * xo), cnt = __ldg(bounds + 2 * xo + 1);
    const int32_t* k = kk + static_cast<size_t>(xo) * ksize;
    if (cnt <= RS_KREG) {
      int w[RS_KREG];
#pragma unroll
      for (int t = 0; t < RS_KREG; ++t) w[t] = t < cnt ? __ldg(k + t) : 0;
      for (int r = 0; r < n_rows; ++r) {
        // the window's 24 bytes as seven aligned 32-bit words, realigned with funnel shifts: a third of the shared-memory
        // instructions of byte loads (bytes past the window meet zero weights; the buffer has slack behind the last row)
        const uint32_t a = smem_u32(s_in + r * in_bytes + xmin * 3);
        const uint32_t base = a & ~3u, sh = (a & 3u) * 8u;
        uint32_t u[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u[j]) : "r"(base + 4 * j));
        int a0 = 1 << (RS_PREC - 1), a1 = a0, a2 = a0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const uint32_t v = __funnelshift_r(u[j], u[j + 1], sh);       // bytes 4j .. 4j+3 of the window
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int i = 4 * j + b, t = i / 3, c = i - 3 * t;           // compile-time after unrolling
            const int px = static_cast<int>((v >> (8 * b)) & 255u);
            if (c == 0) a0 += px * w[t]; else if (c == 1) a1 += px * w[t]; else a2 += px * w[t];
          }
        }
        uint8_t* o = s_out + r * out_bytes + xo * 3;
        o[0] = static_cast<uint8_t>(clip8(a0)); o[1] = static_cast<uint8_t>(clip8(a1)); o[2] = static_cast<uint8_t>(clip8(a2));
      }
    } else {
      for (int r = 0; r < n_rows; ++r) {
        const uint8_t* p = s_in + r * in_bytes + xmin * 3;
        int a0 = 1 << (RS_PREC - 1), a1 = a0, a2 = a0;
        for (int t = 0; t < cnt; ++t) {
          const int w = __ldg(k + t);
          a0 += p[3 * t] * w; a1 += p[3 * t + 1] * w; a2 += p[3 * t + 2] * w;
        }
        uint8_t* o = s_out + r * out_bytes + xo * 3;
        o[0] = static_cast<uint8_t>(clip8(a0)); o[1] = static_cast<uint8_t>(clip8(a1)); o[2] = static_cast<uint8_t>(clip8(a2));
      }
    }
  }
  __syncthreads();
  {
    const int body = out_total > out_head ? (out_total - out_head) >> 4 : 0;
    for (int v = threadIdx.x; v < body; v += blockDim.x)
      reinterpret_cast<uint4*>(gout + out_head)[v] = *reinterpret_cast<const uint4*>(s_out + out_head + 16 * v);
    const int tail0 = out_head + 16 * body;
    for (int i = threadIdx.x; i < out_head && i < out_total; i += blockDim.x) gout[i] = s_out[i];
    for (int i = tail0 + threadIdx.x; i < out_total; i += blockDim.x) gout[i] = s_out[i];
  }
}

// vertical pass: src [n, H, row_bytes] -> dst [n, OH, row_bytes]; grid = (ceil(row_words / 256), ceil(OH / 8), n): the output rows and
// their taps are uniform per CTA (no index arithmetic per thread), one thread per 4 consecutive bytes, taps read as coalesced words
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int OH, int row_words,
                                                       const int32_t* __restrict__ kk, const int32_t* __restrict__ bounds, int ksize) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= row_words) return;
  const int f = blockIdx.z;
  for (int yo = blockIdx.y * RS_VROWS; yo < OH && yo < (blockIdx.y + 1) * RS_VROWS; ++yo) {
    const int ymin = __ldg(bounds + 2 * yo), cnt = __ldg(bounds + 2 * yo + 1);
    const int32_t* k = kk + static_cast<size_t>(yo) * ksize;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(src) + (static_cast<long long>(f) * H + ymin) * row_words + w;
    int a0 = 1 << (RS_PREC - 1), a1 = a0, a2 = a0, a3 = a0;
    if (cnt <= RS_KREG) {
      uint32_t u[RS_KREG];
      int wt[RS_KREG];
#pragma unroll
      for (int i = 0; i < RS_KREG; ++i) {                  // every tap row requested before the first one is used
        wt[i] = i < cnt ? __ldg(k + i) : 0;
        u[i] = i < cnt ? __ldg(p + static_cast<long long>(i) * row_words) : 0u;
      }
#pragma unroll
      for (int i = 0; i < RS_KREG; ++i) {
        a0 += static_cast<int>(u[i] & 255u) * wt[i]; a1 += static_cast<int>((u[i] >> 8) & 255u) * wt[i];
        a2 += static_cast<int>((u[i] >> 16) & 255u) * wt[i]; a3 += static_cast<int>(u[i] >> 24) * wt[i];
      }
    } else {
      for (int i = 0; i < cnt; ++i) {
        const uint32_t u = __ldg(p + static_cast<long long>(i) * row_words);
        const int wgt = __ldg(k + i);
        a0 += static_cast<int>(u & 255u) * wgt; a1 += static_cast<int>((u >> 8) & 255u) * wgt;
        a2 += static_cast<int>((u >> 16) & 255u) * wgt; a3 += static_cast<int>(u >> 24) * wgt;
      }
    }
    reinterpret_cast<uint32_t*>(dst)[(static_cast<long long>(f) * OH + yo) * row_words + w] = clip8(a0) | (clip8(a1) << 8) | (clip8(a2) << 16) | (clip8(a3) << 24);
  }
}

}  // namespace

int resize_bilinear_u8(const uint8_t* src, int n, int H, int W, uint8_t* tmp, uint8_t* dst, int OH, int OW, const int32_t* kx, const int32_t* bx,
                       int ksize_x, const int32_t* ky, const int32_t* by, int ksize_y, cudaStream_t s) {
  VC_REQUIRE(n >= 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "resize: n=%d %dx%d -> %dx%d", n, H, W, OH, OW);
  VC_REQUIRE((OW * 3) % 4 == 0, "resize: output row of %d bytes is not a multiple of 4", OW * 3);
  VC_REQUIRE(ksize_x > 0 && ksize_y > 0 && kx && bx && ky && by, "resize: coefficient tables missing");
  if (n == 0) return 0;
  const uint8_t* vsrc = src;
  int vH = H;
  if (W != OW) {
    VC_REQUIRE(tmp != nullptr || H == OH, "resize: the horizontal pass needs a [n,%d,%d,3] scratch buffer", H, OW);
    uint8_t* hdst = (H == OH) ? dst : tmp;
    const long long rows = static_cast<long long>(n) * H;
    const size_t smem = static_cast<size_t>(RS_ROWS) * (W * 3 + OW * 3) + 96;
    VC_REQUIRE(smem <= 200 * 1024, "resize: rows of %d pixels do not fit shared memory", W);
    static size_t attr = 48 * 1024;
    if (smem > attr) {
      VC_CUDA_OK(cudaFuncSetAttribute(resize_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = 200 * 1024;
    }
    VC_LAUNCH("resize_h", static_cast<double>(rows) * (W + OW) * 3.0, s,
              (resize_h_kernel<<<static_cast<unsigned>((rows + RS_ROWS - 1) / RS_ROWS), 256, smem, s>>>(src, hdst, rows, W, OW, kx, bx, ksize_x)));
    VC_CUDA_OK(cudaGetLastError());
    vsrc = hdst;
  }
  if (vH != OH) {
    const int row_words = OW * 3 / 4;
    VC_REQUIRE(OH <= 65535 && n <= 65535, "resize: grid limits (OH=%d, n=%d)", OH, n);
    const int bt = row_words < 256 ? ((row_words + 31) / 32) * 32 : 256;
    VC_LAUNCH("resize_v", static_cast<double>(n) * (vH + OH) * OW * 3.0, s,
              (resize_v_kernel<<<dim3((row_words + bt - 1) / bt, (OH + RS_VROWS - 1) / RS_VROWS, n), bt, 0, s>>>(vsrc, dst, vH, OH, row_words, ky, by, ksize_y)));
    VC_CUDA_OK(cudaGetLastError());
  } else if (W == OW && dst != src) {
    VC_CUDA_OK(cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * H * W * 3, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

}  // namespace vc
