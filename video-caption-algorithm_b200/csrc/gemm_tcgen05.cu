// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epi(A[M,K] * W[N,K]^T)
//
//   * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a 4-deep
//     mbarrier ring, 128x256x64 tiles;
//   * one elected thread issues tcgen05.mma (cta_group::1, M=128, N=256, K=16),
//     fp32 accumulators live in TMEM, double-buffered (2 x 256 columns) so the
//     epilogue of tile i overlaps the MMAs of tile i+1;
//   * 8 epilogue warps read TMEM with tcgen05.ld (one accumulator row per thread)
//     and fuse bias / GELU / residual-add / patch-embed scatter before storing.
//
// This is the dense contraction of the ViT frame encoder (reference:
// src/models/video_encoder.py:288-326 -> torchvision/timm Linear + Conv2d patch
// embed) and of the GPT-2 prefill (transformers Conv1D).  nn.Linear weights are
// [out,in] = [N,K] K-major, which is exactly the B operand layout tcgen05 wants.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <cudaTypedefs.h>
#include <mutex>
#include <unordered_map>

namespace vc {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, UK = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TMEM_COLS = 512;         // 2 accumulators x 256 fp32 columns
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct GemmParams {
  int M, N, K;
  int mode;
  const float* bias;      // [N] or null
  void* out;              // bf16 [M,ldo] or fp32 [rows,ldo]
  int ldo;                // leading dimension of out (elements)
  const float* aux;       // patch-embed: pos embedding [tokens, N]
  int rows_per_group;     // patch-embed: patches per frame (196); out row = g*(rpg+1)+1+r
  uint64_t desc_hi;       // smem descriptor without the start address (see umma_desc_sw128)
  uint32_t k_adv;         // start-address step per K=16 slice, in 16-byte units
  uint32_t idesc;         // tcgen05 instruction descriptor
};

template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int col0, bool row_ok, const uint32_t (&acc)[32]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  if (MODE == VC_EPI_BIAS_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (MODE == VC_EPI_BIAS_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
  }
  if (!row_ok) return;
  if (MODE == VC_EPI_BIAS || MODE == VC_EPI_BIAS_GELU_ERF || MODE == VC_EPI_BIAS_GELU_TANH) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 w;
      w.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
      w.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
      w.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
      w.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
      o[j] = w;
    }
  } else if (MODE == VC_EPI_BIAS_RESID_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 r = o[j];
      r.x += v[4 * j + 0]; r.y += v[4 * j + 1]; r.z += v[4 * j + 2]; r.w += v[4 * j + 3];
      o[j] = r;
    }
  } else if (MODE == VC_EPI_BIAS_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else if (MODE == VC_EPI_PATCH_EMBED) {
    const int g = row / p.rows_per_group, r = row - g * p.rows_per_group;
    const size_t orow = static_cast<size_t>(g) * (p.rows_per_group + 1) + 1 + r;
    const float4* pe = reinterpret_cast<const float4*>(p.aux + static_cast<size_t>(1 + r) * p.N + col0);
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 e = __ldg(pe + j);
      o[j] = make_float4(v[4 * j + 0] + e.x, v[4 * j + 1] + e.y, v[4 * j + 2] + e.z, v[4 * j + 3] + e.w);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must sit on 1024-byte boundaries
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int k_blocks = p.K / BK;
  const int total = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b); }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (one lane)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(&tm_a, &full_bar[stage], sa, kb * BK, m0);
          tma_load_2d(&tm_b, &full_bar[stage], sa + A_BYTES, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (one lane)
    if (lane == 0) {
      const uint32_t idesc = p.idesc;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t da = p.desc_hi | static_cast<uint64_t>((a_addr & 0x3FFFFu) >> 4);
          const uint64_t db = p.desc_hi | static_cast<uint64_t>(((a_addr + A_BYTES) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // +32 B per K=16 slice inside the 128 B swizzle row (start address is in 16 B units)
            tc_mma_bf16(d_tmem, da + p.k_adv * k, db + p.k_adv * k, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);            // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);                // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps
    const int ew = warp - 2;            // 0..7
    const int quarter = warp & 3;       // TMEM lane quarter this warp may read
    const int half = ew >> 2;           // which 128 columns of the 256-wide tile
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col_in_tile = half * 128 + c * 32;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col_in_tile, r);
        tmem_ld_wait();
        const int col0 = n0 + col_in_tile;
        if (col0 < p.N) epilogue_chunk<MODE>(p, row, col0, row_ok, r);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

int get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  return g_encode ? 0 : -1;
}

// K-major bf16 [rows, K] matrix, box = [BK, box_rows], 128-byte swizzle.
int make_map(CUtensorMap* tm, const void* ptr, int rows, int K, int box_rows) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ptr=%p", static_cast<int>(r), rows, K, ptr);
    return -3;
  }
  return 0;
}

int g_num_sms = 0;
// bring-up overrides (vc_debug_gemm_override); 0 = built-in encoding
uint64_t g_dbg_desc_hi = 0;
uint32_t g_dbg_k_adv = 0, g_dbg_idesc = 0;
std::mutex g_cfg_mu;
bool g_attr_set[8] = {false};

template <int MODE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, cudaStream_t stream) {
  {
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    if (!g_attr_set[MODE]) {
      VC_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      g_attr_set[MODE] = true;
    }
  }
  static const char* const kNames[] = {"gemm_bias", "gemm_gelu_erf", "gemm_gelu_tanh", "gemm_resid", "gemm_f32", "gemm_patch"};
  {
    KernelScope ks(kNames[MODE], 2.0 * p.M * static_cast<double>(p.N) * p.K, stream);
    gemm_tcgen05_kernel<MODE><<<grid, THREADS, SMEM_BYTES, stream>>>(ta, tb, p);
  }
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

void gemm_debug_override(uint64_t desc_hi, uint32_t k_adv, uint32_t idesc) {
  g_dbg_desc_hi = desc_hi; g_dbg_k_adv = k_adv; g_dbg_idesc = idesc;
}

int gemm_bf16(const void* A, const void* W, const float* bias, int M, int N, int K, int mode, void* out, int ldo,
              const float* aux, int rows_per_group, int max_ctas, cudaStream_t stream) {
  VC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  VC_REQUIRE(K % BK == 0, "gemm: K=%d must be a multiple of %d (pad the operand)", K, BK);
  VC_REQUIRE(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
  VC_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "gemm: operands must be 16-byte aligned");
  VC_REQUIRE(get_encode() == 0, "gemm: cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  if (g_num_sms == 0) {
    int dev = 0;
    VC_CUDA_OK(cudaGetDevice(&dev));
    VC_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap ta, tb;
  if (int e = make_map(&ta, A, M, K, BM)) return e;
  if (int e = make_map(&tb, W, N, K, BN)) return e;
  GemmParams p{M, N, K, mode, bias, out, ldo, aux, rows_per_group,
               g_dbg_desc_hi ? g_dbg_desc_hi : umma_desc_sw128_hi(),
               g_dbg_k_adv ? g_dbg_k_adv : 2u,
               g_dbg_idesc ? g_dbg_idesc : umma_idesc_bf16(BM, BN)};
  const int total = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int grid = total < g_num_sms ? total : g_num_sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  switch (mode) {
    case VC_EPI_BIAS: return launch<VC_EPI_BIAS>(ta, tb, p, grid, stream);
    case VC_EPI_BIAS_GELU_ERF: return launch<VC_EPI_BIAS_GELU_ERF>(ta, tb, p, grid, stream);
    case VC_EPI_BIAS_GELU_TANH: return launch<VC_EPI_BIAS_GELU_TANH>(ta, tb, p, grid, stream);
    case VC_EPI_BIAS_RESID_F32: return launch<VC_EPI_BIAS_RESID_F32>(ta, tb, p, grid, stream);
    case VC_EPI_BIAS_F32: return launch<VC_EPI_BIAS_F32>(ta, tb, p, grid, stream);
    case VC_EPI_PATCH_EMBED:
      VC_REQUIRE(aux != nullptr && rows_per_group > 0, "gemm: patch-embed epilogue needs pos embedding and patches/frame");
      return launch<VC_EPI_PATCH_EMBED>(ta, tb, p, grid, stream);
    default: set_error("gemm: unknown epilogue mode %d", mode); return -1;
  }
}

}  // namespace vc
