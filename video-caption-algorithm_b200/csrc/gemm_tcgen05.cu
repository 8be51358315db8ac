// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epi(A[M,K] * W[N,K]^T)
//
//   * CTA pairs (cluster of 2, tcgen05 cta_group::2): one 256x256 output tile per pair, each CTA
//     stages its 128 rows of A and its 128 rows of W per 64-wide K block, so a pair moves 64 KB
//     per 8.4 MFLOP (128 flop/B) instead of the 85 flop/B of a single-CTA 128x256 tile — the
//     L2->SM fill rate, not the tensor pipe, was the limit of the single-CTA version;
//   * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a 4-deep mbarrier
//     ring; both CTAs' loads complete on the leader's "full" barrier, the MMA's tcgen05.commit
//     multicasts the "empty" arrive to both CTAs;
//   * one elected thread of the leader CTA issues tcgen05.mma (M=256, N=256, K=16); fp32
//     accumulators live in TMEM (128 lanes x 256 columns per CTA), double-buffered so the epilogue
//     of tile i overlaps the MMAs of tile i+1;
//   * 16 epilogue warps per CTA read TMEM with tcgen05.ld (one accumulator row per thread),
//     transpose 32x32 blocks through swizzled shared memory and then apply bias / GELU /
//     residual-add / patch-embed scatter with fully coalesced global accesses.
//
// This is the dense contraction of the ViT frame encoder (reference:
// src/models/video_encoder.py:288-326 -> torchvision/timm Linear + Conv2d patch
// embed) and of the GPT-2 prefill (transformers Conv1D).  nn.Linear weights are
// [out,in] = [N,K] K-major, which is exactly the B operand layout tcgen05 wants.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#ifndef VC_RESID_WARPS
#define VC_RESID_WARPS 16
#define VC_RESID_REGS 96
#endif
#ifndef VC_LNF_REGS
#define VC_LNF_REGS 96
#endif

namespace vc {

int make_tmap_bf16_kmajor(CUtensorMap* tm, const void* ptr, int rows, int K, int box_rows);

namespace {

constexpr int BM = 128;                // rows of A per CTA; a pair covers 256
constexpr int BK = 64, UK = 16;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
// Two tile widths (BN = columns per pair tile; each CTA stages BN/2 rows of W):
//   256: the encoder's shapes (thousands of tiles).  4 ring stages measure the same as 6 on every ViT shape (the ring only has to
//        cover the L2 latency).
//    64: few-row products (GPT-2 prefill / grouped decode steps of 128-1280 rows: ONE row tile): N/64 pairs instead of N/256
//        stream the weights, and a deeper ring (7 x 20 KB) because each pair walks its whole K with nothing else to overlap.
template <int BN>
struct Tile {
  static constexpr int STAGES = BN == 256 ? 4 : 7;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = BN == 256 ? 512 : 128;      // 2 accumulators x BN fp32 columns (power of two >= 32)
};
// Epilogue warps per CTA: a multiple of 4 (a warp reads the TMEM lane quarter warp % 4).  The bf16-output epilogues are
// instruction-bound (ncu: the GELU epilogue executes 3x the instructions of the bias one and holds the tensor pipe at 68 %
// active), so they get 16 warps = two 32-column chunks each.  The residual + statistics epilogue waits on HBM and keeps a block
// of residual values in flight per warp.  It ran with 12 warps x 128 registers (chunks 0-2, 3-5, 6-7: uneven) because
// 16 warps x 104 registers "does not launch" — the limit is per SM sub-partition: 18 warps put 5 on one scheduler, and
// 5 x 32 x 104 > 16,384 registers.  16 warps x 96 registers fits (one 8-byte spill) and measures 4 % faster over proj + fc2
// (13.6-13.8 vs 14.1-14.8 ms per pass, two builds alternated on one box): 64 KB of residual reads in flight instead of 48 KB,
// and two blocks for every warp.
__host__ __device__ constexpr int epi_warps(int mode) { return mode == VC_EPI_RESID_STATS ? VC_RESID_WARPS : 16; }
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;  // one 32x32 fp32 block per warp
__host__ __device__ constexpr int gemm_threads(int mode) { return (2 + epi_warps(mode)) * 32; }
// Split-K (KS CTA pairs of one cluster share a 64-column tile, each walks K/KS): a 4-stage ring is enough for K/KS, and every CTA
// owns 128/KS rows of its row half for the final epilogue: it receives the KS partial accumulators of those rows in RED_BYTES.
constexpr int KS_STAGES = 4;
constexpr int RED_PITCH = 64 + 4;                       // floats per row of a received partial (272 B: conflict-free float4 rows)
constexpr int RED_BYTES = 128 * RED_PITCH * 4;          // KS sources x 128/KS rows
__host__ __device__ constexpr int gemm_smem(int mode, int bn, int ks = 1) {
  return (bn == 256 ? Tile<256>::STAGES * Tile<256>::STAGE_BYTES : (ks > 1 ? KS_STAGES : Tile<64>::STAGES) * Tile<64>::STAGE_BYTES) +
         epi_warps(mode) * EPI_STAGE_BYTES + (ks > 1 ? RED_BYTES : 0) + 1024 /*align*/ + 256 /*barriers*/;
}

struct GemmParams {
  int M, N, K;
  int mode;
  const float* bias;      // [N] or null
  void* out;              // bf16 [M,ldo] or fp32 [rows,ldo]
  int ldo;                // leading dimension of out (elements)
  const float* aux;       // patch-embed: pos embedding [tokens, N]
  int rows_per_group;     // patch-embed: patches per frame (196); out row = g*(rpg+1)+1+r
  // LayerNorm-folded and residual/statistics epilogues (GemmExtra, vc_kernels.h)
  const float* cs;        // [N] column sums of the folded weights
  const float2* stats;    // [M] (mean, rstd) of the rows of A's source
  const void* w_ptr;      // W (for the L2 prefetch of this CTA's weight rows ahead of the dependency wait; 64-column tiles)
  int pre_flags;          // bit 0: weight tiles of the first ring stages before the dependency wait; bit 1: L2 prefetch of the rest
  float* xres;            // fp32 residual stream [M, N], read-modify-written (RESID_STATS)
  __nv_bfloat16* xb_out;  // bf16 [M, N]: bf16 copy of the new residual = A operand of the next folded product
  float2* pstats;         // [N/32][M] partial (sum, sum of squares) of the new residual rows, one slot per 32-column chunk
  float2* stats_out;      // RESID_STATS, few rows: the LAST CTA to finish turns the partials into (mean, rstd) per row here
  unsigned int* done;     // ... counter of finished CTAs (zero before the launch; the last CTA resets it)
  float eps;
};

// Predicated global accesses as single instructions: a C++ `if` around the store lets the compiler sink the whole
// epilogue math into a divergent region per row, which serialises the eight unrolled iterations.
__device__ __forceinline__ void st_pred_v2(void* ptr, uint32_t a, uint32_t b, bool ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.global.v2.b32 [%1], {%2, %3};\n\t}" ::"r"(static_cast<uint32_t>(ok)), "l"(ptr), "r"(a), "r"(b)
               : "memory");
}
__device__ __forceinline__ void st_pred_v4(void* ptr, float4 v, bool ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.global.v4.f32 [%1], {%2, %3, %4, %5};\n\t}" ::"r"(static_cast<uint32_t>(ok)), "l"(ptr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ld_pred_v4(const void* ptr, bool ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t@p ld.global.v4.f32 {%0, %1, %2, %3}, [%5];\n\t}"
               : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
               : "r"(static_cast<uint32_t>(ok)), "l"(ptr)
               : "memory");
  return v;
}

// One 32-row x 32-column accumulator block of one warp.  Phase 1: thread = row, raw fp32 accumulators into
// the warp's staging block (16-byte chunks XOR-swizzled by row: conflict-free both ways).  Phase 2: thread =
// (row group, 4-column group): 8 lanes cover one 128-byte row segment, so every global access is a full line.
template <int MODE>
__device__ __forceinline__ void epilogue_block(const GemmParams& p, uint8_t* stg, int lane, int row_base, int col0, const uint32_t (&acc)[32]) {
  const uint32_t sbase = smem_u32(stg);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t a = sbase + lane * 128 + ((j ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(acc[4 * j]), "r"(acc[4 * j + 1]), "r"(acc[4 * j + 2]), "r"(acc[4 * j + 3])
                 : "memory");
  }
  __syncwarp();
  const int cc = lane & 7, rsub = lane >> 3;
  const int col = col0 + cc * 4;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + rsub;
    float4 v;
    const uint32_t a = sbase + rr * 128 + ((cc ^ (rr & 7)) << 4);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    if (MODE == VC_EPI_BIAS_GELU_ERF) {
      v.x = gelu_erf_fast(v.x); v.y = gelu_erf_fast(v.y); v.z = gelu_erf_fast(v.z); v.w = gelu_erf_fast(v.w);
    } else if (MODE == VC_EPI_BIAS_GELU_TANH) {
      v.x = gelu_tanh_fast(v.x); v.y = gelu_tanh_fast(v.y); v.z = gelu_tanh_fast(v.z); v.w = gelu_tanh_fast(v.w);
    }
    const int row = row_base + rr;
    const bool ok = row < p.M;
    const int rowc = ok ? row : 0;
    if (MODE == VC_EPI_BIAS || MODE == VC_EPI_BIAS_GELU_ERF || MODE == VC_EPI_BIAS_GELU_TANH) {
      st_pred_v2(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(rowc) * p.ldo + col, pack_bf16(v.x, v.y), pack_bf16(v.z, v.w), ok);
    } else if (MODE == VC_EPI_BIAS_RESID_F32) {
      float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(rowc) * p.ldo + col;
      const float4 r = ld_pred_v4(o, ok);
      st_pred_v4(o, make_float4(r.x + v.x, r.y + v.y, r.z + v.z, r.w + v.w), ok);
    } else if (MODE == VC_EPI_BIAS_F32) {
      st_pred_v4(reinterpret_cast<float*>(p.out) + static_cast<size_t>(rowc) * p.ldo + col, v, ok);
    } else if (MODE == VC_EPI_PATCH_EMBED) {
      const int g = rowc / p.rows_per_group, r = rowc - g * p.rows_per_group;
      const size_t orow = static_cast<size_t>(g) * (p.rows_per_group + 1) + 1 + r;
      const float4 e = __ldg(reinterpret_cast<const float4*>(p.aux + static_cast<size_t>(1 + r) * p.N + col));
      st_pred_v4(reinterpret_cast<float*>(p.out) + orow * p.ldo + col, make_float4(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w), ok);
    }
  }
  __syncwarp();   // the next block reuses the staging buffer
}

// ---- LayerNorm folded into this product (A = bf16 of the un-normalised residual, W' = gamma (.) W):
//      out = act( rstd_row * (acc - mean_row * cs[col]) + bias'[col] )                       (VC_EPI_LNF_*)
// st[i] = (mean, rstd) of row row_base + 4 i + (lane >> 3): loaded once per tile by the caller (the rows of a warp do not
// change from one column chunk to the next).
template <int MODE>
__device__ __forceinline__ void epilogue_block_lnf(const GemmParams& p, uint8_t* stg, int lane, int row_base, int col0, const uint32_t (&acc)[32],
                                                   const float2 (&st)[8], const float4 b, const float4 c) {
  const uint32_t sbase = smem_u32(stg);
  const int cc = lane & 7, rsub = lane >> 3;
  const int col = col0 + cc * 4;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t a = sbase + lane * 128 + ((j ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(acc[4 * j]), "r"(acc[4 * j + 1]), "r"(acc[4 * j + 2]), "r"(acc[4 * j + 3])
                 : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + rsub;
    float4 v;
    const uint32_t a = sbase + rr * 128 + ((cc ^ (rr & 7)) << 4);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    const float mean = st[i].x, rstd = st[i].y;
    v.x = fmaf(rstd, v.x - mean * c.x, b.x); v.y = fmaf(rstd, v.y - mean * c.y, b.y);
    v.z = fmaf(rstd, v.z - mean * c.z, b.z); v.w = fmaf(rstd, v.w - mean * c.w, b.w);
    if (MODE == VC_EPI_LNF_GELU_ERF) {
      v.x = gelu_erf_fast(v.x); v.y = gelu_erf_fast(v.y); v.z = gelu_erf_fast(v.z); v.w = gelu_erf_fast(v.w);
    } else if (MODE == VC_EPI_LNF_GELU_TANH) {
      v.x = gelu_tanh_fast(v.x); v.y = gelu_tanh_fast(v.y); v.z = gelu_tanh_fast(v.z); v.w = gelu_tanh_fast(v.w);
    }
    const int row = row_base + rr;
    const bool ok = row < p.M;
    st_pred_v2(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(ok ? row : 0) * p.ldo + col, pack_bf16(v.x, v.y), pack_bf16(v.z, v.w), ok);
  }
  __syncwarp();   // the next block reuses the staging buffer
}

// ---- VC_EPI_RESID_STATS: the fp32 residual stream is updated HERE,  t = x + bf16(acc + bias);  x = t;  xb_out = bf16(t),
//      and the rows' (sum, sum of squares) leave as partials, so no stand-alone residual-add / LayerNorm pass over the stream
//      exists any more.  The Linear output is rounded to bf16 before the fp32 add — what autocast does.
// The residual values of a block (xv, 8 x 16 bytes per lane) are requested ONE BLOCK AHEAD by the caller — the first block of a
// tile while its MMAs are still running — so the epilogue does not pay a memory round trip per block (a read-modify-write
// issued inside the row loop made this product latency-bound in round 1: 250-420 TFLOP/s).
__device__ __forceinline__ void resid_prefetch(const GemmParams& p, int lane, int row_base, int col0, float4 (&xv)[8]) {
  const int col = col0 + (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row_base + i * 4 + rsub;
    const bool ok = row < p.M && col0 < p.N;
    xv[i] = ld_pred_v4(p.xres + static_cast<size_t>(ok ? row : 0) * p.N + (ok ? col : 0), ok);
  }
}
__device__ __forceinline__ void epilogue_park(uint8_t* stg, int lane, const uint32_t (&acc)[32]) {
  const uint32_t sbase = smem_u32(stg);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t a = sbase + lane * 128 + ((j ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(acc[4 * j]), "r"(acc[4 * j + 1]), "r"(acc[4 * j + 2]), "r"(acc[4 * j + 3])
                 : "memory");
  }
  __syncwarp();
}
__device__ __forceinline__ void epilogue_block_resid(const GemmParams& p, uint8_t* stg, int lane, int row_base, int col0, const float4 (&xv)[8],
                                                     const float4 b) {
  const uint32_t sbase = smem_u32(stg);
  const int cc = lane & 7, rsub = lane >> 3;
  const int col = col0 + cc * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + rsub;
    float4 v;
    const uint32_t a = sbase + rr * 128 + ((cc ^ (rr & 7)) << 4);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    const float2 e01 = unpack_bf16(pack_bf16(v.x + b.x, v.y + b.y)), e23 = unpack_bf16(pack_bf16(v.z + b.z, v.w + b.w));
    const float4 t = make_float4(xv[i].x + e01.x, xv[i].y + e01.y, xv[i].z + e23.x, xv[i].w + e23.y);
    const int row = row_base + rr;
    const bool ok = row < p.M;
    const size_t o = static_cast<size_t>(ok ? row : 0) * p.N + col;
    st_pred_v4(p.xres + o, t, ok);
    st_pred_v2(p.xb_out + o, pack_bf16(t.x, t.y), pack_bf16(t.z, t.w), ok);
    // this row's 32 columns: the 8 lanes that share the row are adjacent
    float s1 = (t.x + t.y) + (t.z + t.w), s2 = (t.x * t.x + t.y * t.y) + (t.z * t.z + t.w * t.w);
#pragma unroll
    for (int o2 = 1; o2 < 8; o2 <<= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o2);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o2);
    }
    // one slot per (32-column chunk, row): the grouping does not depend on the tile width or on which warp ran the chunk, so the
    // statistics — and with them every later bit — are the same whatever batch a row rides in
    if (cc == 0 && ok) p.pstats[static_cast<size_t>(col0 >> 5) * p.M + row] = make_float2(s1, s2);
  }
  __syncwarp();   // the next block reuses the staging buffer
}

// KS > 1 (RESID_STATS with 64-column tiles only): split-K inside a cluster of KS CTA pairs.  A few-row product with N = 768 has
// only 12 column tiles; 12 pairs each streaming all of A (256 rows x K) are bound by the L2 -> SM fill of those 24 SMs
// (15-21 us per product at 256 rows).  With KS pairs per tile every pair walks K/KS, leaves its 128 x 64 partial accumulator in
// TMEM, and the epilogue warps SEND each 32 x 32 block to the CTA that owns those rows (`st.shared::cluster`, one mbarrier arrive
// per lane); the owner adds the KS partials in a fixed order and runs the ordinary residual + statistics epilogue on its
// 128/KS rows.  One tile per cluster (grid = tiles x 2 KS CTAs).
template <int MODE, int BN, int KS = 1>
__global__ void __cluster_dims__(2 * KS, 1, 1) __maxnreg__(MODE == VC_EPI_RESID_STATS ? VC_RESID_REGS : (MODE >= VC_EPI_LNF_BIAS ? VC_LNF_REGS : 80))   // x 576 threads: 5 warps on one scheduler x 32 x 96 <= 16,384 registers
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmParams p) {
  static_assert(KS == 1 || (MODE == VC_EPI_RESID_STATS && BN == 64 && (KS == 2 || KS == 4)), "split-K: residual epilogue, 64-column tiles");
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must sit on 1024-byte boundaries (same offsets in both CTAs of the pair)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int STAGES = KS > 1 ? KS_STAGES : Tile<BN>::STAGES, B_BYTES = Tile<BN>::B_BYTES, STAGE_BYTES = Tile<BN>::STAGE_BYTES, TMEM_COLS = Tile<BN>::TMEM_COLS;
  (void)B_BYTES;
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;
  constexpr int EPI_WARPS = epi_warps(MODE);
  // Dependents may become resident early (their barrier init / TMEM allocation then overlaps this kernel's tail); this kernel itself
  // touches no global memory before pdl_wait() below.  No-ops when the kernel is not part of a programmatic-dependent-launch chain.
  pdl_launch_dependents();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + EPI_WARPS * EPI_STAGE_BYTES);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;    // per CTA: its smem slot is free (multicast commit)
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] per CTA: accumulator ready (multicast commit)
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] leader only: accumulator drained by both CTAs' epilogues
  uint64_t* red_bar = tempty_bar + 2;         // split-K: the partials of this CTA's rows have arrived (256 lane arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(red_bar + 1);
  float* red_buf = reinterpret_cast<float*>(epi_stage + EPI_WARPS * EPI_STAGE_BYTES + 256);   // split-K only (gemm_smem)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();   // KS == 1: 0 / 1
  const uint32_t rank = crank & 1;            // row half of the pair tile
  const uint32_t lead_rank = crank & ~1u;     // the pair's leader inside the cluster
  const int kq = static_cast<int>(crank >> 1);   // which K share (split-K), 0 otherwise
  const bool leader = rank == 0;
  const uint16_t pair_mask = static_cast<uint16_t>(0x3u << lead_rank);
  // work items: a pair tile per CTA pair, or (split-K) per cluster
  const int pair = blockIdx.x / (2 * KS), n_pairs = gridDim.x / (2 * KS);
  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int k_blocks = p.K / BK / KS;         // K blocks this pair walks ...
  const int kb_first = kq * k_blocks;         // ... starting here
  const int total = KS > 1 ? min(m_tiles * n_tiles, n_pairs) : m_tiles * n_tiles;   // split-K: one tile per cluster (host: grid = tiles)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * EPI_WARPS); }
    mbar_init(red_bar, 256);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b); }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, TMEM_COLS); tmem_relinquish_pair(); }
  tc_fence_before();
  cluster_sync_all();          // barriers initialised and TMEM allocated in both CTAs before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    // ------------------------------------------------ TMA producer (whole warp loops, one elected lane issues)
    int stage = 0; uint32_t phase = 0;
    const uint32_t full_leader0 = mapa_shared(smem_u32(&full_bar[0]), lead_rank);
    // W is a constant: with 64-column tiles (few-row chains, where a kernel is a chain of latencies) the weight halves of the
    // first tile's first ring stages are requested BEFORE the dependency wait, and the rest of this CTA's weight rows is pulled
    // into L2, so that after the wait only A has to arrive.
    int pre = 0;
    if (BN == 64 && pair < total && p.pre_flags != 0) {
      const int n0 = (pair % n_tiles) * BN + static_cast<int>(rank) * (BN / 2);
      pre = (p.pre_flags & 1) ? (k_blocks < STAGES ? k_blocks : STAGES) : 0;
      for (int kb = 0; kb < pre; ++kb) {
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&full_bar[kb], 2 * STAGE_BYTES);
          tma_load_2d_pair(&tm_b, full_leader0 + kb * 8, smem + kb * STAGE_BYTES + A_BYTES, (kb_first + kb) * BK, n0);
        }
        __syncwarp();
      }
      if ((p.pre_flags & 2) && k_blocks > pre && lane == 0) {
        // rows of this CTA's half tile, shared out over the K shares of the cluster (each row: K contiguous bf16)
        constexpr int ROWS = BN / 2 / KS;
        const int r0 = n0 + kq * ROWS;
        const int rows = r0 + ROWS <= p.N ? ROWS : (p.N > r0 ? p.N - r0 : 0);
        const uint8_t* src = static_cast<const uint8_t*>(p.w_ptr) + static_cast<size_t>(r0) * p.K * 2;
        for (long long off = 0, bytes = static_cast<long long>(rows) * p.K * 2; off < bytes; off += 32768)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + off), "r"(static_cast<uint32_t>(bytes - off < 32768 ? bytes - off : 32768))
                       : "memory");
      }
    }
    pdl_wait();                // the predecessor's outputs (A) are complete and visible from here on
    for (int t = pair; t < total; t += n_pairs) {
      const int m0 = (t / n_tiles) * (2 * BM) + static_cast<int>(rank) * BM;
      const int n0 = (t % n_tiles) * BN + static_cast<int>(rank) * (BN / 2);
      for (int kb = 0; kb < k_blocks; ++kb) {
        const bool w_there = t == pair && kb < pre;     // this stage's weights (and its expect) were issued ahead of the wait
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * STAGE_BYTES;
          if (leader && !w_there) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);   // both CTAs' bytes land on this barrier
          const uint32_t full_leader = full_leader0 + stage * 8;
          tma_load_2d_pair(&tm_a, full_leader, sa, (kb_first + kb) * BK, m0);
          if (!w_there) tma_load_2d_pair(&tm_b, full_leader, sa + A_BYTES, (kb_first + kb) * BK, n0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    pdl_wait();
    // ------------------------------------------------ MMA issuer (leader CTA; whole warp loops, one elected lane issues)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      constexpr uint64_t desc_hi = umma_desc_sw128_hi();
      const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = pair; t < total; t += n_pairs) {
        mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = smem_lo + stage * (STAGE_BYTES >> 4);
            const uint64_t da = desc_hi | a_lo;
            const uint64_t db = desc_hi | (a_lo + (A_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              // +32 B per K=16 slice inside the 128 B swizzle row (start address is in 16 B units)
              tc_mma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
            tc_commit_pair(&empty_bar[stage], pair_mask);   // both CTAs' slots reusable once these MMAs retire
            if (kb == k_blocks - 1) tc_commit_pair(&tfull_bar[acc], pair_mask);   // accumulator complete in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps (both CTAs)
    pdl_wait();                         // residual stream / statistics of the predecessor are complete and visible from here on
    const int ew = warp - 2;            // 0..7
    const int quarter = warp & 3;       // TMEM lane quarter this warp may read
    const int part = ew >> 2;           // which column chunks of the 256-wide tile: 0 -> 0..2, 1 -> 3..5, 2 -> 6..7
    // 32-column chunks of the tile dealt to the warp groups: 8 chunks / 4 groups = 2 each, / 3 groups = (0-2, 3-5, 6-7); a 64-wide
    // tile has 2 chunks, so some groups get none (they still take part in the accumulator hand-off)
    constexpr int CHUNKS = BN / 32, GROUPS = EPI_WARPS / 4;
    const int c_begin = (part * CHUNKS + GROUPS - 1) / GROUPS, c_end = ((part + 1) * CHUNKS + GROUPS - 1) / GROUPS;
    uint8_t* stg = epi_stage + ew * EPI_STAGE_BYTES;
    const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty_bar[0]), lead_rank);
    const uint32_t tempty_leader1 = mapa_shared(smem_u32(&tempty_bar[1]), lead_rank);
    int acc = 0; uint32_t acc_phase = 0;
    float4 xnext[8];
    if (MODE == VC_EPI_RESID_STATS && KS == 1 && pair < total)
      resid_prefetch(p, lane, (pair / n_tiles) * (2 * BM) + static_cast<int>(rank) * BM + quarter * 32, (pair % n_tiles) * BN + c_begin * 32, xnext);
    for (int t = pair; t < total; t += n_pairs) {
      const int m0 = (t / n_tiles) * (2 * BM) + static_cast<int>(rank) * BM;
      const int n0 = (t % n_tiles) * BN;
      const int row_base = m0 + quarter * 32;
      float2 st[8];
      if (MODE >= VC_EPI_LNF_BIAS && MODE <= VC_EPI_LNF_GELU_TANH) {
        // the rows' LayerNorm statistics, fetched while the MMAs of the tile run
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = row_base + i * 4 + (lane >> 3);
          st[i] = __ldcg(p.stats + (row < p.M ? row : 0));     // written by the kernel before this one: not the read-only path
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if constexpr (KS > 1) {
        constexpr int QPO = 4 / KS;                       // TMEM lane quarters (32 rows) per owning CTA
        const int owner = quarter / QPO;                  // K share index of the CTA that finishes these rows (same row half)
        const bool mine = owner == kq;
        float4 xcur[8];
        if (c_begin < c_end) {
          const int c = c_begin;                          // 64-column tile: one 32-column chunk per warp with a chunk
          if (mine) resid_prefetch(p, lane, row_base, n0 + c * 32, xcur);      // in flight while the partials travel
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c * 32, r);
          tmem_ld_wait();
          // this lane's row (quarter * 32 + lane of the row half), 32 columns -> slot kq of the owner's receive buffer
          const uint32_t dst = mapa_shared(smem_u32(red_buf + ((kq * QPO + (quarter - owner * QPO)) * 32 + lane) * RED_PITCH + c * 32),
                                           static_cast<uint32_t>(owner * 2) + rank);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + j * 16), "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]),
                         "r"(r[4 * j + 3])
                         : "memory");
          mbar_arrive_cluster(mapa_shared(smem_u32(red_bar), static_cast<uint32_t>(owner * 2) + rank));   // releases this lane's stores
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_signal(acc ? tempty_leader1 : tempty_leader0);
        if (c_begin < c_end && mine) {
          const int c = c_begin;
          mbar_wait_cluster(red_bar, 0);                  // all KS x QPO x 2 blocks of this CTA's rows are in
          const float* src = red_buf + ((quarter - owner * QPO) * 32 + lane) * RED_PITCH + c * 32;
          float4 a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = *reinterpret_cast<const float4*>(src + 4 * j);
#pragma unroll
          for (int s2 = 1; s2 < KS; ++s2) {                // fixed order: K shares 0, 1, ...
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(src + s2 * QPO * 32 * RED_PITCH + 4 * j);
              a[j].x += b.x; a[j].y += b.y; a[j].z += b.z; a[j].w += b.w;
            }
          }
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            r[4 * j] = __float_as_uint(a[j].x); r[4 * j + 1] = __float_as_uint(a[j].y);
            r[4 * j + 2] = __float_as_uint(a[j].z); r[4 * j + 3] = __float_as_uint(a[j].w);
          }
          const int col0 = n0 + c * 32;
          if (col0 < p.N && row_base < p.M) {
            epilogue_park(stg, lane, r);
            epilogue_block_resid(p, stg, lane, row_base, col0, xcur, __ldg(reinterpret_cast<const float4*>(p.bias + col0 + (lane & 7) * 4)));
          }
        }
      } else if (MODE == VC_EPI_RESID_STATS) {
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = n0 + c * 32;
          const bool live = col0 < p.N && row_base < p.M;
          float4 fb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live) fb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + (lane & 7) * 4));     // travels under the TMEM load
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c * 32, r);
          tmem_ld_wait();
          if (live) epilogue_park(stg, lane, r);
          float4 xcur[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) xcur[i] = xnext[i];
          // the block after this one: next chunk of this tile, or the first chunk of this pair's next tile
          if (c + 1 < c_end) {
            resid_prefetch(p, lane, row_base, n0 + (c + 1) * 32, xnext);
          } else if (t + n_pairs < total) {
            const int t2 = t + n_pairs;
            resid_prefetch(p, lane, (t2 / n_tiles) * (2 * BM) + static_cast<int>(rank) * BM + quarter * 32, (t2 % n_tiles) * BN + c_begin * 32, xnext);
          }
          if (live) epilogue_block_resid(p, stg, lane, row_base, col0, xcur, fb);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_signal(acc ? tempty_leader1 : tempty_leader0);
      } else {
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          const int col_in_tile = c * 32;
          const int col0 = n0 + col_in_tile;
          float4 fb = make_float4(0.f, 0.f, 0.f, 0.f), fc = fb;
          if (MODE >= VC_EPI_LNF_BIAS && col0 < p.N) {            // the chunk's folded bias and column sums travel under the TMEM load
            fb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + (lane & 7) * 4));
            fc = __ldg(reinterpret_cast<const float4*>(p.cs + col0 + (lane & 7) * 4));
          }
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col_in_tile, r);
          tmem_ld_wait();
          if (col0 < p.N && row_base < p.M) {
            if (MODE >= VC_EPI_LNF_BIAS) epilogue_block_lnf<MODE>(p, stg, lane, row_base, col0, r, st, fb, fc);
            else epilogue_block<MODE>(p, stg, lane, row_base, col0, r);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_signal(acc ? tempty_leader1 : tempty_leader0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (MODE == VC_EPI_RESID_STATS && p.stats_out != nullptr) {
      // Few rows (GPT-2 chains): no separate finalize kernel.  Every CTA counts itself done once its epilogue warps have stored
      // their partials; the last one combines the N/32 partials of every row (same arithmetic as ln_stats_finalize_kernel).
      __shared__ unsigned int s_last;
      __threadfence();                                                            // this thread's partials are out before the count
      asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");          // the epilogue warps of this CTA
      if (threadIdx.x == 64) {
        __threadfence();
        s_last = atomicAdd(p.done, 1u) == gridDim.x - 1 ? 1u : 0u;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
      if (s_last) {
        __threadfence();
        const int parts = p.N >> 5;
        for (int row = static_cast<int>(threadIdx.x) - 64; row < p.M; row += EPI_WARPS * 32) {
          double sm = 0.0, sq = 0.0;
          for (int c0 = 0; c0 < parts; c0 += 8) {           // eight loads in flight, summed in ascending chunk order
            float2 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = c0 + k < parts ? __ldcg(p.pstats + static_cast<size_t>(c0 + k) * p.M + row) : make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { sm += v[k].x; sq += v[k].y; }
          }
          const double mean = sm / p.N;
          double var = sq / p.N - mean * mean;
          var = var > 0.0 ? var : 0.0;
          p.stats_out[row] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps))));
        }
        if (threadIdx.x == 64) *p.done = 0u;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();          // the peer's smem / TMEM stay alive until every MMA and epilogue is done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
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

int make_map(CUtensorMap* tm, const void* ptr, int rows, int K, int box_rows) { return make_tmap_bf16_kmajor(tm, ptr, rows, K, box_rows); }

int g_num_sms = 0;
std::mutex g_cfg_mu;
PerDeviceOnce g_attr_set[16][4];

template <int MODE, int BN, int KS = 1>
int launch_bn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, cudaStream_t stream) {
  {
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    if (g_attr_set[MODE][BN == 256 ? 0 : (KS == 1 ? 1 : (KS == 2 ? 2 : 3))].first()) {
      VC_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<MODE, BN, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem(MODE, BN, KS)));
      VC_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<MODE, BN, KS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
  }
  static const char* const kNames[] = {"gemm_bias", "gemm_gelu_erf", "gemm_gelu_tanh", "gemm_resid", "gemm_f32", "gemm_patch",
                                       "gemm_lnf_bias", "gemm_lnf_gelu_erf", "gemm_lnf_gelu_tanh", "gemm_resid_stats"};
  {
    KernelScope ks(kNames[MODE], 2.0 * p.M * static_cast<double>(p.N) * p.K, stream);
    // programmatic dependent launch: inside the GPT-2 chains the next kernel's prologue overlaps this kernel's tail; after a
    // kernel that never triggers (the encoder's plain launches) it behaves like an ordinary launch
    VC_CUDA_OK(launch_pdl(gemm_tcgen05_kernel<MODE, BN, KS>, dim3(grid), dim3(gemm_threads(MODE)), static_cast<size_t>(gemm_smem(MODE, BN, KS)), stream, ta, tb,
                          p));
  }
  VC_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int MODE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, int bn, cudaStream_t stream) {
  return bn == 256 ? launch_bn<MODE, 256>(ta, tb, p, grid, stream) : launch_bn<MODE, 64>(ta, tb, p, grid, stream);
}

}  // namespace

// K-major bf16 [rows, K] matrix, box = [64, box_rows], 128-byte swizzle (shared with the decode kernel).
int make_tmap_bf16_kmajor(CUtensorMap* tm, const void* ptr, int rows, int K, int box_rows) {
  VC_REQUIRE(get_encode() == 0, "cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d box_rows=%d ptr=%p", static_cast<int>(r), rows, K, box_rows, ptr);
    return -3;
  }
  return 0;
}

// Same, rows `ld` elements apart (ld >= K): the last position of every sequence of [n_seq, L, K] activations (lm_head, decode_chain.cu).
int make_tmap_bf16_kmajor_ld(CUtensorMap* tm, const void* ptr, int rows, int K, long long ld, int box_rows) {
  VC_REQUIRE(get_encode() == 0, "cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%lld box_rows=%d ptr=%p", static_cast<int>(r), rows, K, ld, box_rows, ptr);
    return -3;
  }
  return 0;
}

static int current_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  static int sms_of[64] = {0};
  if (dev >= 0 && dev < 64 && sms_of[dev] != 0) return sms_of[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64) sms_of[dev] = n;
  return n;
}

// Tile width the GEMM uses for an [M, N] output: 256 columns when that already gives every CTA pair a tile, else 64
// (few-row products).  VC_GEMM_BN=64|256 forces one (A/B).
int gemm_tile_width(int M, int N) {
  static const int force_bn = getenv("VC_GEMM_BN") ? atoi(getenv("VC_GEMM_BN")) : 0;
  if (force_bn == 64 || force_bn == 256) return force_bn;
  const int m_tiles = (M + 2 * BM - 1) / (2 * BM);
  const int sms = current_sms();
  return m_tiles * ((N + 255) / 256) >= (sms > 0 ? sms : 148) / 2 ? 256 : 64;
}
// partial (sum, sum of squares) slots per row the RESID_STATS epilogue writes: one per 32-column chunk (ln_stats_finalize's `parts`)
int gemm_resid_parts(int N) { return N / 32; }

int gemm_bf16(const void* A, const void* W, const float* bias, int M, int N, int K, int mode, void* out, int ldo,
              const float* aux, int rows_per_group, int max_ctas, cudaStream_t stream) {
  return gemm_bf16_ex(A, W, bias, M, N, K, mode, out, ldo, aux, rows_per_group, max_ctas, nullptr, stream);
}

int gemm_bf16_ex(const void* A, const void* W, const float* bias, int M, int N, int K, int mode, void* out, int ldo,
                 const float* aux, int rows_per_group, int max_ctas, const GemmExtra* ex, cudaStream_t stream) {
  VC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  VC_REQUIRE(mode < VC_EPI_LNF_BIAS || (ex != nullptr && bias != nullptr), "gemm: epilogue %d needs the extra operands", mode);
  VC_REQUIRE(K % BK == 0, "gemm: K=%d must be a multiple of %d (pad the operand)", K, BK);
  VC_REQUIRE(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
  VC_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "gemm: operands must be 16-byte aligned");
  VC_REQUIRE(get_encode() == 0, "gemm: cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  g_num_sms = current_sms();
  VC_REQUIRE(g_num_sms > 0, "gemm: cannot query the device");
  // tile width: 256 columns when that already gives every CTA pair a tile, else 64 (few-row products: one row tile)
  const int m_tiles = (M + 2 * BM - 1) / (2 * BM);
  const int BN = gemm_tile_width(M, N);
  CUtensorMap ta, tb;
  if (int e = make_map(&ta, A, M, K, BM)) return e;
  if (int e = make_map(&tb, W, N, K, BN / 2)) return e;
  GemmParams p{M, N, K, mode, bias, out, ldo, aux, rows_per_group, nullptr, nullptr, W, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f};
  if (ex != nullptr) {
    p.cs = ex->cs; p.stats = reinterpret_cast<const float2*>(ex->stats); p.xres = ex->xres;
    p.xb_out = static_cast<__nv_bfloat16*>(ex->xb_out);
    p.pstats = reinterpret_cast<float2*>(ex->pstats);
    p.stats_out = reinterpret_cast<float2*>(ex->stats_out); p.done = ex->done; p.eps = ex->eps;
  }
  static const int pre_flags = getenv("VC_GEMM_PRE") ? atoi(getenv("VC_GEMM_PRE")) : 3;
  p.pre_flags = pre_flags;
  const int total = m_tiles * ((N + BN - 1) / BN);   // pair tiles
  // VC_ENCODER_SMS=n leaves SMs free for kernels of another stream (the decode chain of the previous batch)
  static const int sm_cap = getenv("VC_ENCODER_SMS") ? atoi(getenv("VC_ENCODER_SMS")) : 0;
  int pairs = (sm_cap > 0 && sm_cap < g_num_sms ? sm_cap : g_num_sms) / 2;
  if (total < pairs) pairs = total;
  if (max_ctas > 0 && 2 * pairs > max_ctas) pairs = max_ctas / 2 > 0 ? max_ctas / 2 : 1;
  const int grid = 2 * pairs;
  if (mode == VC_EPI_RESID_STATS && BN == 64 && max_ctas <= 0 && sm_cap <= 0) {
    // few column tiles (N = 768: 12 per row tile): split K over the CTA pairs of a cluster, one tile per cluster
    static const int env_ks = getenv("VC_GEMM_KS") ? atoi(getenv("VC_GEMM_KS")) : 0;        // 1 = never split (A/B)
    const int force_ks = ex != nullptr && ex->split_k > 0 ? ex->split_k : env_ks;
    const int kb = K / BK;
    int ks = total * 8 <= g_num_sms && kb % 4 == 0 ? 4 : (total * 4 <= g_num_sms && kb % 2 == 0 ? 2 : 1);
    if (force_ks == 1 || ((force_ks == 2 || force_ks == 4) && kb % force_ks == 0)) ks = force_ks;
    if (ks == 4) return launch_bn<VC_EPI_RESID_STATS, 64, 4>(ta, tb, p, total * 8, stream);
    if (ks == 2) return launch_bn<VC_EPI_RESID_STATS, 64, 2>(ta, tb, p, total * 4, stream);
  }
  switch (mode) {
    case VC_EPI_BIAS: return launch<VC_EPI_BIAS>(ta, tb, p, grid, BN, stream);
    case VC_EPI_BIAS_GELU_ERF: return launch<VC_EPI_BIAS_GELU_ERF>(ta, tb, p, grid, BN, stream);
    case VC_EPI_BIAS_GELU_TANH: return launch<VC_EPI_BIAS_GELU_TANH>(ta, tb, p, grid, BN, stream);
    case VC_EPI_BIAS_RESID_F32: return launch<VC_EPI_BIAS_RESID_F32>(ta, tb, p, grid, BN, stream);
    case VC_EPI_BIAS_F32: return launch<VC_EPI_BIAS_F32>(ta, tb, p, grid, BN, stream);
    case VC_EPI_PATCH_EMBED:
      VC_REQUIRE(aux != nullptr && rows_per_group > 0, "gemm: patch-embed epilogue needs pos embedding and patches/frame");
      return launch<VC_EPI_PATCH_EMBED>(ta, tb, p, grid, BN, stream);
    case VC_EPI_LNF_BIAS: return launch<VC_EPI_LNF_BIAS>(ta, tb, p, grid, BN, stream);
    case VC_EPI_LNF_GELU_ERF: return launch<VC_EPI_LNF_GELU_ERF>(ta, tb, p, grid, BN, stream);
    case VC_EPI_LNF_GELU_TANH: return launch<VC_EPI_LNF_GELU_TANH>(ta, tb, p, grid, BN, stream);
    case VC_EPI_RESID_STATS: return launch<VC_EPI_RESID_STATS>(ta, tb, p, grid, BN, stream);
    default: set_error("gemm: unknown epilogue mode %d", mode); return -1;
  }
}

}  // namespace vc
