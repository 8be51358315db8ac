// ViT self-attention on tcgen05: one (frame, head) per work item, S = Q K^T and O = P V on the 5th-gen tensor
// cores with both accumulators in TMEM, P handed back to the tensor core THROUGH TMEM (A-from-TMEM MMA), nothing
// but Q/K/V tiles in shared memory.
//
//   warp 0      TMA producer: Q (2 x [128 x 64]), K and V ([NK x 64]) of the next item into a 2-deep ring
//   warp 1      MMA issuer:   S_t = Q_t K^T (M=128, N=NK, K=64), then O_t = P_t V (A = P from TMEM, B = V MN-major)
//   warps 2-5   softmax + output of query tile 0 (rows 0..127 of the frame)
//   warps 6-9   softmax + output of query tile 1 (rows 128..255)
//
// TMEM, per query tile (256 columns): S fp32 in [0, NK); after a thread has read its row it overwrites the first
// NK/2 columns with P as packed bf16 pairs; O fp32 accumulates in [128, 192) (S is dead by then).  NK = tokens
// rounded up to 16 (197 -> 208); keys >= tokens are masked to -inf, so their P is exactly 0.
//
// tokens in (256, 264] (ViT-L/14 has 257): the MMA covers keys 0..255 (N = 256 is the instruction's limit and the TMEM
// budget of a tile); the few keys beyond are folded in on the CUDA cores of the softmax warps — score = q_row . k from the
// Q tile in shared memory, and p * v added to the O row in the output pass — and the query rows beyond 255 are computed by
// vit_row_attention (one warp per row).
//
// Softmax is fp32, exp2 with the 1/sqrt(64) * log2(e) scale folded in, row sum accumulated in fp32 before the bf16
// rounding of P, output normalised at the end — the same arithmetic as the mma.sync kernel this replaces for
// tokens <= 256 (reference: nn.MultiheadAttention / timm Attention -> SDPA, src/models/video_encoder.py:112-121).
// The exponentials bound the kernel: 128 x NK ex2 per tile at 16 MUFU/clk/SM ~ 1.7k cycles vs ~0.8k of MMA.
#include "vc_common.cuh"
#include "vc_kernels.h"

#include <cudaTypedefs.h>
#include <cstdlib>

namespace vc {

int make_tmap_bf16_kmajor(CUtensorMap* tm, const void* ptr, int rows, int K, int box_rows);

namespace {

constexpr int AT_THREADS = 352;            // warp 0 TMA, warps 1 and 10 MMA issuers (one per query tile), warps 2-9 softmax
constexpr int AT_MAX_EXTRA = 8;              // keys / query rows beyond 256 handled outside the MMA
constexpr int AT_Q_TILE = 128 * 128;            // bytes: 128 rows x 64 dims bf16
constexpr int AT_KV_MAX = 256 * 128;            // bytes reserved for K (and V): up to 256 keys
constexpr int AT_BUF = 2 * AT_Q_TILE + 2 * AT_KV_MAX;   // one item: Q0 Q1 K V = 96 KB
constexpr int AT_SMEM = 2 * AT_BUF + 1024 + 256;

struct AttParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* qkv;   // for the keys beyond the 256 the MMA covers (ViT-L/14: key 256)
  int n_frames, tokens, heads, D, NK, m_tiles, extra;
};

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x16b(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// D[tmem] (+)= A[tmem, bf16 pairs] * B[smem desc]
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(AT_THREADS, 1)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const AttParams p) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* qkv_full = reinterpret_cast<uint64_t*>(smem + 2 * AT_BUF);   // [2]
  uint64_t* qkv_empty = qkv_full + 2;                                    // [2]
  uint64_t* s_full = qkv_empty + 2;                                      // [2] per query tile
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_frames * p.heads;
  const int NK = p.NK, m_tiles = p.m_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qkv_full[i], 1); mbar_init(&qkv_empty[i], m_tiles);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    const uint32_t bytes = m_tiles * AT_Q_TILE + 2 * NK * 128;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t par = (it >> 1) & 1;
      const int frame = item / p.heads, head = item - frame * p.heads;
      const int row0 = frame * p.tokens;
      mbar_wait(&qkv_empty[buf], par ^ 1);
      if (elect_one()) {
        uint8_t* b = smem + buf * AT_BUF;
        mbar_arrive_expect_tx(&qkv_full[buf], bytes);
        tma_load_2d(&tm_q, &qkv_full[buf], b, head * 64, row0);
        if (m_tiles == 2) tma_load_2d(&tm_q, &qkv_full[buf], b + AT_Q_TILE, head * 64, row0 + 128);
        tma_load_2d(&tm_kv, &qkv_full[buf], b + 2 * AT_Q_TILE, p.D + head * 64, row0);
        tma_load_2d(&tm_kv, &qkv_full[buf], b + 2 * AT_Q_TILE + AT_KV_MAX, 2 * p.D + head * 64, row0);
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 10) {
    // ------------------------------------------------ MMA issuers: one warp per query tile, so the two tiles of a (frame,
    // head) are NOT in lockstep.  Tile 1 starts half a period late (after tile 0's first softmax) and stays there: while
    // one tile's softmax warps own the MUFU pipes, the other tile is in its MMA / output / wait phases.
    const int t = warp == 1 ? 0 : 1;
    if (t < m_tiles && static_cast<int>(blockIdx.x) < n_items) {
      constexpr uint64_t desc_hi = umma_desc_sw128_hi();
      const uint32_t idesc_s = umma_idesc_bf16(128, NK);
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major: [keys][64 dims]
      const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
      if (t == 1) mbar_wait(&p_full[0], 0);
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t par_buf = (it >> 1) & 1, par = it & 1;
        const uint32_t b_lo = smem_lo + buf * (AT_BUF >> 4);
        mbar_wait(&qkv_full[buf], par_buf);
        mbar_wait(&o_empty[t], par ^ 1);           // previous item's output of this tile has left TMEM
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dq = desc_hi | (b_lo + t * (AT_Q_TILE >> 4));
          const uint64_t dk = desc_hi | (b_lo + (2 * AT_Q_TILE >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base + t * 256, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          tc_commit(&s_full[t]);
        }
        __syncwarp();
        mbar_wait(&p_full[t], par);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dv = desc_hi | (b_lo + ((2 * AT_Q_TILE + AT_KV_MAX) >> 4));
          const int ksteps = NK >> 4;
          for (int k = 0; k < ksteps; ++k)
            tc_mma_bf16_ts(tmem_base + t * 256 + 128, tmem_base + t * 256 + 8 * k, dv + 128 * k, idesc_pv, k != 0);
          tc_commit(&o_full[t]);
          tc_commit(&qkv_empty[buf]);              // every MMA of this tile that reads the buffer has retired (count = tiles)
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ softmax + output, one query tile per warp group
    const int t = (warp - 2) >> 2;                  // query tile of this warp
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    if (t < m_tiles) {
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
      const int row_in_frame = t * 128 + quarter * 32 + lane;
      const float scale = 0.125f * 1.4426950408889634f;   // head_dim^-0.5 * log2(e)
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        const int frame = item / p.heads, head = item - frame * p.heads;
        // keys beyond the MMA's 256: raw scores q_row . k_j on the CUDA cores, before waiting for the tensor core
        float sx[AT_MAX_EXTRA];
        uint4 vx0[8];
        if (p.extra > 0) {
          const int buf = it & 1;
          mbar_wait(&qkv_full[buf], (it >> 1) & 1);            // the Q tile of this item has landed
          const uint8_t* qrow = smem + buf * AT_BUF + t * AT_Q_TILE + (quarter * 32 + lane) * 128;
          const int rsw = (quarter * 32 + lane) & 7;
          float q[64];
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint4 u = *reinterpret_cast<const uint4*>(qrow + ((c8 ^ rsw) << 4));
            const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
            q[c8 * 8 + 0] = a.x; q[c8 * 8 + 1] = a.y; q[c8 * 8 + 2] = b.x; q[c8 * 8 + 3] = b.y;
            q[c8 * 8 + 4] = cc.x; q[c8 * 8 + 5] = cc.y; q[c8 * 8 + 6] = d.x; q[c8 * 8 + 7] = d.y;
          }
          for (int e = 0; e < p.extra; ++e) {
            const uint4* kp = reinterpret_cast<const uint4*>(p.qkv + (static_cast<size_t>(frame) * p.tokens + NK + e) * 3 * p.D + p.D + head * 64);
            float acc = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              const uint4 u = __ldg(kp + c8);
              const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
              acc = fmaf(q[c8 * 8 + 0], a.x, acc); acc = fmaf(q[c8 * 8 + 1], a.y, acc); acc = fmaf(q[c8 * 8 + 2], b.x, acc);
              acc = fmaf(q[c8 * 8 + 3], b.y, acc); acc = fmaf(q[c8 * 8 + 4], cc.x, acc); acc = fmaf(q[c8 * 8 + 5], cc.y, acc);
              acc = fmaf(q[c8 * 8 + 6], d.x, acc); acc = fmaf(q[c8 * 8 + 7], d.y, acc);
            }
            sx[e] = acc;
          }
          // V row of the first extra key: requested now (q is dead), used in the output pass — its L2 latency hides behind
          // the S MMA and the softmax instead of sitting between the PV MMA and the store
          const uint4* vp0 = reinterpret_cast<const uint4*>(p.qkv + (static_cast<size_t>(frame) * p.tokens + NK) * 3 * p.D + 2 * p.D + head * 64);
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) vx0[c8] = __ldg(vp0 + c8);
        }
        mbar_wait(&s_full[t], par);
        tc_fence_after();
        // A warp whose 32 query rows all lie beyond the frame's tokens (rows 224-255 at 197 tokens) skips both softmax passes:
        // its TMEM lanes keep the raw scores as "P", the PV MMA turns them into output rows nobody stores.
        const bool dead = t * 128 + quarter * 32 >= p.tokens;
        // pass 1: row max over the valid keys (64 columns per TMEM load: the loop is latency-, not bandwidth-bound)
        float mx = -INFINITY;
        for (int e = 0; e < p.extra; ++e) mx = fmaxf(mx, sx[e]);
        int c = dead ? NK : 0;
        for (; c + 64 <= NK; c += 64) {
          uint32_t r[64];
          tmem_ld_32x64(trow + c, r);
          tmem_ld_wait();
          if (c + 64 <= p.tokens) {
#pragma unroll
            for (int j = 0; j < 64; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (c + j < p.tokens) mx = fmaxf(mx, __uint_as_float(r[j]));
          }
        }
        for (; c < NK; c += 16) {
          uint32_t r[16];
          tmem_ld_32x16b(trow + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c + j < p.tokens) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        const float m2 = mx * scale;
        // pass 2: p = 2^(s*scale - m2); row sum in fp32; P as bf16 pairs over S columns this thread has already consumed
        // (keys [c, c+64) -> packed columns [c/2, c/2+32), always at or below the columns being read)
        float sum = 0.f;
        c = dead ? NK : 0;
        for (; c + 64 <= NK; c += 64) {
          uint32_t r[64], pk[32];
          tmem_ld_32x64(trow + c, r);
          tmem_ld_wait();
          const bool full = c + 64 <= p.tokens;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e0 = fast_ex2(fmaf(__uint_as_float(r[2 * j]), scale, -m2));
            float e1 = fast_ex2(fmaf(__uint_as_float(r[2 * j + 1]), scale, -m2));
            if (!full) {
              if (c + 2 * j >= p.tokens) e0 = 0.f;
              if (c + 2 * j + 1 >= p.tokens) e1 = 0.f;
            }
            sum += e0 + e1;
            pk[j] = pack_bf16(e0, e1);
          }
          tmem_st_32x32(trow + (c >> 1), pk);
        }
        for (; c < NK; c += 16) {
          uint32_t r[16], pk[8];
          tmem_ld_32x16b(trow + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float e0 = c + 2 * j < p.tokens ? fast_ex2(fmaf(__uint_as_float(r[2 * j]), scale, -m2)) : 0.f;
            const float e1 = c + 2 * j + 1 < p.tokens ? fast_ex2(fmaf(__uint_as_float(r[2 * j + 1]), scale, -m2)) : 0.f;
            sum += e0 + e1;
            pk[j] = pack_bf16(e0, e1);
          }
          tmem_st_32x8(trow + (c >> 1), pk);
        }
        for (int e = 0; e < p.extra; ++e) {
          sx[e] = fast_ex2(fmaf(sx[e], scale, -m2));        // now the (unrounded) probability of the extra key
          sum += sx[e];
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        // output: O / sum
        mbar_wait(&o_full[t], par);
        tc_fence_after();
        const float inv = 1.f / sum;
        const bool row_ok = row_in_frame < p.tokens;     // (tiles cover rows 0..255; rows beyond: vit_row_attention)
        __nv_bfloat16* orow = p.out + (static_cast<size_t>(frame) * p.tokens + row_in_frame) * p.D + head * 64;
        {
          uint32_t r[64];
          tmem_ld_32x64(trow + 128, r);
          tmem_ld_wait();
          for (int e = 0; e < p.extra; ++e) {               // + p_j * v_j of the keys the MMA did not cover
            const uint4* vp = reinterpret_cast<const uint4*>(p.qkv + (static_cast<size_t>(frame) * p.tokens + NK + e) * 3 * p.D + 2 * p.D + head * 64);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              const uint4 u = e == 0 ? vx0[c8] : __ldg(vp + c8);
              const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
              r[c8 * 8 + 0] = __float_as_uint(fmaf(sx[e], a.x, __uint_as_float(r[c8 * 8 + 0])));
              r[c8 * 8 + 1] = __float_as_uint(fmaf(sx[e], a.y, __uint_as_float(r[c8 * 8 + 1])));
              r[c8 * 8 + 2] = __float_as_uint(fmaf(sx[e], b.x, __uint_as_float(r[c8 * 8 + 2])));
              r[c8 * 8 + 3] = __float_as_uint(fmaf(sx[e], b.y, __uint_as_float(r[c8 * 8 + 3])));
              r[c8 * 8 + 4] = __float_as_uint(fmaf(sx[e], cc.x, __uint_as_float(r[c8 * 8 + 4])));
              r[c8 * 8 + 5] = __float_as_uint(fmaf(sx[e], cc.y, __uint_as_float(r[c8 * 8 + 5])));
              r[c8 * 8 + 6] = __float_as_uint(fmaf(sx[e], d.x, __uint_as_float(r[c8 * 8 + 6])));
              r[c8 * 8 + 7] = __float_as_uint(fmaf(sx[e], d.y, __uint_as_float(r[c8 * 8 + 7])));
            }
          }
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint4 w;
              w.x = pack_bf16(__uint_as_float(r[8 * q + 0]) * inv, __uint_as_float(r[8 * q + 1]) * inv);
              w.y = pack_bf16(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv);
              w.z = pack_bf16(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv);
              w.w = pack_bf16(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv);
              reinterpret_cast<uint4*>(orow)[q] = w;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[t]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_at_sms = 0;
PerDeviceOnce g_at_attr;

}  // namespace

bool vit_attention_tc_supported(int tokens, int heads, int head_dim) {
  return head_dim == 64 && tokens >= 1 && tokens <= 256 + AT_MAX_EXTRA && heads >= 1;
}

int vit_attention_tc(const void* qkv, void* out, int n_frames, int tokens, int heads, cudaStream_t s) {
  const int D = heads * 64;
  const int NK = tokens > 256 ? 256 : (tokens + 15) / 16 * 16;     // keys covered by the MMA
  const int extra = tokens > 256 ? tokens - 256 : 0;               // keys and query rows handled outside it
  const int m_tiles = tokens > 128 ? 2 : 1;
  if (g_at_sms == 0) {
    int dev = 0;
    VC_CUDA_OK(cudaGetDevice(&dev));
    VC_CUDA_OK(cudaDeviceGetAttribute(&g_at_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (g_at_attr.first()) VC_CUDA_OK(cudaFuncSetAttribute(vit_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
  CUtensorMap tq, tkv;
  const int rows = n_frames * tokens;
  int e;
  if ((e = make_tmap_bf16_kmajor(&tq, qkv, rows, 3 * D, 128))) return e;
  if ((e = make_tmap_bf16_kmajor(&tkv, qkv, rows, 3 * D, NK))) return e;
  AttParams p{static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(qkv), n_frames, tokens, heads, D, NK, m_tiles, extra};
  const int items = n_frames * heads;
  static const int sm_cap = getenv("VC_ENCODER_SMS") ? atoi(getenv("VC_ENCODER_SMS")) : 0;
  const int sms = sm_cap > 0 && sm_cap < g_at_sms ? sm_cap : g_at_sms;
  const int grid = items < sms ? items : sms;
  {
    KernelScope ks("vit_attention", 4.0 * n_frames * heads * static_cast<double>(tokens) * tokens * 64, s);
    vit_attention_tc_kernel<<<grid, AT_THREADS, AT_SMEM, s>>>(tq, tkv, p);
  }
  VC_CUDA_OK(cudaGetLastError());
  if (extra > 0) return vit_row_attention(qkv, out, n_frames, tokens, heads, 256, extra, s);   // query rows 256..tokens-1
  return 0;
}

}  // namespace vc
