// Micro-benchmark: how long does it take EVERY CTA of a grid to pull the same [64 x 768] bf16 activation block (96 KB, freshly
// written by a previous kernel, so L2-resident) into its shared memory?  This is the dependent fetch of the decode chain's
// full-K products (csrc/decode_chain.cu).  Variants: 16-byte cp.async, LDG.128 + STS, one bulk copy per row, bulk copies
// multicast across a cluster of 2 / 4 / 8 CTAs (each CTA fetches 1/C of the rows for everybody).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fill.bin tools/ubench_fill.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ROWS = 64, KB = 1536, PITCH = KB + 64, THREADS = 256;

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void producer(uint4* x, int n16, unsigned v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) x[i] = make_uint4(v + i, v, i, 1);
}

// mode 0: cp.async 16B; 1: LDG.128 + STS; 2: bulk copy per row (own rows = all); 3: bulk multicast, cluster size = csize
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) consumer(const uint8_t* __restrict__ x, int rows, unsigned long long* times, unsigned* sink, int csize) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + ROWS * PITCH);
  const int tid = threadIdx.x;
  unsigned rank = 0;
  if (MODE >= 2) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (MODE == 3) {
      asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
  }
  const unsigned long long t0 = gtime();
  if (MODE == 0) {
    constexpr int CPR = KB / 16;
    for (int i = tid; i < rows * CPR; i += THREADS) {
      const int r = i / CPR, c = i - r * CPR;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(sm + r * PITCH + c * 16)), "l"(x + (size_t)r * KB + c * 16) : "memory");
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  } else if (MODE == 1) {
    constexpr int CPR = KB / 16;
    uint4 v[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      const int i = tid + j * THREADS;
      if (i < rows * CPR) v[j] = *reinterpret_cast<const uint4*>(x + (size_t)(i / CPR) * KB + (i % CPR) * 16);
    }
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      const int i = tid + j * THREADS;
      if (i < rows * CPR) *reinterpret_cast<uint4*>(sm + (i / CPR) * PITCH + (i % CPR) * 16) = v[j];
    }
    __syncthreads();
  } else if (MODE == 2) {
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(rows * KB) : "memory");
    if (tid < rows)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + tid * PITCH)),
                   "l"(x + (size_t)tid * KB), "r"(KB), "r"(s32(bar))
                   : "memory");
  } else {
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(rows * KB) : "memory");
    const int per = rows / csize;
    if (tid < per) {
      const int r = rank * per + tid;
      const unsigned short mask = static_cast<unsigned short>((1u << csize) - 1);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                       s32(sm + r * PITCH)),
                   "l"(x + (size_t)r * KB), "r"(KB), "r"(s32(bar)), "h"(mask)
                   : "memory");
    }
  }
  if (MODE >= 2) {
    unsigned ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)) : "memory");
  }
  const unsigned long long t1 = gtime();
  unsigned acc = 0;
  for (int i = tid; i < rows * KB / 4; i += THREADS) acc += reinterpret_cast<const unsigned*>(sm)[(i / (KB / 4)) * (PITCH / 4) + i % (KB / 4)];
  if (acc == 0x12345u) sink[0] = acc;
  if (tid == 0) { times[2 * blockIdx.x] = t0; times[2 * blockIdx.x + 1] = t1; }
  if (MODE == 3) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int MODE>
void run(const char* name, int grid, int csize, uint8_t* x, unsigned long long* d_times, unsigned* sink) {
  const size_t smem = ROWS * PITCH + 64;
  CK(cudaFuncSetAttribute(consumer<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (MODE == 3) CK(cudaFuncSetAttribute(consumer<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  double best_max = 1e9, best_med = 1e9, wall = 1e9;
  for (int it = 0; it < 6; ++it) {
    producer<<<64, 256>>>(reinterpret_cast<uint4*>(x), ROWS * KB / 16, it);   // freshly written by other SMs -> L2
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = MODE == 3 ? csize : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, consumer<MODE>, (const uint8_t*)x, ROWS, d_times, sink, csize));
    CK(cudaDeviceSynchronize());
    static unsigned long long h[2 * 512];
    CK(cudaMemcpy(h, d_times, sizeof(unsigned long long) * 2 * grid, cudaMemcpyDeviceToHost));
    unsigned long long first = ~0ULL, last = 0;
    double mx = 0, sum = 0;
    for (int b = 0; b < grid; ++b) {
      const double d = (h[2 * b + 1] - h[2 * b]) / 1e3;
      mx = d > mx ? d : mx; sum += d;
      first = h[2 * b] < first ? h[2 * b] : first; last = h[2 * b + 1] > last ? h[2 * b + 1] : last;
    }
    if (it >= 1) {
      best_max = mx < best_max ? mx : best_max;
      best_med = sum / grid < best_med ? sum / grid : best_med;
      wall = (last - first) / 1e3 < wall ? (last - first) / 1e3 : wall;
    }
  }
  printf("%-28s grid %3d cluster %d : per-CTA fetch mean %.2f us, slowest CTA %.2f us, first start -> last arrival %.2f us\n", name, grid, csize, best_med,
         best_max, wall);
}

int main() {
  uint8_t* x; unsigned long long* t; unsigned* sink;
  CK(cudaMalloc(&x, ROWS * KB)); CK(cudaMalloc(&t, sizeof(unsigned long long) * 2 * 512)); CK(cudaMalloc(&sink, 4));
  for (int grid : {1, 16, 144}) {
    run<0>("cp.async 16B", grid, 1, x, t, sink);
    run<1>("LDG.128 + STS", grid, 1, x, t, sink);
    run<2>("bulk copy per row", grid, 1, x, t, sink);
  }
  run<3>("bulk multicast", 144, 2, x, t, sink);
  run<3>("bulk multicast", 144, 4, x, t, sink);
  run<3>("bulk multicast", 128, 8, x, t, sink);
  run<3>("bulk multicast", 96, 8, x, t, sink);
  run<3>("bulk multicast", 128, 16, x, t, sink);
  return 0;
}
