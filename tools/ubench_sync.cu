// Micro-benchmark: cost of one dependent "phase" (read what other CTAs wrote, write, device-wide barrier) on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_sync tools/ubench_sync.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct P {
  unsigned int* bar;      // [0] counter, [32] flag (separate 128-byte lines)
  float* a; float* b;     // ping-pong data, n floats each
  int n, iters, mode, work, sleep_ns;
  const void* big;
};

__device__ __forceinline__ void bar_counter(unsigned int* bar, unsigned int& epoch, int sleep_ns) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const unsigned int target = epoch * gridDim.x;
    unsigned int v;
    for (;;) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (sleep_ns) __nanosleep(sleep_ns);
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

// last arriver publishes a flag on another line; pollers never touch the counter line
__device__ __forceinline__ void bar_flag(unsigned int* bar, unsigned int& epoch, int sleep_ns) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
    if (old == epoch * gridDim.x - 1) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 32), "r"(epoch) : "memory");
    } else {
      unsigned int v;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar + 32) : "memory");
        if (v >= epoch) break;
        if (sleep_ns) __nanosleep(sleep_ns);
      }
    }
  }
  __syncthreads();
}

// acquire load polling the counter directly (no separate fence)
__device__ __forceinline__ void bar_acq(unsigned int* bar, unsigned int& epoch, int sleep_ns) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const unsigned int target = epoch * gridDim.x;
    unsigned int v;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (sleep_ns) __nanosleep(sleep_ns);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(384) k(P p) {
  unsigned int epoch = 0;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  float* src = p.a; float* dst = p.b;
  for (int it = 0; it < p.iters; ++it) {
    if (p.work == 1) {
      // every thread reads a value another CTA wrote in the previous phase, writes one value
      for (int i = gt; i < p.n; i += nt) {
        const int j = (i + 4099) % p.n;
        dst[i] = __ldcg(src + j) + 1.f;
      }
    } else if (p.work == 2) {
      // two dependent L2 round trips (pointer chase flavour) before the store
      for (int i = gt; i < p.n; i += nt) {
        const int j = (i + 4099) % p.n;
        const float v = __ldcg(src + j);
        const int j2 = (j + 77 + (v > 1e30f ? 1 : 0)) % p.n;
        dst[i] = v + __ldcg(src + j2);
      }
    }
    if (p.work >= 3) {
      // stage 96 KB from L2: work 3 = every CTA reads the SAME region, 4 = a private region per CTA, 5 = 8 replicas shared by CTA groups
      const uint4* base = reinterpret_cast<const uint4*>(p.big) + (p.work == 3 ? 0 : p.work == 4 ? blockIdx.x * 6144 : (blockIdx.x & 7) * 6144);
      uint4 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = threadIdx.x + i * blockDim.x;
        if (c < 6144) v[i] = __ldcg(base + c);
      }
      unsigned int acc = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = threadIdx.x + i * blockDim.x;
        if (c < 6144) acc += v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
      }
      if (acc == 0x12345678u) dst[gt % p.n] = 1.f;
    }
    if (p.mode == 0) bar_counter(p.bar, epoch, p.sleep_ns);
    else if (p.mode == 1) bar_flag(p.bar, epoch, p.sleep_ns);
    else if (p.mode == 2) bar_acq(p.bar, epoch, p.sleep_ns);
    else __syncthreads();
    float* t = src; src = dst; dst = t;
  }
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned int* bar; float *a, *b;
  const int n = 1 << 16;
  CK(cudaMalloc(&bar, 1024)); CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
  CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 2000;
  printf("SMs %d; us per phase (iters %d)\n", sms, iters);
  const char* mn[] = {"red+poll(relaxed)+fence", "atom+flag", "red+poll(acquire)", "no grid barrier"};
  for (int per_sm = 1; per_sm <= 2; ++per_sm)
    for (int mode = 0; mode < 4; ++mode)
      for (int work = 0; work <= 2; ++work)
        for (int sl = 0; sl <= 0; sl += 20) {
          if (mode == 3 && sl) continue;
          P p{bar, a, b, n, iters, mode, work, sl, nullptr};
          void* args[] = {&p};
          float best = 1e9f;
          for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemset(bar, 0, 1024));
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel((void*)k, dim3(sms * per_sm), dim3(128), args, 0, 0));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
          }
          printf("ctas/sm %d  %-26s work %d sleep %2d ns : %6.2f us\n", per_sm, mn[mode], work, sl, best * 1e3f / iters);
        }
  void* big; CK(cudaMalloc(&big, 148 * 6144 * 16)); CK(cudaMemset(big, 1, 148 * 6144 * 16));
  for (int work = 3; work <= 5; ++work)
    for (int mode = 2; mode <= 3; ++mode) {
      P p{bar, a, b, n, iters, mode, work, 0, big};
      void* args[] = {&p};
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(bar, 0, 1024));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void*)k, dim3(sms), dim3(384), args, 0, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      printf("384 thr, stage 96 KB: work %d (%s) %-20s : %6.2f us\n", work, work == 3 ? "same region" : work == 4 ? "private region" : "8 replicas", mn[mode], best * 1e3f / iters);
    }
  return 0;
}
