// Micro-benchmark: ex2.approx.ftz.f32 (MUFU.EX2) and FFMA throughput per SM on sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>
__global__ void k(float* out, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      else if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(0.999f), "f"(-0.001f));
      else { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(0.999f), "f"(-0.001f));
             asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(0.999f), "f"(-0.001f)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(0.999f), "f"(-0.001f)); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 12345.f) out[0] = s;
}

template <int MODE>
void run(const char* name, int sms, int warps, double ops_per_iter_elem) {
  float* out; CK(cudaMalloc(&out, 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 2048;
  k<MODE><<<sms, warps * 32>>>(out, 8);
  CK(cudaEventRecord(e0));
  k<MODE><<<sms, warps * 32>>>(out, iters);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double per_sm_per_clk = 16.0 * iters * warps * 32 * ops_per_iter_elem / (ms * 1e-3 * 1.9e9);
  printf("%-22s warps/SM %2d: %6.1f ops/clk/SM (at 1.9 GHz)\n", name, warps, per_sm_per_clk);
  CK(cudaFree(out));
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int w : {4, 8, 16, 32}) {
    run<0>("ex2", sms, w, 1.0);
    run<1>("ffma", sms, w, 1.0);
    run<2>("ex2 + 3 ffma (ex2/clk)", sms, w, 1.0);
  }
  return 0;
}
