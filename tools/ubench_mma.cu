// Micro-benchmark: legacy mma.sync.m16n8k16 bf16 throughput on sm_100a vs warps per SM and independent accumulator chains.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int CH>
__global__ void k(float* out, int iters) {
  float c[CH][4];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.f) out[0] = s;
}

template <int CH>
void run(int sms, int warps_per_sm) {
  float* out; CK(cudaMalloc(&out, 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 4096;
  const int threads = 128, ctas = sms * warps_per_sm / 4;
  k<CH><<<ctas, threads>>>(out, 16);
  CK(cudaEventRecord(e0));
  k<CH><<<ctas, threads>>>(out, iters);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double flop = 4096.0 * CH * iters * (double)ctas * 4;
  printf("warps/SM %2d  chains %d : %7.1f TFLOP/s  (%.1f cycles per mma per warp at 1.9 GHz)\n", warps_per_sm, CH, flop / ms / 1e9,
         ms * 1e-3 * 1.9e9 / (iters * CH));
  CK(cudaFree(out));
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int w : {4, 8, 16, 32}) { run<1>(sms, w); run<4>(sms, w); run<8>(sms, w); run<16>(sms, w); }
  return 0;
}
