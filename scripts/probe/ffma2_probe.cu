// FFMA2 (fma.rn.f32x2) latency / issue-rate probe for sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE, int ILP>
__global__ void k(float* out, long long* cyc, int iters) {
  float s = 1.0f + threadIdx.x * 1e-7f;
  uint64_t a2[ILP]; float a1[ILP];
  for (int j = 0; j < ILP; ++j) { a2[j] = pk(s + j, s - j); a1[j] = s + j; }
  uint64_t b2 = pk(0.999f, 1.001f), c2 = pk(1e-3f, -1e-3f);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      if (MODE == 2) a2[j] = fma2(a2[j], b2, c2);
      else a1[j] = fma1(a1[j], 0.999f, 1e-3f);
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int j = 0; j < ILP; ++j) { acc += a1[j]; float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a2[j])); acc += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int ILP> void run(int threads, const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 4 * 148 * 1024); cudaMalloc(&cyc, 8);
  int iters = 4096;
  k<MODE, ILP><<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
  k<MODE, ILP><<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / iters;
  printf("%-8s ILP=%d warps/SM=%2d: %.2f cycles per iteration -> %.2f cycles/instr/warp, %.2f warp-instr/clk/SM\n", name, ILP, threads / 32, per, per / ILP,
         (double)ILP * (threads / 32) / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, 1>(32, "FFMA");  run<2, 1>(32, "FFMA2");
  run<1, 8>(128, "FFMA"); run<2, 8>(128, "FFMA2");
  run<1, 8>(512, "FFMA"); run<2, 8>(512, "FFMA2");
  run<1, 2>(352, "FFMA"); run<2, 2>(352, "FFMA2");
  return 0;
}
