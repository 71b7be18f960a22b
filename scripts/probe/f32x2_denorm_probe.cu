// Do FFMA2 / FMUL2 / FADD2 (fma/mul/add.rn.f32x2) treat subnormals like the scalar IEEE instructions?
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__global__ void k(const float* x, float* out, int n, unsigned* mism) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = x[3 * i], b = x[3 * i + 1], c = x[3 * i + 2];
  u64 A = pk(a, 1.5f), B = pk(b, 2.5f), C = pk(c, -0.75f), R;
  float lo, hi;
  float m1, m2, m3;
  asm("mul.rn.f32 %0, %1, %2;" : "=f"(m1) : "f"(a), "f"(b));
  asm("add.rn.f32 %0, %1, %2;" : "=f"(m2) : "f"(a), "f"(c));
  asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(m3) : "f"(a), "f"(b), "f"(c));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(R) : "l"(A), "l"(B)); up(R, lo, hi);
  if (__float_as_uint(lo) != __float_as_uint(m1)) { atomicAdd(&mism[0], 1u); if (i < 64) { out[4 * i] = a; out[4 * i + 1] = b; out[4 * i + 2] = lo; out[4 * i + 3] = m1; } }
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(R) : "l"(A), "l"(C)); up(R, lo, hi);
  if (__float_as_uint(lo) != __float_as_uint(m2)) atomicAdd(&mism[1], 1u);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(R) : "l"(A), "l"(B), "l"(C)); up(R, lo, hi);
  if (__float_as_uint(lo) != __float_as_uint(m3)) atomicAdd(&mism[2], 1u);
}
int main() {
  const int n = 1 << 22;
  float* h = new float[3 * n];
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s; };
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < 3; ++j) {
      uint32_t r = rnd();
      uint32_t u;
      int mode = i & 3;
      if (mode == 0) u = r & 0x807fffffu;                                   // subnormal operands
      else if (mode == 1) u = (r & 0x807fffffu) | ((20u + (rnd() % 40u)) << 23);  // tiny normals: products underflow
      else if (mode == 2) u = (r & 0x807fffffu) | ((100u + (rnd() % 60u)) << 23); // ordinary magnitudes
      else u = (r & 0x807fffffu) | ((60u + (rnd() % 10u)) << 23);                // ~1e-20: squares are subnormal
      memcpy(&h[3 * i + j], &u, 4);
    }
  }
  float *dx, *dout; unsigned* dm;
  cudaMalloc(&dx, 12ull * n); cudaMalloc(&dout, 4 * 4 * 64); cudaMalloc(&dm, 12);
  cudaMemcpy(dx, h, 12ull * n, cudaMemcpyHostToDevice); cudaMemset(dm, 0, 12); cudaMemset(dout, 0, 1024);
  k<<<n / 256, 256>>>(dx, dout, n, dm);
  unsigned m[3]; float o[256];
  cudaMemcpy(m, dm, 12, cudaMemcpyDeviceToHost); cudaMemcpy(o, dout, 1024, cudaMemcpyDeviceToHost);
  printf("mismatches vs scalar IEEE over %d cases: mul %u  add %u  fma %u\n", n, m[0], m[1], m[2]);
  int shown = 0;
  for (int i = 0; i < 64 && shown < 4; ++i) if (o[4 * i] != 0 || o[4 * i + 1] != 0) { printf("  mul %g * %g: packed %g scalar %g\n", o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]); ++shown; }
  return 0;
}
