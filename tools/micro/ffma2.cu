// Microbenchmark: issue throughput of scalar FFMA vs packed fma.rn.f32x2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_ffma(float* out, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&d)) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return d;
}
__global__ void k_ffma2(float* out, float a, float b) {
  float t = threadIdx.x;
  float2 x0 = make_float2(t, t + 1), x1 = make_float2(t + 2, t + 3), x2 = make_float2(t + 4, t + 5), x3 = make_float2(t + 6, t + 7);
  float2 A = make_float2(a, a), B = make_float2(b, b);
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    x0 = ffma2(x0, A, B); x1 = ffma2(x1, A, B); x2 = ffma2(x2, A, B); x3 = ffma2(x3, A, B);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y;
}
// mixed: 4 scalar FFMA + 2 packed + 4 integer ops per iteration, to see whether packing frees issue slots for other pipes
__global__ void k_mix_scalar(float* out, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7;
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    u0 = (u0 ^ u1) + 0x9E3779B9u; u1 = (u1 ^ u2) + 0xBB67AE85u; u2 = (u2 ^ u3) + 0x12345u; u3 = (u3 ^ u0) + 0x54321u;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (float)(u0 ^ u1 ^ u2 ^ u3);
}
__global__ void k_mix_packed(float* out, float a, float b) {
  float t = threadIdx.x;
  float2 x0 = make_float2(t, t + 1), x1 = make_float2(t + 2, t + 3), x2 = make_float2(t + 4, t + 5), x3 = make_float2(t + 6, t + 7);
  float2 A = make_float2(a, a), B = make_float2(b, b);
  unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7;
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    x0 = ffma2(x0, A, B); x1 = ffma2(x1, A, B); x2 = ffma2(x2, A, B); x3 = ffma2(x3, A, B);
    u0 = (u0 ^ u1) + 0x9E3779B9u; u1 = (u1 ^ u2) + 0xBB67AE85u; u2 = (u2 ^ u3) + 0x12345u; u3 = (u3 ^ u0) + 0x54321u;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y + (float)(u0 ^ u1 ^ u2 ^ u3);
}
template <typename F> float timeit(F f) {
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(s); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e); return ms / 5;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 512 * 4);
  const int grid = 148 * 4, block = 512;  // 16 warps/SMSP
  double fl = (double)grid * block * ITERS * 8 * 2;
  float t1 = timeit([&] { k_ffma<<<grid, block>>>(out, 1.0001f, 0.5f); });
  float t2 = timeit([&] { k_ffma2<<<grid, block>>>(out, 1.0001f, 0.5f); });
  float t3 = timeit([&] { k_mix_scalar<<<grid, block>>>(out, 1.0001f, 0.5f); });
  float t4 = timeit([&] { k_mix_packed<<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("scalar FFMA : %.3f ms  %.1f TFLOP/s\n", t1, fl / t1 / 1e9);
  printf("packed FFMA2: %.3f ms  %.1f TFLOP/s\n", t2, fl / t2 / 1e9);
  printf("mix scalar  : %.3f ms\nmix packed  : %.3f ms\n", t3, t4);
  return 0;
}
