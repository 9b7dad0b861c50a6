// Microbenchmark: MUFU throughput on sm_100a for the ops the K1 / K2 kernels lean on -- tanh.approx.f32 (K2's bound),
// ex2, rsqrt, sin -- in results per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define OP_TANH 0
#define OP_EX2 1
#define OP_RSQ 2
#define OP_SIN 3
#define OP_TANH_BF16X2 4
template <int OP> __device__ __forceinline__ float op(float x) {
  float r;
  if (OP == OP_TANH) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  else if (OP == OP_EX2) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  else if (OP == OP_RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  else if (OP == OP_SIN) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  else {
    unsigned u = __float_as_uint(x), v;
    asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(v) : "r"(u));
    r = __uint_as_float(v);
  }
  return r;
}
template <int OP> __global__ void k(float* out) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = 0.1f * (threadIdx.x + j) + 0.3f;
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = op<OP>(x[j]);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, float* out, int sms, double mhz) {
  const int blocks = sms * 4, threads = 256;
  k<OP><<<blocks, threads>>>(out);
  cudaDeviceSynchronize();
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s);
  for (int i = 0; i < 5; ++i) k<OP><<<blocks, threads>>>(out);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e); ms /= 5;
  const double ops = (double)blocks * threads * 8.0 * ITERS;
  printf("%-16s %8.3f ms  %7.2f Gop/s  %6.2f results/clk/SM at %.0f MHz%s\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / (mhz * 1e6) / sms, mhz,
         OP == OP_TANH_BF16X2 ? "  (x2 elements per result)" : "");
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  float* out; cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 4 * 256);
  run<OP_TANH>("tanh.approx.f32", out, p.multiProcessorCount, mhz);
  run<OP_EX2>("ex2.approx.f32", out, p.multiProcessorCount, mhz);
  run<OP_RSQ>("rsqrt.approx.f32", out, p.multiProcessorCount, mhz);
  run<OP_SIN>("sin.approx.f32", out, p.multiProcessorCount, mhz);
  run<OP_TANH_BF16X2>("tanh.bf16x2", out, p.multiProcessorCount, mhz);
  return 0;
}
