// Microbenchmark: tcgen05.ld throughput on sm_100a -- the TMEM read rate the policy kernels' epilogues lean on (K2 reads four
// 128 x 128 fp32 accumulators per 128-row tile).  One CTA per SM allocates all 512 TMEM columns; W warps (4, 8 or 16: a warp may
// only read the lane quarter w % 4) each issue ITERS x tcgen05.ld.32x32b.{x16, x32, x64} over their quarter, with the loads of one
// step in flight together (one tcgen05.wait::ld per step).  Reports bytes per clock per SM, from clock64 around the loop, and the
// same with a tanh.approx per element between load and wait-free use (the epilogue's real mix).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../fpv-drone-rl-agent_b200/csrc -o tmem_ld tmem_ld.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc05.cuh"
#define ITERS 512

__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&r)[64]) {
  uint32_t(&a)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
  uint32_t(&b)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
  tc05::tmem_ld32(taddr, a);
  tc05::tmem_ld32(taddr + 32, b);
}

// MODE 0: x16 per step, 1: x32 per step, 2: two x32 per step (64 columns in flight); TANH: apply tanh.approx to every element
template <int MODE, bool TANH>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk) {
  __shared__ uint32_t slot;
  const uint32_t warp = threadIdx.x >> 5;
  if (warp == 0) tc05::tmem_alloc(&slot, 512);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t base = slot + (((warp & 3) * 32u) << 16);
  constexpr int COLS = MODE == 0 ? 16 : (MODE == 1 ? 32 : 64);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    const uint32_t col = (uint32_t)((i * COLS + (warp >> 2) * 64) & 511 & ~(COLS - 1));
    uint32_t r[64];
    if (MODE == 0) { uint32_t(&a)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]); tc05::tmem_ld16(base + col, a); }
    else if (MODE == 1) { uint32_t(&a)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]); tc05::tmem_ld32(base + col, a); }
    else ld64(base + (col & 448), r);
    tc05::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < COLS; ++j) {
      float v = __uint_as_float(r[j]);
      if (TANH) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(v) : "f"(v));
      acc += v;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc(slot, 512);
}

template <int MODE, bool TANH> void run(const char* name, int warps, float* out, long long* clk, int sms) {
  k<MODE, TANH><<<sms, warps * 32>>>(out, clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[256];
  cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < sms; ++i) mean += (double)h[i];
  mean /= sms;
  const int cols = MODE == 0 ? 16 : (MODE == 1 ? 32 : 64);
  const double bytes = (double)warps * 32 * cols * 4 * ITERS;  // per CTA = per SM
  printf("%-34s %2d warps: %7.1f B/clk/SM  (%6.2f elements/clk/SM, %8.0f clk)\n", name, warps, bytes / mean, bytes / 4 / mean, mean);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float* out;
  long long* clk;
  cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 512);
  cudaMalloc(&clk, sizeof(long long) * 256);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  for (int w : {4, 8, 16}) {
    run<0, false>("tcgen05.ld 32x32b.x16", w, out, clk, p.multiProcessorCount);
    run<1, false>("tcgen05.ld 32x32b.x32", w, out, clk, p.multiProcessorCount);
    run<2, false>("tcgen05.ld 2 x 32x32b.x32 in flight", w, out, clk, p.multiProcessorCount);
    run<1, true>("tcgen05.ld x32 + tanh per element", w, out, clk, p.multiProcessorCount);
    run<2, true>("tcgen05.ld 2 x x32 + tanh", w, out, clk, p.multiProcessorCount);
  }
  return 0;
}
