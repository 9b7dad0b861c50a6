// Microbenchmark (sm_100a): issue rate of scalar FFMA / packed FFMA2 forms, alone and mixed with ALU-pipe and MUFU work,
// in warp-instructions per clock per SM sub-partition (clock64 inside the kernel, so DVFS does not matter).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_issue ffma2_issue.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
typedef unsigned long long u64;
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// KIND 0: FFMA reg,reg,reg   1: FFMA reg,imm,reg   2: FFMA2 reg,reg,reg   3: FFMA2 reg,imm,reg   4: FFMA2 reg,R.F32,reg
// 5: 8 FFMA + 8 FMNMX (alu)  6: 4 FFMA2 + 8 FMNMX   7: 4 FFMA2 + 4 FFMA   8: 4 FFMA2 + 2 MUFU   9: 8 FFMA + 2 MUFU
// 10: 4 FFMA2 + 8 FMNMX + 2 MUFU   11: 8 FFMA + 8 FMNMX + 2 MUFU   12: 4 FFMA2 + 8 IMAD(lo)
template <int KIND>
__global__ void k(float* out, u64* cyc, float a, float b, int n_iter) {
  const float t = threadIdx.x * 1e-3f;
  float x0 = t, x1 = t + 1, x2 = t + 2, x3 = t + 3, x4 = t + 4, x5 = t + 5, x6 = t + 6, x7 = t + 7;
  float2 p0 = make_float2(x0, x1), p1 = make_float2(x2, x3), p2 = make_float2(x4, x5), p3 = make_float2(x6, x7);
  float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 1.0001f);
  float m0 = t, m1 = t + .5f, m2 = t + .25f, m3 = t + .125f, m4 = t, m5 = t + .5f, m6 = t + .25f, m7 = t + .125f;
  float u0 = t + 0.1f, u1 = t + 0.2f;
  unsigned i0 = threadIdx.x, i1 = i0 * 3, i2 = i0 * 5, i3 = i0 * 7, i4 = i0 + 9, i5 = i0 * 11, i6 = i0 * 13, i7 = i0 * 17;
  __syncthreads();
  const u64 c0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n_iter; ++i) {
    if (KIND == 0 || KIND == 5 || KIND == 9 || KIND == 11) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    if (KIND == 1) {
      x0 = fmaf(x0, 1.0001f, b); x1 = fmaf(x1, 1.0001f, b); x2 = fmaf(x2, 1.0001f, b); x3 = fmaf(x3, 1.0001f, b);
      x4 = fmaf(x4, 1.0001f, b); x5 = fmaf(x5, 1.0001f, b); x6 = fmaf(x6, 1.0001f, b); x7 = fmaf(x7, 1.0001f, b);
    }
    if (KIND == 2 || KIND == 6 || KIND == 7 || KIND == 8 || KIND == 10 || KIND == 12) {
      p0 = f2fma(p0, A, B); p1 = f2fma(p1, A, B); p2 = f2fma(p2, A, B); p3 = f2fma(p3, A, B);
    }
    if (KIND == 3) {
      const float2 I = make_float2(1.0001f, 1.0001f);
      p0 = f2fma(p0, I, B); p1 = f2fma(p1, I, B); p2 = f2fma(p2, I, B); p3 = f2fma(p3, I, B);
    }
    if (KIND == 4) {
      const float2 S = make_float2(a, a);
      p0 = f2fma(p0, S, B); p1 = f2fma(p1, S, B); p2 = f2fma(p2, S, B); p3 = f2fma(p3, S, B);
    }
    if (KIND == 7) { x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b); }
    if (KIND == 5 || KIND == 6 || KIND == 10 || KIND == 11) {
      m0 = fminf(m0, m1 + 0.f) ; m1 = fmaxf(m1, m2); m2 = fminf(m2, m3); m3 = fmaxf(m3, m4);
      m4 = fminf(m4, m5); m5 = fmaxf(m5, m6); m6 = fminf(m6, m7); m7 = fmaxf(m7, m0);
    }
    if (KIND == 8 || KIND == 9 || KIND == 10 || KIND == 11) { u0 = __sinf(u0); u1 = __cosf(u1); }
    if (KIND == 12) {
      i0 = i0 * 0xD2511F53u + i1; i1 = i1 * 0xCD9E8D57u + i2; i2 = i2 * 0x9E3779B9u + i3; i3 = i3 * 0xBB67AE85u + i4;
      i4 = i4 * 0xD2511F53u + i5; i5 = i5 * 0xCD9E8D57u + i6; i6 = i6 * 0x9E3779B9u + i7; i7 = i7 * 0xBB67AE85u + i0;
    }
  }
  const u64 c1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y +
                                               m0 + m1 + m2 + m3 + m4 + m5 + m6 + m7 + u0 + u1 + (float)(i0 ^ i1 ^ i2 ^ i3 ^ i4 ^ i5 ^ i6 ^ i7);
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int KIND>
void run(const char* name, int n_inst, float* out, u64* cyc) {
  for (int warps_per_smsp : {1, 2, 4, 8}) {
    const int block = 128 * warps_per_smsp;  // one block per SM
    k<KIND><<<148, block>>>(out, cyc, 1.0001f, 0.5f, ITERS);
    k<KIND><<<148, block>>>(out, cyc, 1.0001f, 0.5f, ITERS);
    cudaDeviceSynchronize();
    u64 h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    // warp-instructions issued per SMSP = warps_per_smsp * ITERS * n_inst
    printf("%-44s warps/SMSP %d : %.3f warp-inst/clk/SMSP  (%.2f clk per iteration per warp)\n", name, warps_per_smsp,
           warps_per_smsp * (double)ITERS * n_inst / avg, avg / ITERS);
  }
}

int main() {
  float* out; u64* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("FFMA  r,r,r            x8", 8, out, cyc);
  run<1>("FFMA  r,imm,r          x8", 8, out, cyc);
  run<2>("FFMA2 r,r,r            x4", 4, out, cyc);
  run<3>("FFMA2 r,imm,r          x4", 4, out, cyc);
  run<4>("FFMA2 r,R.F32,r        x4", 4, out, cyc);
  run<5>("8 FFMA + 8 FMNMX", 16, out, cyc);
  run<6>("4 FFMA2 + 8 FMNMX", 12, out, cyc);
  run<7>("4 FFMA2 + 4 FFMA", 8, out, cyc);
  run<8>("4 FFMA2 + 2 MUFU", 6, out, cyc);
  run<9>("8 FFMA + 2 MUFU", 10, out, cyc);
  run<10>("4 FFMA2 + 8 FMNMX + 2 MUFU", 14, out, cyc);
  run<11>("8 FFMA + 8 FMNMX + 2 MUFU", 18, out, cyc);
  run<12>("4 FFMA2 + 8 IMAD", 12, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
