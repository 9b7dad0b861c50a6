// ppo_update_kernels.cu -- K5: the PPO update (SB3 PPO.train, reached by train_hover.py:60 model.learn) for the
// [128,128] tanh actor-critic, hand-written for sm_100a.  C-ABI: include/ppo_b200.h (ppo_update_*).
//
// One optimiser step on a minibatch of 128-row tiles is four launches:
//   1. update_prepare_kernel   minibatch size, advantage mean / std (SB3 normalises advantages per minibatch)
//   2. update_fwdbwd_kernel    forward + loss + backward of BOTH networks on the tensor cores (tcgen05, bf16 operands, fp32
//                              accumulation in TMEM); per-CTA partial gradients to a scratch buffer
//   3. update_reduce_kernel    sum of the CTA partials -> flat fp32 gradient (+ its squared norm)
//      (multi-GPU: the caller all-reduces the flat gradient here -- the only collective of the path -- then ppo_update_grad_norm)
//   4. update_adam_kernel      global-norm clip + Adam on the flat fp32 master parameters, and the bf16 re-pack of the
//                              weights into the PpoPolicy buffers the rollout's forward kernel reads
//
// update_fwdbwd_kernel, per 128-sample tile and network (policy pass over all of the CTA's tiles, then value pass):
//   forward   Z1 = X W1^T -> H1 = tanh(Z1 + b1) -> Z2 = H1 W2^T -> H2 = tanh(Z2 + b2) -> out = H2 W3^T          (3 MMA groups)
//   loss      per row: Gaussian log-prob, ratio, clipped surrogate / value MSE -> dOut [128 x 16] (bf16)
//   backward  dH2 = dOut W3,  dW3 += H2^T dOut;   dZ2 = dH2 (1 - H2^2)
//             dH1 = dZ2 W2,   dW2 += H1^T dZ2,  db2 += 1^T dZ2;   dZ1 = dH1 (1 - H1^2)
//             dW1 += dZ1^T X   (X carries a constant-1 column, so its row of dW1 is db1)
// No transposed copies exist: the interleaved no-swizzle layout [K/8][rows][8] that makes an activation tile a K-major
// A operand (K = features) is, read with the MN-major bit of the instruction descriptor, also the transposed operand
// (K = samples) the weight-gradient GEMMs need, and a weight matrix W[out][in] staged K-major for the forward is the
// MN-major B operand of dH = dZ W.  (Canonical layouts: K-major ((8,m),(T,2)):((1T,SBO),(1,LBO)); MN-major
// ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)) -- the same 128-byte core matrices with the roles of LBO and SBO exchanged.)
// Weight gradients accumulate in TMEM across all tiles of a pass (fp32) and leave the SM once per pass.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>

#include "../../include/ppo_b200.h"
#include "../../include/quadx_b200.h"
#include "qx_internal.h"
#include "tc05.cuh"

namespace ppo_upd {

using namespace tc05;

constexpr int kHid = PPO_HIDDEN, kIn = PPO_IN_PAD, kHead = PPO_HEAD_PAD;
constexpr int kThreads = 256;

// instruction descriptor, kind::f16, D = fp32, A = B = bf16, with the major-ness bits (15: A, 16: B; 1 = MN-major)
__host__ __device__ constexpr uint32_t idesc(uint32_t m, uint32_t n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// a wait that cannot hang the GPU: a protocol error in the MMA / barrier chain traps instead of spinning forever
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

// ---------------------------------------------------------------------------
// flat parameter vector (fp32 master copy, gradient, Adam moments): PyTorch module order of ActorCritic
//   pi1.W [H][od], pi1.b [H], pi2.W [H][H], pi2.b [H], mu.W [ad][H], mu.b [ad],
//   vf1.W, vf1.b, vf2.W, vf2.b, v.W [1][H], v.b [1], log_std [ad]
// ---------------------------------------------------------------------------
struct Layout {
  int od, ad;
  int w1p, b1p, w2p, b2p, wmu, bmu, w1v, b1v, w2v, b2v, wv, bv, ls, total;
};
__host__ __device__ inline Layout make_layout(int od, int ad) {
  Layout L;
  L.od = od; L.ad = ad;
  int o = 0;
  L.w1p = o; o += kHid * od; L.b1p = o; o += kHid; L.w2p = o; o += kHid * kHid; L.b2p = o; o += kHid;
  L.wmu = o; o += ad * kHid; L.bmu = o; o += ad;
  L.w1v = o; o += kHid * od; L.b1v = o; o += kHid; L.w2v = o; o += kHid * kHid; L.b2v = o; o += kHid;
  L.wv = o; o += kHid; L.bv = o; o += 1; L.ls = o; o += ad;
  L.total = o;
  return L;
}

// workspace header (device): written by the prepare kernel, read by the others
struct Header {
  float inv_b;        // 1 / number of valid rows in the minibatch
  float adv_mean, adv_inv_std;
  float norm2;        // squared L2 norm of the flat gradient
  unsigned int ticket, pad;
  double part[3];     // prepare: sum 1, sum adv, sum adv^2 (fp64)
  unsigned long long step;  // Adam step count
  float lr_scale_m1;  // learning-rate schedule: the Adam kernel steps with lr * (1 + lr_scale_m1), so a zero-filled workspace means lr
  // SB3's target_kl early stop (PPO.train: "if approx_kl_div > 1.5 * target_kl: continue_training = False; break"), on the device so
  // that it works per minibatch inside a captured epoch: the forward/backward kernel sums (ratio - 1) - log ratio and the row count
  // of THIS minibatch, the reduce kernel appends both to the flat gradient (slots [total], [total + 1], so a gradient all-reduce
  // carries them and every rank takes the same decision), and the Adam kernel skips its update -- and every later one until the
  // host re-arms -- once the mean exceeds 1.5 x target_kl.
  float mb_k3, mb_n;
  float target_kl;        // 0 = off (zero-filled workspace)
  unsigned int stopped;   // latched by the Adam kernel
  unsigned int skipped;   // optimiser steps skipped since the last re-arm
  unsigned int ls_floor_on;
  float ls_floor;         // lower bound of log_std after the Adam step (optional)
};
constexpr size_t kHeaderBytes = 256;
static_assert(sizeof(Header) <= kHeaderBytes, "header");

struct UpdArgs {
  PpoPolicy p;
  const float* obs;       // [n_rows, od]  normalised observations as the policy saw them
  const float* actions;   // [n_rows, ad]
  const float* old_logp;  // [n_rows]
  const float* adv;       // [n_rows]
  const float* ret;       // [n_rows]
  const int32_t* tiles;   // [n_tiles] indices of the 128-row tiles of this minibatch
  int32_t n_tiles;
  int64_t n_rows;
  float clip_range, vf_coef, ent_coef;
  int32_t normalize_adv;
  Header* hdr;
  float* scratch;         // [gridDim.x][total] per-CTA partial gradients
  float* loss_stats;      // [8] accumulated: pg, vf, -(logp - old), clipfrac, (ratio - 1) - log ratio, n_rows, 0, 0
  float* logp_out;        // non-null: forward pass of the actor only, log pi(a | s) of every row written here, nothing else touched
};

// ---------------------------------------------------------------------------
// 1. minibatch statistics
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) update_prepare_kernel(const float* __restrict__ adv, const int32_t* __restrict__ tiles, int n_tiles,
                                                             int64_t n_rows, int normalize, Header* __restrict__ hdr) {
  __shared__ double sh[8][3];
  double c = 0.0, s = 0.0, q = 0.0;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t base = (int64_t)tiles[t] * 128;
    if (threadIdx.x < 128) {
      const int64_t r = base + threadIdx.x;
      if (r < n_rows) { const double a = (double)adv[r]; c += 1.0; s += a; q += a * a; }
    }
  }
  for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, o); s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[w][0] = c; sh[w][1] = s; sh[w][2] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { sh[0][0] += sh[k][0]; sh[0][1] += sh[k][1]; sh[0][2] += sh[k][2]; }
    atomicAdd(&hdr->part[0], sh[0][0]); atomicAdd(&hdr->part[1], sh[0][1]); atomicAdd(&hdr->part[2], sh[0][2]);
    __threadfence();
    if (atomicAdd(&hdr->ticket, 1u) == gridDim.x - 1) {  // last block: finalise, re-arm for the next launch
      __threadfence();
      const double n = *reinterpret_cast<volatile double*>(&hdr->part[0]), sa = *reinterpret_cast<volatile double*>(&hdr->part[1]),
                   qa = *reinterpret_cast<volatile double*>(&hdr->part[2]);
      const double mean = n > 0 ? sa / n : 0.0;
      const double var = n > 1 ? fmax((qa - n * mean * mean) / (n - 1.0), 0.0) : 0.0;  // torch.std(): unbiased
      hdr->inv_b = n > 0 ? (float)(1.0 / n) : 0.f;
      hdr->adv_mean = normalize ? (float)mean : 0.f;
      hdr->adv_inv_std = normalize ? (float)(1.0 / (sqrt(var) + 1e-8)) : 1.f;
      hdr->norm2 = 0.f;
      hdr->mb_k3 = 0.f; hdr->mb_n = 0.f;
      hdr->part[0] = hdr->part[1] = hdr->part[2] = 0.0;
      hdr->ticket = 0u;
    }
  }
}

// ---------------------------------------------------------------------------
// 2. forward + loss + backward on the tensor cores
// ---------------------------------------------------------------------------
// shared memory map (bytes)
constexpr uint32_t kSmW1 = 0;                                   // [4][256][16 B]   rows 0..127 pi, 128..255 vf
constexpr uint32_t kSmW2p = kSmW1 + 2 * kHid * kIn * 2;         // [16][128][16 B]
constexpr uint32_t kSmW2v = kSmW2p + kHid * kHid * 2;
constexpr uint32_t kSmW3 = kSmW2v + kHid * kHid * 2;            // [32][16][16 B]   K chunks 0..15 pi half, 16..31 vf half
constexpr uint32_t kSmX = kSmW3 + kHead * 2 * kHid * 2;         // [4][128][16 B]   bf16 observations, column od = 1
constexpr uint32_t kSmH1 = kSmX + 128 * kIn * 2;                // [16][128][16 B]  H1, later dZ1
constexpr uint32_t kSmH2 = kSmH1 + 128 * kHid * 2;              // [16][128][16 B]  H2, later dZ2
constexpr uint32_t kSmD3 = kSmH2 + 128 * kHid * 2;              // [2][128][16 B]   dOut
constexpr uint32_t kSmOnes = kSmD3 + 128 * kHead * 2;           // [2][128][16 B]   all ones
constexpr uint32_t kSmB1 = kSmOnes + 128 * 16 * 2;              // 256 f32
constexpr uint32_t kSmB2 = kSmB1 + 2 * kHid * 4;
constexpr uint32_t kSmB3 = kSmB2 + 2 * kHid * 4;                // 16 f32
constexpr uint32_t kSmAcc = kSmB3 + kHead * 4;                  // 32 f32: [0..15] d bias of the heads (dmu.., dv), [16..19] dlog_std, [24..29] loss stats
constexpr uint32_t kSmTotal = kSmAcc + 32 * 4;
static_assert(kSmTotal <= 227 * 1024, "shared memory");
// TMEM columns
constexpr uint32_t kTmAcc = 0, kTmOut = 128, kTmDW2 = 144, kTmDB2 = 272, kTmDW1 = 400, kTmDW3 = 432, kTmCols = 512;

__device__ __forceinline__ float tanh_fast(float x) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// H = tanh(acc + bias) -> bf16 operand chunks (this thread: row `row`, 64 columns from col0)
__device__ __forceinline__ void epi_tanh(uint32_t tacc_row, const float* __restrict__ bias, uint8_t* sH, uint32_t row, int col0) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int c0 = col0 + 32 * half;
    uint32_t r[32];
    tmem_ld32(tacc_row + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 v;
      const float* b = bias + c0 + 8 * q;
      v.x = pack_bf16x2(tanh_fast(__uint_as_float(r[8 * q + 0]) + b[0]), tanh_fast(__uint_as_float(r[8 * q + 1]) + b[1]));
      v.y = pack_bf16x2(tanh_fast(__uint_as_float(r[8 * q + 2]) + b[2]), tanh_fast(__uint_as_float(r[8 * q + 3]) + b[3]));
      v.z = pack_bf16x2(tanh_fast(__uint_as_float(r[8 * q + 4]) + b[4]), tanh_fast(__uint_as_float(r[8 * q + 5]) + b[5]));
      v.w = pack_bf16x2(tanh_fast(__uint_as_float(r[8 * q + 6]) + b[6]), tanh_fast(__uint_as_float(r[8 * q + 7]) + b[7]));
      *reinterpret_cast<uint4*>(sH + chunk_off(row, (uint32_t)(c0 >> 3) + q, 128)) = v;
    }
  }
}

// dZ = dH (1 - H^2), in place over H (bf16): acc holds dH (fp32)
__device__ __forceinline__ void epi_dtanh(uint32_t tacc_row, uint8_t* sH, uint32_t row, int col0) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int c0 = col0 + 32 * half;
    uint32_t r[32];
    tmem_ld32(tacc_row + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4* ph = reinterpret_cast<uint4*>(sH + chunk_off(row, (uint32_t)(c0 >> 3) + q, 128));
      const uint4 h = *ph;
      uint4 v;
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
      uint32_t out[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float h0 = bf16_lo(hw[j]), h1 = bf16_hi(hw[j]);
        const float d0 = __uint_as_float(r[8 * q + 2 * j]) * fmaf(-h0, h0, 1.f), d1 = __uint_as_float(r[8 * q + 2 * j + 1]) * fmaf(-h1, h1, 1.f);
        out[j] = pack_bf16x2(d0, d1);
      }
      v.x = out[0]; v.y = out[1]; v.z = out[2]; v.w = out[3];
      *ph = v;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) update_fwdbwd_kernel(const UpdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t erow = (warp & 3) * 32 + lane;  // accumulator row of this thread in the epilogues
  const int ecol0 = (int)(warp >> 2) * 64;       // and its 64 columns
  const uint32_t xrow = tid & 127, xh = tid >> 7;
  const int od = a.p.obs_dim, ad = a.p.act_dim;
  const Layout L = make_layout(od, ad);

  stage_weight(smem + kSmW1, (const __nv_bfloat16*)a.p.w1, 2 * kHid, kIn, tid, kThreads);
  stage_weight(smem + kSmW2p, (const __nv_bfloat16*)a.p.w2p, kHid, kHid, tid, kThreads);
  stage_weight(smem + kSmW2v, (const __nv_bfloat16*)a.p.w2v, kHid, kHid, tid, kThreads);
  stage_weight(smem + kSmW3, (const __nv_bfloat16*)a.p.w3, kHead, 2 * kHid, tid, kThreads);
  float* sB1 = reinterpret_cast<float*>(smem + kSmB1);
  float* sB2 = reinterpret_cast<float*>(smem + kSmB2);
  float* sB3 = reinterpret_cast<float*>(smem + kSmB3);
  float* sAcc = reinterpret_cast<float*>(smem + kSmAcc);
  sB1[tid] = a.p.b1[tid]; sB2[tid] = a.p.b2[tid];
  if (tid < kHead) sB3[tid] = a.p.b3[tid];
  if (tid < 32) sAcc[tid] = 0.f;
  for (uint32_t i = tid; i < 128 * 16 * 2 / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem + kSmOnes)[i] = 0x3f803f80u;  // bf16 1.0 x2
  if (warp == 0) tmem_alloc(&tmem_slot, kTmCols);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tbase = tmem_slot;
  const uint32_t trow = tbase + (((warp & 3) * 32u) << 16);  // this warp's lane quarter
  const uint32_t sW1 = smem_u32(smem + kSmW1), sW2p = smem_u32(smem + kSmW2p), sW2v = smem_u32(smem + kSmW2v), sW3 = smem_u32(smem + kSmW3);
  const uint32_t sX = smem_u32(smem + kSmX), sH1 = smem_u32(smem + kSmH1), sH2 = smem_u32(smem + kSmH2), sD3 = smem_u32(smem + kSmD3),
                 sOnes = smem_u32(smem + kSmOnes);
  // K-major steps: LBO = distance between the K chunks, SBO = 128 (next 8 rows).  MN-major: LBO = 128 (next 8 K rows), SBO = distance
  // between the 8-wide MN groups.
  constexpr uint32_t ACT = 128 * 16, W1S = 2 * kHid * 16, W2S = kHid * 16, W3S = kHead * 16;
  const uint32_t id_h = idesc(128, kHid, false, false), id_o = idesc(128, kHead, false, false);
  const uint32_t id_dh2 = idesc(128, kHid, false, true);    // dH = dZ . W   : A K-major, B = W as MN-major
  const uint32_t id_dw = idesc(128, kHid, true, true);      // dW = H^T dZ  : both MN-major
  const uint32_t id_db = idesc(128, kHid, false, true);     // db = 1 . dZ
  const uint32_t id_dw3 = idesc(128, kHead, true, true);    // dW3 = H2^T dOut
  const uint32_t id_dw1 = idesc(128, kIn, true, true);      // dW1 = dZ1^T X
  uint32_t phase = 0;
  const float inv_b = a.hdr->inv_b, adv_mean = a.hdr->adv_mean, adv_inv_std = a.hdr->adv_inv_std;
  float* my_scratch = a.scratch + (size_t)blockIdx.x * L.total;
  // every entry of the CTA's scratch slot is written below (each CTA owns at least one tile of both passes)

  float st_pg = 0.f, st_vf = 0.f, st_kl = 0.f, st_cf = 0.f, st_k3 = 0.f, st_n = 0.f;  // per-thread loss statistics (rows of warps 0..3)
  const bool fwd_only = a.logp_out != nullptr;

  for (int net = 0; net < (fwd_only ? 1 : 2); ++net) {
    const uint32_t sW1n = sW1 + (net ? kHid * 16 : 0), sW2n = net ? sW2v : sW2p, sW3n = sW3 + (net ? 16 * W3S : 0);
    const float* b1 = sB1 + (net ? kHid : 0);
    const float* b2 = sB2 + (net ? kHid : 0);
    int iter = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++iter) {
      const int64_t row0 = (int64_t)a.tiles[t] * 128;
      // ---- X tile: fp32 global -> bf16 operand chunks, column od = 1 (bias row of dW1); 2 threads per row, 16 columns each
      {
        const int64_t r = row0 + xrow;
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = 0.f;
        if (r < a.n_rows) {
          const float* src = a.obs + r * od + 16 * xh;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (16 * (int)xh + j < od) x[j] = __ldg(src + j);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (16 * (int)xh + j == od) x[j] = 1.f;
        }
#pragma unroll
        for (int qq = 0; qq < 2; ++qq)
          *reinterpret_cast<uint4*>(smem + kSmX + chunk_off(xrow, 2 * xh + qq, 128)) =
              make_uint4(pack_bf16x2(x[8 * qq], x[8 * qq + 1]), pack_bf16x2(x[8 * qq + 2], x[8 * qq + 3]), pack_bf16x2(x[8 * qq + 4], x[8 * qq + 5]),
                         pack_bf16x2(x[8 * qq + 6], x[8 * qq + 7]));
      }
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      // ---- L1: acc = X W1^T
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < kIn / 16; ++ks)
          mma_bf16(tbase + kTmAcc, make_desc(sX + ks * 2 * ACT, ACT, 128), make_desc(sW1n + ks * 2 * W1S, W1S, 128), id_h, ks > 0);
        mma_commit(&bar);
      }
      mbar_wait_bounded(&bar, phase); phase ^= 1;
      fence_after_sync();
      epi_tanh(trow + kTmAcc, b1, smem + kSmH1, erow, ecol0);
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      // ---- L2: acc = H1 W2^T
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ++ks)
          mma_bf16(tbase + kTmAcc, make_desc(sH1 + ks * 2 * ACT, ACT, 128), make_desc(sW2n + ks * 2 * W2S, W2S, 128), id_h, ks > 0);
        mma_commit(&bar);
      }
      mbar_wait_bounded(&bar, phase); phase ^= 1;
      fence_after_sync();
      epi_tanh(trow + kTmAcc, b2, smem + kSmH2, erow, ecol0);
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      // ---- L3: out = H2 W3^T (this network's half of the head matrix)
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ++ks)
          mma_bf16(tbase + kTmOut, make_desc(sH2 + ks * 2 * ACT, ACT, 128), make_desc(sW3n + ks * 2 * W3S, W3S, 128), id_o, ks > 0);
        mma_commit(&bar);
      }
      mbar_wait_bounded(&bar, phase); phase ^= 1;
      fence_after_sync();
      // ---- loss: one thread per row (warps 0..3)
      if (warp < 4) {
        uint32_t r[16];
        tmem_ld16(trow + kTmOut, r);
        tmem_ld_wait();
        const int64_t row = row0 + erow;
        const bool valid = row < a.n_rows;
        float d[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) d[j] = 0.f;
        if (net == 0) {
          float dls[4] = {0.f, 0.f, 0.f, 0.f};
          if (valid) {
            float logp = 0.f, z[4], isd[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              z[j] = 0.f; isd[j] = 0.f;
              if (j < ad) {
                const float ls = __ldg(a.p.log_std + j), mean = __uint_as_float(r[j]) + sB3[j];
                isd[j] = __expf(-ls);
                z[j] = (__ldg(a.actions + row * ad + j) - mean) * isd[j];
                logp += -0.5f * z[j] * z[j] - ls - 0.91893853320467f;
              }
            }
            if (fwd_only) a.logp_out[row] = logp;
            const float advn = fwd_only ? 0.f : (__ldg(a.adv + row) - adv_mean) * adv_inv_std;
            const float lr = fwd_only ? 0.f : logp - __ldg(a.old_logp + row), ratio = __expf(lr);
            const float lo = 1.f - a.clip_range, hi = 1.f + a.clip_range;
            const bool clipped = (advn > 0.f && ratio > hi) || (advn < 0.f && ratio < lo);
            const float g = clipped ? 0.f : -advn * ratio * inv_b;  // d loss / d logp
            st_pg += fmaxf(-advn * ratio, -advn * fminf(fmaxf(ratio, lo), hi));
            st_kl += -lr; st_cf += (fabsf(ratio - 1.f) > a.clip_range) ? 1.f : 0.f; st_k3 += (ratio - 1.f) - lr; st_n += 1.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < ad) { d[j] = g * z[j] * isd[j]; dls[j] = g * (z[j] * z[j] - 1.f); }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float s0 = d[j], s1 = dls[j];
            for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
            if (lane == 0 && j < ad) { atomicAdd(&sAcc[j], s0); atomicAdd(&sAcc[16 + j], s1); }
          }
        } else {
          if (valid) {
            const float v = __uint_as_float(r[ad]) + sB3[ad], diff = v - __ldg(a.ret + row);
            st_vf += diff * diff;
            d[ad] = a.vf_coef * 2.f * diff * inv_b;
          }
          float s0 = d[ad];
          for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          if (lane == 0) atomicAdd(&sAcc[ad], s0);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
          *reinterpret_cast<uint4*>(smem + kSmD3 + chunk_off(erow, q, 128)) =
              make_uint4(pack_bf16x2(d[8 * q], d[8 * q + 1]), pack_bf16x2(d[8 * q + 2], d[8 * q + 3]), pack_bf16x2(d[8 * q + 4], d[8 * q + 5]),
                         pack_bf16x2(d[8 * q + 6], d[8 * q + 7]));
      }
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      if (fwd_only) { fence_after_sync(); continue; }
      // ---- B1: dH2 = dOut W3 (K = 16 heads);  dW3 += H2^T dOut (K = 128 samples)
      if (tid == 0) {
        fence_after_sync();
        mma_bf16(tbase + kTmAcc, make_desc(sD3, ACT, 128), make_desc(sW3n, 128, W3S), id_dh2, false);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_bf16(tbase + kTmDW3, make_desc(sH2 + ks * 256, 128, ACT), make_desc(sD3 + ks * 256, 128, ACT), id_dw3, iter > 0 || ks > 0);
        mma_commit(&bar);
      }
      mbar_wait_bounded(&bar, phase); phase ^= 1;
      fence_after_sync();
      epi_dtanh(trow + kTmAcc, smem + kSmH2, erow, ecol0);  // H2 <- dZ2
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      // ---- B2: dH1 = dZ2 W2;  dW2 += H1^T dZ2;  db2 += 1 dZ2
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_bf16(tbase + kTmAcc, make_desc(sH2 + ks * 2 * ACT, ACT, 128), make_desc(sW2n + ks * 256, 128, W2S), id_dh2, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_bf16(tbase + kTmDW2, make_desc(sH1 + ks * 256, 128, ACT), make_desc(sH2 + ks * 256, 128, ACT), id_dw, iter > 0 || ks > 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_bf16(tbase + kTmDB2, make_desc(sOnes, ACT, 128), make_desc(sH2 + ks * 256, 128, ACT), id_db, iter > 0 || ks > 0);
        mma_commit(&bar);
      }
      mbar_wait_bounded(&bar, phase); phase ^= 1;
      fence_after_sync();
      epi_dtanh(trow + kTmAcc, smem + kSmH1, erow, ecol0);  // H1 <- dZ1
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      // ---- B3: dW1 += dZ1^T X
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_bf16(tbase + kTmDW1, make_desc(sH1 + ks * 256, 128, ACT), make_desc(sX + ks * 256, 128, ACT), id_dw1, iter > 0 || ks > 0);
        mma_commit(&bar);
      }
      mbar_wait_bounded(&bar, phase); phase ^= 1;
      fence_after_sync();
    }
    // ---- flush this pass's weight gradients (TMEM, fp32) to the CTA's scratch slot, in the flat parameter layout
    if (iter > 0 && warp < 4 && !fwd_only) {
      const int oW2 = net ? L.w2v : L.w2p, oB2 = net ? L.b2v : L.b2p, oW1 = net ? L.w1v : L.w1p, oB1 = net ? L.b1v : L.b1p;
      // dW2: TMEM row = input feature i, column = output feature o  ->  W2[o][i]
#pragma unroll 1
      for (int c0 = 0; c0 < kHid; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + kTmDW2 + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) my_scratch[oW2 + (c0 + j) * kHid + erow] = __uint_as_float(r[j]);
      }
      // db2: every row of the ones-product holds the column sums; row 0 (warp 0, lane 0) writes them
#pragma unroll 1
      for (int c0 = 0; c0 < kHid; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + kTmDB2 + c0, r);
        tmem_ld_wait();
        if (erow == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) my_scratch[oB2 + c0 + j] = __uint_as_float(r[j]);
        }
      }
      // dW1: row = hidden feature f, column = input k  ->  W1[f][k] (k < od), b1[f] (k == od)
      {
        uint32_t r[32];
        tmem_ld32(trow + kTmDW1, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          if (k < od) my_scratch[oW1 + (int)erow * od + k] = __uint_as_float(r[k]);
          else if (k == od) my_scratch[oB1 + erow] = __uint_as_float(r[k]);
        }
      }
      // dW3: row = hidden feature f, column = head j  ->  Wmu[j][f] / Wv[0][f]
      {
        uint32_t r[16];
        tmem_ld16(trow + kTmDW3, r);
        tmem_ld_wait();
        if (net == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < ad) my_scratch[L.wmu + j * kHid + erow] = __uint_as_float(r[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 5; ++j)
            if (j == ad) my_scratch[L.wv + erow] = __uint_as_float(r[j]);
        }
      }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  // ---- head biases, log_std, loss statistics
  if (fwd_only) {
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, kTmCols);
    return;
  }
  if (warp < 4) {
    float v[6] = {st_pg, st_vf, st_kl, st_cf, st_k3, st_n};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      if (lane == 0) atomicAdd(&sAcc[24 + k], v[k]);
    }
  }
  __syncthreads();
  if (tid < (uint32_t)ad) {
    my_scratch[L.bmu + tid] = sAcc[tid];
    my_scratch[L.ls + tid] = sAcc[16 + tid] + (blockIdx.x == 0 ? -a.ent_coef : 0.f);  // d(-ent_coef * mean entropy) / d log_std = -ent_coef
  }
  if (tid == 0) my_scratch[L.bv] = sAcc[ad];
  if (tid < 6 && a.loss_stats) atomicAdd(&a.loss_stats[tid], sAcc[24 + tid]);
  if (tid == 6) atomicAdd(&a.hdr->mb_k3, sAcc[24 + 4]);
  if (tid == 7) atomicAdd(&a.hdr->mb_n, sAcc[24 + 5]);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, kTmCols);
}

// ---------------------------------------------------------------------------
// 3. sum of the CTA partials -> flat gradient, and its squared norm
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) update_reduce_kernel(const float* __restrict__ scratch, int n_ctas, int total, float* __restrict__ grad,
                                                            Header* __restrict__ hdr) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  float s = 0.f;
  if (p < total) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int c = 0;
    for (; c + 4 <= n_ctas; c += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += scratch[(size_t)(c + k) * total + p];
    }
    for (; c < n_ctas; ++c) acc[0] += scratch[(size_t)c * total + p];
    s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    grad[p] = s;
  }
  if (p == 0 && hdr->target_kl > 0.f) { grad[total] = hdr->mb_k3; grad[total + 1] = hdr->mb_n; }  // see Header
  float q = s * s;
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += sh[k];
    atomicAdd(&hdr->norm2, t);
  }
}

// squared norm of an (all-reduced) gradient scaled by `scale`
__global__ void __launch_bounds__(256) update_norm_kernel(const float* __restrict__ grad, int total, float scale, Header* __restrict__ hdr) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  const float g = p < total ? grad[p] * scale : 0.f;
  float q = g * g;
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += sh[k];
    atomicAdd(&hdr->norm2, t);
  }
}
__global__ void update_zero_norm_kernel(Header* hdr) { hdr->norm2 = 0.f; }

// ---------------------------------------------------------------------------
// 4. clip + Adam + bf16 re-pack
// ---------------------------------------------------------------------------
struct AdamArgs {
  float* params;
  const float* grad;
  float* m;
  float* v;
  Header* hdr;
  int total;
  float lr, beta1, beta2, eps, max_grad_norm, grad_scale;
  PpoPolicy pk;  // destination of the re-pack (written)
};

__global__ void __launch_bounds__(256) update_adam_kernel(const AdamArgs a) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  const unsigned long long t = a.hdr->step + 1ull;
  const float tkl = a.hdr->target_kl;
  const bool stop = a.hdr->stopped != 0u || (tkl > 0.f && a.grad[a.total] > 1.5f * tkl * fmaxf(a.grad[a.total + 1], 1.f));
  if (p < a.total && !stop) {
    const float norm = sqrtf(a.hdr->norm2);
    const float clip = a.max_grad_norm > 0.f ? fminf(1.f, a.max_grad_norm / (norm + 1e-6f)) : 1.f;  // torch clip_grad_norm_
    const float g = a.grad[p] * a.grad_scale * clip;
    const float m = a.beta1 * a.m[p] + (1.f - a.beta1) * g;
    const float v = a.beta2 * a.v[p] + (1.f - a.beta2) * g * g;
    a.m[p] = m; a.v[p] = v;
    const double bc1 = 1.0 - pow((double)a.beta1, (double)t), bc2 = 1.0 - pow((double)a.beta2, (double)t);
    const float denom = sqrtf(v) / (float)sqrt(bc2) + a.eps;
    float w = a.params[p] - (float)((double)a.lr * (1.0 + (double)a.hdr->lr_scale_m1) / bc1) * (m / denom);
    if (a.hdr->ls_floor_on && p >= make_layout(a.pk.obs_dim, a.pk.act_dim).ls) w = fmaxf(w, a.hdr->ls_floor);
    a.params[p] = w;
    // ---- re-pack into the bf16 / fp32 buffers of the forward kernels
    const Layout L = make_layout(a.pk.obs_dim, a.pk.act_dim);
    const int od = L.od, ad = L.ad;
    __nv_bfloat16* w1 = (__nv_bfloat16*)a.pk.w1; __nv_bfloat16* w2p = (__nv_bfloat16*)a.pk.w2p; __nv_bfloat16* w2v = (__nv_bfloat16*)a.pk.w2v;
    __nv_bfloat16* w3 = (__nv_bfloat16*)a.pk.w3;
    float* b1 = (float*)a.pk.b1; float* b2 = (float*)a.pk.b2; float* b3 = (float*)a.pk.b3; float* ls = (float*)a.pk.log_std;
    const __nv_bfloat16 wb = __float2bfloat16_rn(w);
    if (p < L.b1p) { const int f = (p - L.w1p) / od, k = (p - L.w1p) % od; w1[f * kIn + k] = wb; }
    else if (p < L.w2p) b1[p - L.b1p] = w;
    else if (p < L.b2p) w2p[p - L.w2p] = wb;
    else if (p < L.wmu) b2[p - L.b2p] = w;
    else if (p < L.bmu) { const int j = (p - L.wmu) / kHid, f = (p - L.wmu) % kHid; w3[j * 2 * kHid + f] = wb; }
    else if (p < L.w1v) b3[p - L.bmu] = w;
    else if (p < L.b1v) { const int f = (p - L.w1v) / od, k = (p - L.w1v) % od; w1[(kHid + f) * kIn + k] = wb; }
    else if (p < L.w2v) b1[kHid + p - L.b1v] = w;
    else if (p < L.b2v) w2v[p - L.w2v] = wb;
    else if (p < L.wv) b2[kHid + p - L.b2v] = w;
    else if (p < L.bv) w3[ad * 2 * kHid + kHid + (p - L.wv)] = wb;
    else if (p < L.ls) b3[ad] = w;
    else ls[p - L.ls] = w;
  }
  // the step counter advances once per launch: the last block to finish bumps it (every block has read it by then)
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&a.hdr->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    if (stop) { a.hdr->stopped = 1u; a.hdr->skipped += 1u; }
    else a.hdr->step = t;
    a.hdr->ticket = 0u;
  }
}

// test hook: D[128, n] = A^T-or-A times B^T-or-B through MN-major / K-major descriptors
//   a_mn = 0: A given as [128][k] (K-major)   a_mn = 1: A given as [k][128]  (MN-major: M contiguous)
//   b_mn = 0: B given as [n][k]               b_mn = 1: B given as [k][n]
__global__ void __launch_bounds__(128) test_gemm_mn_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
                                                           int n, int k, int a_mn, int b_mn) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * k * 2;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  // K-major operand [rows][k]: chunk (row, kc) at kc * rows * 16 + row * 16.  MN-major operand [k][mn]: the 16-byte chunk holding
  // (k, 8 g .. 8 g + 7) at g * k * 16 + k * 16 -- i.e. the same interleaved layout with the roles of the two indices exchanged.
  if (!a_mn) stage_weight(sA, A, 128, k, tid, 128);
  else stage_weight(sA, A, k, 128, tid, 128);
  if (!b_mn) stage_weight(sB, B, n, k, tid, 128);
  else stage_weight(sB, B, k, n, tid, 128);
  fence_async_smem();
  uint32_t ncols = 32;
  while ((int)ncols < n) ncols <<= 1;
  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint32_t id = idesc(128, n, a_mn != 0, b_mn != 0);
    for (int ks = 0; ks < k / 16; ++ks) {
      const uint64_t da = !a_mn ? make_desc(smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128) : make_desc(smem_u32(sA) + ks * 256, 128, k * 16);
      const uint64_t db = !b_mn ? make_desc(smem_u32(sB) + ks * 2 * n * 16, n * 16, 128) : make_desc(smem_u32(sB) + ks * 256, 128, k * 16);
      mma_bf16(tbase, da, db, id, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait_bounded(&bar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tbase + ((warp * 32u) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)tid * n + c0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, ncols);
}

}  // namespace ppo_upd

static int ufail(int code, const char* msg) { return qx_fail(code, "%s", msg); }

extern "C" int ppo_test_gemm_mn(const void* a, const void* b, float* d, int32_t n, int32_t k, int32_t a_mn, int32_t b_mn, void* stream) {
  if (!a || !b || !d || n < 16 || n > 256 || n % 16 || k < 16 || k % 16) return ufail(QX_EINVAL, "ppo_test_gemm_mn: bad arguments");
  const size_t smem = (size_t)(128 + n) * k * 2;
  if (smem > 200 * 1024) return ufail(QX_EINVAL, "ppo_test_gemm_mn: tile too large");
  cudaFuncSetAttribute(ppo_upd::test_gemm_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ppo_upd::test_gemm_mn_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, d, n, k, a_mn, b_mn);
  return cudaGetLastError() == cudaSuccess ? QX_OK : ufail(QX_ECUDA, "ppo_test_gemm_mn: launch failed");
}

extern "C" int32_t ppo_update_num_params(int32_t obs_dim, int32_t act_dim) {
  if (obs_dim < 1 || obs_dim >= PPO_IN_PAD || act_dim < 1 || act_dim > 4) return -1;
  return ppo_upd::make_layout(obs_dim, act_dim).total;
}

static int upd_sms(int* out) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    return ufail(QX_ECUDA, "ppo_update: cannot query the device");
  *out = sms;
  return QX_OK;
}

// per-device one-time opt-in to ~170 KB of dynamic shared memory
static int upd_smem_optin() {
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return ufail(QX_ECUDA, "ppo_update: cannot query the current device");
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    if (cudaFuncSetAttribute(ppo_upd::update_fwdbwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ppo_upd::kSmTotal) != cudaSuccess)
      return ufail(QX_ECUDA, "ppo_update: cannot reserve shared memory");
    attr_set[dev].store(true, std::memory_order_release);
  }
  return QX_OK;
}

extern "C" int64_t ppo_update_workspace_bytes(int32_t obs_dim, int32_t act_dim) {
  const int32_t n = ppo_update_num_params(obs_dim, act_dim);
  if (n < 0) return -1;
  int sms = 0;
  if (upd_sms(&sms)) sms = 148;
  return (int64_t)ppo_upd::kHeaderBytes + (int64_t)sizeof(float) * n * sms;
}

extern "C" int ppo_update_minibatch(const PpoPolicy* p, const float* obs, const float* actions, const float* old_logp, const float* adv,
                                    const float* ret, const int32_t* tiles_dev, int32_t n_tiles, int64_t n_rows, float clip_range, float vf_coef,
                                    float ent_coef, int32_t normalize_adv, float* grad_out, float* loss_stats, void* workspace, void* stream) {
  if (!p || !obs || !actions || !old_logp || !adv || !ret || !tiles_dev || n_tiles <= 0 || n_rows <= 0 || !grad_out || !workspace)
    return ufail(QX_EINVAL, "ppo_update_minibatch: bad arguments");
  const int32_t total = ppo_update_num_params(p->obs_dim, p->act_dim);
  if (total < 0) return ufail(QX_EINVAL, "ppo_update_minibatch: unsupported observation / action width");
  int sms = 0;
  if (int rc = upd_sms(&sms)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  ppo_upd::Header* hdr = (ppo_upd::Header*)workspace;
  float* scratch = (float*)((uint8_t*)workspace + ppo_upd::kHeaderBytes);
  const int grid = n_tiles < sms ? n_tiles : sms;
  ppo_upd::update_prepare_kernel<<<grid < 64 ? grid : 64, 256, 0, s>>>(adv, tiles_dev, n_tiles, n_rows, normalize_adv, hdr);
  if (int rc = upd_smem_optin()) return rc;
  ppo_upd::UpdArgs a{};
  a.p = *p; a.obs = obs; a.actions = actions; a.old_logp = old_logp; a.adv = adv; a.ret = ret; a.tiles = tiles_dev; a.n_tiles = n_tiles;
  a.n_rows = n_rows; a.clip_range = clip_range; a.vf_coef = vf_coef; a.ent_coef = ent_coef; a.normalize_adv = normalize_adv; a.hdr = hdr;
  a.scratch = scratch; a.loss_stats = loss_stats;
  ppo_upd::update_fwdbwd_kernel<<<grid, ppo_upd::kThreads, ppo_upd::kSmTotal, s>>>(a);
  ppo_upd::update_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(scratch, grid, total, grad_out, hdr);
  return cudaGetLastError() == cudaSuccess ? QX_OK : ufail(QX_ECUDA, "ppo_update_minibatch: launch failed");
}

extern "C" int ppo_update_recompute_logp(const PpoPolicy* p, const float* obs, const float* actions, const int32_t* tiles_dev, int32_t n_tiles,
                                         int64_t n_rows, float* logp_out, void* workspace, void* stream) {
  if (!p || !obs || !actions || !tiles_dev || n_tiles <= 0 || n_rows <= 0 || !logp_out || !workspace)
    return ufail(QX_EINVAL, "ppo_update_recompute_logp: bad arguments");
  if (ppo_update_num_params(p->obs_dim, p->act_dim) < 0) return ufail(QX_EINVAL, "ppo_update_recompute_logp: unsupported observation / action width");
  int sms = 0;
  if (int rc = upd_sms(&sms)) return rc;
  if (int rc = upd_smem_optin()) return rc;
  ppo_upd::UpdArgs a{};
  a.p = *p; a.obs = obs; a.actions = actions; a.tiles = tiles_dev; a.n_tiles = n_tiles; a.n_rows = n_rows; a.clip_range = 0.2f;
  a.hdr = (ppo_upd::Header*)workspace; a.scratch = (float*)((uint8_t*)workspace + ppo_upd::kHeaderBytes); a.logp_out = logp_out;
  const int grid = n_tiles < sms ? n_tiles : sms;
  ppo_upd::update_fwdbwd_kernel<<<grid, ppo_upd::kThreads, ppo_upd::kSmTotal, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? QX_OK : ufail(QX_ECUDA, "ppo_update_recompute_logp: launch failed");
}

extern "C" int ppo_update_grad_norm(const float* grad, int32_t n_params, float grad_scale, void* workspace, void* stream) {
  if (!grad || n_params <= 0 || !workspace) return ufail(QX_EINVAL, "ppo_update_grad_norm: bad arguments");
  ppo_upd::Header* hdr = (ppo_upd::Header*)workspace;
  ppo_upd::update_zero_norm_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hdr);
  ppo_upd::update_norm_kernel<<<(n_params + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grad, n_params, grad_scale, hdr);
  return cudaGetLastError() == cudaSuccess ? QX_OK : ufail(QX_ECUDA, "ppo_update_grad_norm: launch failed");
}

extern "C" int ppo_update_adam(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t n_params, float lr, float beta1,
                               float beta2, float eps, float max_grad_norm, float grad_scale, const PpoPolicy* packed_out, void* workspace,
                               void* stream) {
  if (!params || !grad || !exp_avg || !exp_avg_sq || !packed_out || !workspace) return ufail(QX_EINVAL, "ppo_update_adam: bad arguments");
  if (n_params != ppo_update_num_params(packed_out->obs_dim, packed_out->act_dim)) return ufail(QX_EINVAL, "ppo_update_adam: parameter count does not match the policy");
  ppo_upd::AdamArgs a{};
  a.params = params; a.grad = grad; a.m = exp_avg; a.v = exp_avg_sq; a.hdr = (ppo_upd::Header*)workspace; a.total = n_params;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_grad_norm = max_grad_norm; a.grad_scale = grad_scale; a.pk = *packed_out;
  ppo_upd::update_adam_kernel<<<(n_params + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? QX_OK : ufail(QX_ECUDA, "ppo_update_adam: launch failed");
}

extern "C" int ppo_update_set_lr_scale(void* workspace, float scale, void* stream) {
  if (!workspace || !(scale >= 0.f)) return ufail(QX_EINVAL, "ppo_update_set_lr_scale: bad arguments");
  ppo_upd::Header* hdr = (ppo_upd::Header*)workspace;
  const float v = scale - 1.0f;
  if (cudaMemcpyAsync(&hdr->lr_scale_m1, &v, sizeof(v), cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess)
    return ufail(QX_ECUDA, "ppo_update_set_lr_scale: copy failed");
  cudaStreamSynchronize((cudaStream_t)stream);  // v is on this frame
  return QX_OK;
}

extern "C" int ppo_update_kl_stop(void* workspace, float target_kl, int32_t* stopped_out, int32_t* skipped_out, void* stream) {
  if (!workspace) return ufail(QX_EINVAL, "ppo_update_kl_stop: null workspace");
  ppo_upd::Header* hdr = (ppo_upd::Header*)workspace;
  cudaStream_t s = (cudaStream_t)stream;
  if (stopped_out || skipped_out) {
    unsigned int v[2] = {0u, 0u};
    if (cudaMemcpyAsync(v, &hdr->stopped, sizeof(v), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
      return ufail(QX_ECUDA, "ppo_update_kl_stop: copy failed");
    if (stopped_out) *stopped_out = (int32_t)v[0];
    if (skipped_out) *skipped_out = (int32_t)v[1];
  }
  if (target_kl >= 0.f) {  // (re-)arm: set the threshold, clear the latch and the counter
    struct { float tkl; unsigned int stopped, skipped; } v = {target_kl, 0u, 0u};
    static_assert(offsetof(ppo_upd::Header, stopped) == offsetof(ppo_upd::Header, target_kl) + 4 && offsetof(ppo_upd::Header, skipped) == offsetof(ppo_upd::Header, target_kl) + 8, "layout");
    if (cudaMemcpyAsync(&hdr->target_kl, &v, sizeof(v), cudaMemcpyHostToDevice, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
      return ufail(QX_ECUDA, "ppo_update_kl_stop: copy failed");
  }
  return QX_OK;
}

extern "C" int ppo_update_set_log_std_floor(void* workspace, int32_t on, float floor, void* stream) {
  if (!workspace) return ufail(QX_EINVAL, "ppo_update_set_log_std_floor: null workspace");
  ppo_upd::Header* hdr = (ppo_upd::Header*)workspace;
  struct { unsigned int on; float v; } x = {on ? 1u : 0u, floor};
  static_assert(offsetof(ppo_upd::Header, ls_floor) == offsetof(ppo_upd::Header, ls_floor_on) + 4, "layout");
  if (cudaMemcpyAsync(&hdr->ls_floor_on, &x, sizeof(x), cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess || cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
    return ufail(QX_ECUDA, "ppo_update_set_log_std_floor: copy failed");
  return QX_OK;
}

extern "C" int ppo_update_step_count(void* workspace, int64_t set_to, int64_t* out, void* stream) {
  if (!workspace) return ufail(QX_EINVAL, "ppo_update_step_count: null workspace");
  ppo_upd::Header* hdr = (ppo_upd::Header*)workspace;
  if (set_to >= 0) {
    unsigned long long v = (unsigned long long)set_to;
    if (cudaMemcpyAsync(&hdr->step, &v, sizeof(v), cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess) return ufail(QX_ECUDA, "ppo_update_step_count: copy failed");
    cudaStreamSynchronize((cudaStream_t)stream);
  }
  if (out) {
    unsigned long long v = 0;
    cudaStreamSynchronize((cudaStream_t)stream);
    if (cudaMemcpy(&v, &hdr->step, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return ufail(QX_ECUDA, "ppo_update_step_count: copy failed");
    *out = (int64_t)v;
  }
  return QX_OK;
}
