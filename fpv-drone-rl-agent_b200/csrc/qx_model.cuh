// qx_model.cuh -- fp32 device model of one QuadX drone + the hover / yaw env layer.
//
// Everything here runs in registers of ONE thread per env.  What is computed
// follows, step by step,
//   * /root/reference/simulation/hover.py:224-358 (obs, reward, termination, step)
//   * hover.py:72-113 (reset), yaw.py:57-149 (yaw task)
//   * the PyFlyt 0.21.0 QuadX model (mode 0, plus the outer loops of modes -1..7) and pybullet's multibody integrator
//     as restated in SURVEY.md section 9 (third-party, parity unpinned),
//     parameterised by cf2x.yaml:1-54 and cf2x.urdf:10-68.
// It is written from that description for this hardware, not translated from
// the CPU oracle: single precision, MUFU intrinsics, no libm slow paths inside
// the sub-step loop, Philox counters instead of a stateful generator.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qx {

// ---------------------------------------------------------------------------
// kernel-side constants, derived once on the host from QxConfig
// ---------------------------------------------------------------------------
struct DevConfig {
  int32_t task, n_sub_step, n_sub_reset, ctrl_every, max_steps, floor_grace, render, auto_reset, noise,
      state_stale, gyro, obs_dim, act_dim;
  float h, lag_alpha, one_m_alpha, noise_ratio, noise_k, thrust_k, pwm_idle;  // noise_k = -2 ln2 noise_ratio^2
  float torque_k[4], mx[4], my[4], map[16];
  float drag_c, drag_pqr, kp[3], kiT[3], kd_T[3], lim[3];
  float inv_mass, g, I[3], invI[3], vmax, floor_z;
  float hI[3], hm, hg, hh, hh2, ndrag_c, ndrag_pqr, kq1, kq2;  // h/I, h/m, h g, h/2, (h/2)^2, -drag, series coefficients * h/2
  float gk[3];                                                 // gyroscopic term: -(h / I_k) (I_(k+2) - I_(k+1)), or 0 when off
  float cam_sd, cam_cd, inv_tan, res, half_res, inv_half_res, inv_res2, cam_near, cam_off[3], margin, panel[12];
  float inv_agent_dt, dome2, floor_thr, target_area, target_ratio, act_scale[3];
  float start_pos[3], start_rpy[3], spawn_thr, spawn_pos_noise, spawn_yaw_noise;
  uint32_t seed_lo, seed_hi, env_lo, env_hi;  // env id of local env 0
  // PyFlyt flight modes (cascade instantiation only): outer-loop gains as kp[w] ki*T[w] kd/T[w] lim[w]
  int32_t flight_mode, need_euler;
  float thrust_scale, thrust_bias;
  float att[12], vel[8], lpos[8], zpos[4], zvel[4];
};

}  // namespace qx
// The reference's own parameter set (hover.py / cf2x.yaml / cf2x.urdf literals) as compile-time constants: a second
// instantiation of the step kernel overwrites the model fields of its DevConfig copy with these literals, so they
// become immediates instead of ~33 constant-bank loads per Aviary.step and a few dozen live registers.  It is selected
// at qx_create only when derive() produced exactly these bits; any other configuration runs the generic kernel.
#include "qx_ref_constants.cuh"
namespace qx {

enum : uint32_t {
  F_CONTACT = 1u,
  F_TERM = 2u,
  F_TRUNC = 4u,
  F_OOB = 8u,
  F_ONFLOOR = 16u,
  F_LOWZ = 32u,
};
enum : uint32_t { STREAM_STEP = 0, STREAM_RESET = 1, STREAM_SPAWN = 2 };

// The 44 carried words of one env (QX_STATE_WORDS); plane p, lane l = word 4p+l.
struct Env {
  float px, py, pz;           // world position
  float qx, qy, qz, qw;       // body->world quaternion (x,y,z,w)
  float vx, vy, vz;           // world linear velocity
  float wx, wy, wz;           // BODY-frame angular velocity (see physics_substep)
  float thr[4];               // motor throttle (PyFlyt Motors.throttle)
  float pi[3], pe[3];         // rate PID integral, previous error
  float swb[3], svb[3];       // Aviary.state rows 0 and 2 (body rates / velocity snapshot)
  float peul[3];              // previous_ang_pos, hover.py:354
  float pa[4];                // prev_action, hover.py:357
  float pcx, pcy, parea, pratio;  // hover.py:270-272
  int32_t step_count;         // hover.py:356
  uint32_t rng_ctr;           // env steps since creation (Philox counter word 2)
  float ep_ret;               // Monitor: running episode return
  uint32_t flags;
  // transient (not stored): pose part of the Aviary.state snapshot
  float sqx, sqy, sqz, sqw, spx, spy, spz;
  // cascade instantiation only (planes 11..16; spx/spy/spz are carried there too): (integral, previous error) of
  // ang_pos 0/3, lin_vel 6/8, lin_pos 10/12, z_vel 14/15, z_pos 16/17
  float cp[18];
};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ void load_env(Env& e, const float4* __restrict__ st, int64_t n, int64_t i) {
  float4 a = ldg4(st + 0 * n + i), b = ldg4(st + 1 * n + i), c = ldg4(st + 2 * n + i), d = ldg4(st + 3 * n + i);
  float4 f = ldg4(st + 4 * n + i), g = ldg4(st + 5 * n + i), h = ldg4(st + 6 * n + i), k = ldg4(st + 7 * n + i);
  float4 l = ldg4(st + 8 * n + i), m = ldg4(st + 9 * n + i), o = ldg4(st + 10 * n + i);
  e.px = a.x; e.py = a.y; e.pz = a.z; e.qx = a.w;
  e.qy = b.x; e.qz = b.y; e.qw = b.z; e.vx = b.w;
  e.vy = c.x; e.vz = c.y; e.wx = c.z; e.wy = c.w;
  e.wz = d.x; e.thr[0] = d.y; e.thr[1] = d.z; e.thr[2] = d.w;
  e.thr[3] = f.x; e.pi[0] = f.y; e.pi[1] = f.z; e.pi[2] = f.w;
  e.pe[0] = g.x; e.pe[1] = g.y; e.pe[2] = g.z; e.swb[0] = g.w;
  e.swb[1] = h.x; e.swb[2] = h.y; e.svb[0] = h.z; e.svb[1] = h.w;
  e.svb[2] = k.x; e.peul[0] = k.y; e.peul[1] = k.z; e.peul[2] = k.w;
  e.pa[0] = l.x; e.pa[1] = l.y; e.pa[2] = l.z; e.pa[3] = l.w;
  e.pcx = m.x; e.pcy = m.y; e.parea = m.z; e.pratio = m.w;
  e.step_count = __float_as_int(o.x); e.rng_ctr = __float_as_uint(o.y); e.ep_ret = o.z; e.flags = __float_as_uint(o.w);
  e.sqx = e.qx; e.sqy = e.qy; e.sqz = e.qz; e.sqw = e.qw; e.spx = e.px; e.spy = e.py; e.spz = e.pz;
}

__device__ __forceinline__ void store_env(const Env& e, float4* __restrict__ st, int64_t n, int64_t i) {
  st[0 * n + i] = make_float4(e.px, e.py, e.pz, e.qx);
  st[1 * n + i] = make_float4(e.qy, e.qz, e.qw, e.vx);
  st[2 * n + i] = make_float4(e.vy, e.vz, e.wx, e.wy);
  st[3 * n + i] = make_float4(e.wz, e.thr[0], e.thr[1], e.thr[2]);
  st[4 * n + i] = make_float4(e.thr[3], e.pi[0], e.pi[1], e.pi[2]);
  st[5 * n + i] = make_float4(e.pe[0], e.pe[1], e.pe[2], e.swb[0]);
  st[6 * n + i] = make_float4(e.swb[1], e.swb[2], e.svb[0], e.svb[1]);
  st[7 * n + i] = make_float4(e.svb[2], e.peul[0], e.peul[1], e.peul[2]);
  st[8 * n + i] = make_float4(e.pa[0], e.pa[1], e.pa[2], e.pa[3]);
  st[9 * n + i] = make_float4(e.pcx, e.pcy, e.parea, e.pratio);
  st[10 * n + i] = make_float4(__int_as_float(e.step_count), __uint_as_float(e.rng_ctr), e.ep_ret, __uint_as_float(e.flags));
}

// Flight modes != 0: six more planes -- the outer loops' PID memory and the position row of the Aviary.state
// snapshot, which those loops read at the first control update of the next step.
constexpr int kBasePlanes = 11, kCascadePlanes = 17;
__device__ __forceinline__ void load_cascade(Env& e, const float4* __restrict__ st, int64_t n, int64_t i) {
  float4 v[6];
#pragma unroll
  for (int p = 0; p < 6; ++p) v[p] = ldg4(st + (int64_t)(kBasePlanes + p) * n + i);
  const float* w = reinterpret_cast<const float*>(v);
#pragma unroll
  for (int k = 0; k < 18; ++k) e.cp[k] = w[k];
  e.spx = w[18]; e.spy = w[19]; e.spz = w[20];
}
__device__ __forceinline__ void store_cascade(const Env& e, float4* __restrict__ st, int64_t n, int64_t i) {
  st[(int64_t)(kBasePlanes + 0) * n + i] = make_float4(e.cp[0], e.cp[1], e.cp[2], e.cp[3]);
  st[(int64_t)(kBasePlanes + 1) * n + i] = make_float4(e.cp[4], e.cp[5], e.cp[6], e.cp[7]);
  st[(int64_t)(kBasePlanes + 2) * n + i] = make_float4(e.cp[8], e.cp[9], e.cp[10], e.cp[11]);
  st[(int64_t)(kBasePlanes + 3) * n + i] = make_float4(e.cp[12], e.cp[13], e.cp[14], e.cp[15]);
  st[(int64_t)(kBasePlanes + 4) * n + i] = make_float4(e.cp[16], e.cp[17], e.spx, e.spy);
  st[(int64_t)(kBasePlanes + 5) * n + i] = make_float4(e.spz, 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------
// Philox4x32-10: same integer stream as oracle/quadx_model.py:philox4x32_10
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// MUFU-only approximations (1-2 ulp): no Newton refinement, no slow paths
__device__ __forceinline__ float frcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fsqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float frsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ((x >> 9) + 0.5) * 2^-23 without an int->float conversion (reset pose noise)
__device__ __forceinline__ float u01(uint32_t x) { return __uint_as_float((x >> 9) | 0x3f800000u) - 0.99999994f; }
// 16-bit field -> (k + 0.5) * 2^-16: field in the top of the mantissa of [1,2), minus (1 - 2^-17)
__device__ __forceinline__ float u01_lo16(uint32_t w) { return __uint_as_float(((w << 7) & 0x007fff80u) | 0x3f800000u) - 0.99999237060546875f; }
__device__ __forceinline__ float u01_hi16(uint32_t w) { return __uint_as_float(((w >> 9) & 0x007fff80u) | 0x3f800000u) - 0.99999237060546875f; }

// 16-bit field in the top of the mantissa of [1,2): m = 1 + k 2^-16
__device__ __forceinline__ float m12_lo16(uint32_t w) { return __uint_as_float(((w << 7) & 0x007fff80u) | 0x3f800000u); }
__device__ __forceinline__ float m12_hi16(uint32_t w) { return __uint_as_float(((w >> 9) & 0x007fff80u) | 0x3f800000u); }

// 4 x noise_ratio * N(0,1) from two 32-bit words, like oracle normal4(): each word is one Box-Muller pair
// (low half -> radius, high half -> angle).  u = m - (1 - 2^-17) in (0,1);  scale folded into the radius:
// noise_ratio sqrt(-2 ln u) = sqrt(noise_k lg2 u);  angle 2 pi u - pi = 2 pi m - (2 pi (1 - 2^-17) + pi), and
// sin(t - pi) = -sin t, cos(t - pi) = -cos t keeps the MUFU argument in [-pi, pi].
__device__ __forceinline__ void normal4_scaled(uint32_t w0, uint32_t w1, float noise_k, float n[4]) {
  const float ra = -fsqrt(noise_k * __log2f(m12_lo16(w0) - 0.99999237060546875f));
  const float rb = -fsqrt(noise_k * __log2f(m12_lo16(w1) - 0.99999237060546875f));
  const float ta = fmaf(6.28318530718f, m12_hi16(w0), -9.42473002f);
  const float tb = fmaf(6.28318530718f, m12_hi16(w1), -9.42473002f);
  n[0] = ra * __cosf(ta); n[1] = ra * __sinf(ta); n[2] = rb * __cosf(tb); n[3] = rb * __sinf(tb);
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ---------------------------------------------------------------------------
// PyFlyt QuadX.update_control, mode 0: rate PID -> motor mix -> saturation
// ---------------------------------------------------------------------------
// pwm[] returns lag_alpha * pwm (see the end of the function)
__device__ __forceinline__ void control_update(Env& e, const DevConfig& c, const float sp[4], float pwm[4]) {
  float cmd[4];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float err = sp[a] - e.swb[a];
    e.pi[a] = clampf(fmaf(c.kiT[a], err, e.pi[a]), -c.lim[a], c.lim[a]);
    const float d = c.kd_T[a] * (err - e.pe[a]);
    cmd[a] = clampf(fmaf(c.kp[a], err, e.pi[a]) + d, -c.lim[a], c.lim[a]);
    e.pe[a] = err;
  }
  cmd[3] = sp[3];
#pragma unroll
  for (int m = 0; m < 4; ++m)
    pwm[m] = c.map[4 * m + 0] * cmd[0] + c.map[4 * m + 1] * cmd[1] + c.map[4 * m + 2] * cmd[2] + c.map[4 * m + 3] * cmd[3];
  const float high = fmaxf(fmaxf(pwm[0], pwm[1]), fmaxf(pwm[2], pwm[3]));
  if (high > 1.0f) {
    const float inv = frcp(high);
#pragma unroll
    for (int m = 0; m < 4; ++m) pwm[m] *= inv;
  }
  const float low = fminf(fminf(pwm[0], pwm[1]), fminf(pwm[2], pwm[3]));
  if (low < c.pwm_idle) {
    const float k = (c.pwm_idle - low) * frcp(1.0f - low);
#pragma unroll
    for (int m = 0; m < 4; ++m) pwm[m] = fmaf(1.0f - pwm[m], k, pwm[m]);
  }
  // the motor lag thr += alpha (pwm - thr) is applied as thr (1 - alpha) + (alpha pwm): hoist alpha pwm out of the sub-steps
#pragma unroll
  for (int m = 0; m < 4; ++m) pwm[m] *= c.lag_alpha;
}

// ---------------------------------------------------------------------------
// PyFlyt PID.step for one channel; g = kp[w] ki*T[w] kd/T[w] lim[w], channel k
// ---------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ float pid_channel(const float* g, int k, float& mi, float& me, float state, float sp) {
  const float err = sp - state, lim = g[3 * W + k];
  mi = clampf(fmaf(g[W + k], err, mi), -lim, lim);
  const float d = g[2 * W + k] * (err - me);
  me = err;
  return clampf(fmaf(g[k], err, mi) + d, -lim, lim);
}

// QuadX.update_control for flight modes != 0: the outer loops turn the setpoint into (rate commands, thrust) and the
// mode-0 controller above finishes the job.  seul = Euler row of the Aviary.state snapshot.  Mode -1 bypasses
// everything (setpoint = motor pwm, no mixing, no saturation).
__device__ __forceinline__ void control_update_cascade(Env& e, const DevConfig& c, const float sp[4], const float seul[3], float pwm[4]) {
  const int mode = c.flight_mode;
  if (mode == -1) {
#pragma unroll
    for (int m = 0; m < 4; ++m) pwm[m] = sp[m] * c.lag_alpha;
    return;
  }
  float a[4] = {sp[0], sp[1], sp[2], sp[3]};
  if (mode == 1 || mode == 3) {  // angles -> rates
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] = pid_channel<3>(c.att, k, e.cp[k], e.cp[3 + k], seul[k], a[k]);
  } else if (mode >= 4) {
    if (mode == 7) {  // ground-frame position -> ground-frame velocity
      a[0] = pid_channel<2>(c.lpos, 0, e.cp[10], e.cp[12], e.spx, a[0]);
      a[1] = pid_channel<2>(c.lpos, 1, e.cp[11], e.cp[13], e.spy, a[1]);
    }
    if (mode >= 6) {  // ground frame -> yaw-rotated local frame
      float sy, cy;
      sincosf(seul[2], &sy, &cy);
      const float gx = a[0], gy = a[1];
      a[0] = cy * gx + sy * gy;
      a[1] = cy * gy - sy * gx;
    }
    // local velocity -> tilt angles: +x needs +pitch, +y needs -roll
    const float ang0 = pid_channel<2>(c.vel, 0, e.cp[6], e.cp[8], e.svb[0], a[0]);
    const float ang1 = pid_channel<2>(c.vel, 1, e.cp[7], e.cp[9], e.svb[1], a[1]);
    a[0] = -ang1; a[1] = ang0;
    a[0] = pid_channel<3>(c.att, 0, e.cp[0], e.cp[3], seul[0], a[0]);
    a[1] = pid_channel<3>(c.att, 1, e.cp[1], e.cp[4], seul[1], a[1]);
    if (mode == 7) a[2] = pid_channel<3>(c.att, 2, e.cp[2], e.cp[5], seul[2], a[2]);  // yaw is an angle only in mode 7
  }
  if (mode == 2 || mode == 3 || mode == 4 || mode == 7) a[3] = pid_channel<1>(c.zpos, 0, e.cp[16], e.cp[17], e.spz, a[3]);
  if (mode != 0) a[3] = pid_channel<1>(c.zvel, 0, e.cp[14], e.cp[15], e.svb[2], a[3]);
  a[3] = __saturatef(a[3]);
  control_update(e, c, a, pwm);
}

// ---------------------------------------------------------------------------
// one 1/physics_hz sub-step: Motors.physics_update + drag + update_state +
// pybullet.stepSimulation (free flight) + the declared floor stand-in.
//
// Bullet integrates the WORLD angular velocity, w' = w + h R a_b, and then
// q' = exp(h w'/2) q.  With w = R w_b that is exactly  w_b' = w_b + h a_b  and
// q' = q exp(h w_b'/2)  (a rotation about u leaves u unchanged), so the
// angular half is carried in the body frame and needs no rotation matrix.
// (Bullet's +-100 clamp acts on world components; here on body components --
// the two differ only beyond 100 rad/s.)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void physics_substep(Env& e, const DevConfig& c, const float apwm[4], const float nz[4], const bool last) {
  // motors: first-order lag, multiplicative noise, thrust and torques
  float fz = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float t = fmaf(e.thr[m], c.one_m_alpha, apwm[m]);
    t = fmaf(nz[m], t, t);                // nz already scaled by noise_ratio
    e.thr[m] = t;
    const float rr = fabsf(t) * t;        // rpm |rpm| / max_rpm^2
    fz += rr;
    tx = fmaf(c.my[m], rr, tx);           // r x F = (y F, -x F, 0)
    ty = fmaf(c.mx[m], rr, ty);
    tz = fmaf(c.torque_k[m], rr, tz);
  }
  fz *= c.thrust_k; tx *= c.thrust_k; ty *= -c.thrust_k;
  // drag from the (stale) snapshot, body frame
  const float fbx = c.ndrag_c * fabsf(e.svb[0]) * e.svb[0];
  const float fby = c.ndrag_c * fabsf(e.svb[1]) * e.svb[1];
  const float fbz = fmaf(c.ndrag_c * fabsf(e.svb[2]), e.svb[2], fz);
  if (!(e.flags & F_CONTACT)) {
    tx = fmaf(c.ndrag_pqr * fabsf(e.swb[0]), e.swb[0], tx);
    ty = fmaf(c.ndrag_pqr * fabsf(e.swb[1]), e.swb[1], ty);
    tz = fmaf(c.ndrag_pqr * fabsf(e.swb[2]), e.swb[2], tz);
  }
  // rotation matrix of the current attitude
  const float x = e.qx, y = e.qy, z = e.qz, w = e.qw;
  const float x2 = x + x, y2 = y + y, z2 = z + z;
  // 18 instructions: each diagonal entry is two dependent FFMA, each off-diagonal pair shares one product
  const float wx = w * x2, wy = w * y2, wz = w * z2;
  const float r00 = fmaf(-y2, y, fmaf(-z2, z, 1.f)), r11 = fmaf(-x2, x, fmaf(-z2, z, 1.f)), r22 = fmaf(-x2, x, fmaf(-y2, y, 1.f));
  const float r01 = fmaf(x, y2, -wz), r10 = fmaf(x, y2, wz);
  const float r02 = fmaf(x, z2, wy), r20 = fmaf(x, z2, -wy);
  const float r12 = fmaf(y, z2, -wx), r21 = fmaf(y, z2, wx);
  if (c.state_stale) {  // QuadX.update_state runs before stepSimulation
    e.swb[0] = e.wx; e.swb[1] = e.wy; e.swb[2] = e.wz;
    e.svb[0] = r00 * e.vx + r10 * e.vy + r20 * e.vz;
    e.svb[1] = r01 * e.vx + r11 * e.vy + r21 * e.vz;
    e.svb[2] = r02 * e.vx + r12 * e.vy + r22 * e.vz;
    if (last) { e.sqx = x; e.sqy = y; e.sqz = z; e.sqw = w; e.spx = e.px; e.spy = e.py; e.spz = e.pz; }  // only the final pose is read
  }
  // angular half, body frame
  // w x (I w) for a diagonal inertia is ((I2 - I1) wy wz, (I0 - I2) wz wx, (I1 - I0) wx wy); gk = -h / I_k times those
  // differences (0 when the gyroscopic term is off), so each axis is one product and two FFMA
  const float wyz = e.wy * e.wz, wzx = e.wz * e.wx, wxy = e.wx * e.wy;
  e.wx = fmaf(c.gk[0], wyz, fmaf(c.hI[0], tx, e.wx));
  e.wy = fmaf(c.gk[1], wzx, fmaf(c.hI[1], ty, e.wy));
  e.wz = fmaf(c.gk[2], wxy, fmaf(c.hI[2], tz, e.wz));
  // linear half, world frame: semi-implicit Euler
  const float hm = c.hm;
  e.vx = fmaf(hm, r00 * fbx + r01 * fby + r02 * fbz, e.vx);
  e.vy = fmaf(hm, r10 * fbx + r11 * fby + r12 * fbz, e.vy);
  e.vz = fmaf(hm, r20 * fbx + r21 * fby + r22 * fbz, e.vz) - c.hg;
  // btMultiBody's +-max_coord_vel clamp: one max over the six components, the clamp itself is a cold path
  if (fmaxf(fmaxf(fmaxf(fabsf(e.vx), fabsf(e.vy)), fmaxf(fabsf(e.vz), fabsf(e.wx))), fmaxf(fabsf(e.wy), fabsf(e.wz))) > c.vmax) {
    e.vx = clampf(e.vx, -c.vmax, c.vmax); e.vy = clampf(e.vy, -c.vmax, c.vmax); e.vz = clampf(e.vz, -c.vmax, c.vmax);
    e.wx = clampf(e.wx, -c.vmax, c.vmax); e.wy = clampf(e.wy, -c.vmax, c.vmax); e.wz = clampf(e.wz, -c.vmax, c.vmax);
  }
  e.px = fmaf(c.h, e.vx, e.px);
  e.py = fmaf(c.h, e.vy, e.py);
  e.pz = fmaf(c.h, e.vz, e.pz);
  // q <- q exp(h w_b / 2): series in s = (|w| h / 2)^2 (no sqrt / sin / cos), then renormalise
  // sin(x)/x * h/2 and cos(x) in s = x^2, x = |w| h / 2 <= 0.37 at the +-100 rad/s clamp: truncation < 4e-9
  // The product is formed as q + q (dq - 1): the increment is ~|w| h / 2 relative to q, so its rounding is negligible and
  // each component takes one rounding at |q| scale instead of one per FFMA of the full product -- the attitude random
  // walk over thousands of sub-steps is what drives the open-loop position drift against the float64 oracle.
  const float s = (e.wx * e.wx + e.wy * e.wy + e.wz * e.wz) * c.hh2;
  const float kq = fmaf(s, fmaf(s, c.kq2, c.kq1), c.hh);
  const float dwm1 = s * fmaf(s, fmaf(s, -1.f / 720.f, 1.f / 24.f), -0.5f);  // cos(|w| h / 2) - 1
  const float dx = e.wx * kq, dy = e.wy * kq, dz = e.wz * kq;
  const float nx = x + (w * dx + x * dwm1 + y * dz - z * dy);
  const float ny = y + (w * dy - x * dz + y * dwm1 + z * dx);
  const float nzq = z + (w * dz + x * dy - y * dx + z * dwm1);
  const float nw = w + (w * dwm1 - x * dx - y * dy - z * dz);
  const float inv = frsqrt(nx * nx + ny * ny + nzq * nzq + nw * nw);
  e.qx = nx * inv; e.qy = ny * inv; e.qz = nzq * inv; e.qw = nw * inv;
  // floor stand-in: sticky plane at floor_z (zeroes world v_xy and world w_xy)
  if (e.pz < c.floor_z) {
    e.pz = c.floor_z; e.vz = fmaxf(e.vz, 0.f); e.vx = 0.f; e.vy = 0.f;
    const float X = e.qx, Y = e.qy, Z = e.qz, W = e.qw;
    const float a20 = 2.f * (X * Z - W * Y), a21 = 2.f * (Y * Z + W * X), a22 = 1.f - 2.f * (X * X + Y * Y);
    const float wzw = a20 * e.wx + a21 * e.wy + a22 * e.wz;  // world yaw rate survives
    e.wx = a20 * wzw; e.wy = a21 * wzw; e.wz = a22 * wzw;
    e.flags |= F_CONTACT;
  } else {
    e.flags &= ~F_CONTACT;
  }
  if (!c.state_stale) {
    const float X = e.qx, Y = e.qy, Z = e.qz, W = e.qw;
    const float a00 = 1.f - 2.f * (Y * Y + Z * Z), a01 = 2.f * (X * Y - W * Z), a02 = 2.f * (X * Z + W * Y);
    const float a10 = 2.f * (X * Y + W * Z), a11 = 1.f - 2.f * (X * X + Z * Z), a12 = 2.f * (Y * Z - W * X);
    const float a20 = 2.f * (X * Z - W * Y), a21 = 2.f * (Y * Z + W * X), a22 = 1.f - 2.f * (X * X + Y * Y);
    e.swb[0] = e.wx; e.swb[1] = e.wy; e.swb[2] = e.wz;
    e.svb[0] = a00 * e.vx + a10 * e.vy + a20 * e.vz;
    e.svb[1] = a01 * e.vx + a11 * e.vy + a21 * e.vz;
    e.svb[2] = a02 * e.vx + a12 * e.vy + a22 * e.vz;
    if (last) { e.sqx = X; e.sqy = Y; e.sqz = Z; e.sqw = W; e.spx = e.px; e.spy = e.py; e.spz = e.pz; }
  }
}

// pybullet.getEulerFromQuaternion (ZYX, with its gimbal guard)
__device__ __forceinline__ void quat_to_euler(float x, float y, float z, float w, float& roll, float& pitch, float& yaw) {
  const float sarg = -2.f * (x * z - w * y);
  if (sarg <= -0.99999f) {
    roll = 0.f; pitch = -1.57079632679f; yaw = 2.f * atan2f(x, -y);
  } else if (sarg >= 0.99999f) {
    roll = 0.f; pitch = 1.57079632679f; yaw = 2.f * atan2f(-x, y);
  } else {
    roll = atan2f(2.f * (y * z + w * x), w * w - x * x - y * y + z * z);
    pitch = asinf(sarg);
    yaw = atan2f(2.f * (x * y + w * z), w * w + x * x - y * y - z * z);
  }
}

// pybullet.getQuaternionFromEuler, hover.py:233
__device__ __forceinline__ void euler_to_quat(float roll, float pitch, float yaw, float& x, float& y, float& z, float& w) {
  float sr, cr, sp, cp, sy, cy;
  __sincosf(0.5f * roll, &sr, &cr);
  __sincosf(0.5f * pitch, &sp, &cp);
  __sincosf(0.5f * yaw, &sy, &cy);
  x = sr * cp * cy - cr * sp * sy;
  y = cr * sp * cy + sr * cp * sy;
  z = cr * cp * sy - sr * sp * cy;
  w = cr * cp * cy + sr * sp * sy;
}

// ---------------------------------------------------------------------------
// Camera: analytic stand-in for Camera.capture_image + detect_rectangle
// (hover.py:157-222, 241-248).  The four corners of the red face are projected
// through the PyFlyt FPV camera (body Euler angles with the pitch offset by the
// tilt; FOV 90, 128x128) and the features are taken on the pixel lattice the
// way cv2.findContours / contourArea / boundingRect would report them.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void vision(const Env& e, const DevConfig& c, bool& vis, float& cx, float& cy, float& area,
                                       float& ratio) {
  const float x = e.qx, y = e.qy, z = e.qz, w = e.qw;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  const float r00 = 1.f - 2.f * (yy + zz), r01 = 2.f * (xy - wz), r02 = 2.f * (xz + wy);
  const float r10 = 2.f * (xy + wz), r11 = 1.f - 2.f * (xx + zz), r12 = 2.f * (yz - wx);
  const float r20 = 2.f * (xz - wy), r21 = 2.f * (yz + wx), r22 = 1.f - 2.f * (xx + yy);
  // sin / cos of the roll Euler angle = (r21, r22) / cos(pitch)
  float sph = 0.f, cph = 1.f;
  if (fabsf(r20) < 0.99999f) {
    const float inv = frsqrt(r21 * r21 + r22 * r22);
    sph = r21 * inv; cph = r22 * inv;
  }
  // camera axes in the body frame: Rx(-roll) Ry(-tilt) Rx(roll) applied to x, z
  const float sd = c.cam_sd, cd = c.cam_cd;
  const float fx = cd, fy = sd * sph, fzz = sd * cph;
  const float ux = -sd * cph, uy = sph * cph * (cd - 1.f), uz = fmaf(sph, sph, cd * cph * cph);
  // right = fwd x up
  const float rx = fy * uz - fzz * uy, ry = fzz * ux - fx * uz, rz = fx * uy - fy * ux;
  float pxs[4], pys[4];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float dx = c.panel[3 * k + 0] - e.px, dy = c.panel[3 * k + 1] - e.py, dz = c.panel[3 * k + 2] - e.pz;
    const float bx = r00 * dx + r10 * dy + r20 * dz - c.cam_off[0];
    const float by = r01 * dx + r11 * dy + r21 * dz - c.cam_off[1];
    const float bz = r02 * dx + r12 * dy + r22 * dz - c.cam_off[2];
    const float depth = fx * bx + fy * by + fzz * bz;
    ok = ok && (depth > c.cam_near);
    const float k1 = c.inv_tan * frcp(fmaxf(depth, 1e-9f));
    pxs[k] = fmaf((rx * bx + ry * by + rz * bz) * k1, c.half_res, c.half_res);
    pys[k] = fmaf(-(ux * bx + uy * by + uz * bz) * k1, c.half_res, c.half_res);
  }
  const float xmin = fminf(fminf(pxs[0], pxs[1]), fminf(pxs[2], pxs[3]));
  const float xmax = fmaxf(fmaxf(pxs[0], pxs[1]), fmaxf(pxs[2], pxs[3]));
  const float ymin = fminf(fminf(pys[0], pys[1]), fminf(pys[2], pys[3]));
  const float ymax = fmaxf(fmaxf(pys[0], pys[1]), fmaxf(pys[2], pys[3]));
  ok = ok && xmin >= c.margin && ymin >= c.margin && xmax <= c.res - c.margin && ymax <= c.res - c.margin;
  float a2 = 0.f, per = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int kn = (k + 1) & 3;
    a2 += pxs[k] * pys[kn] - pxs[kn] * pys[k];
    const float ex = pxs[kn] - pxs[k], ey = pys[kn] - pys[k];
    per += fsqrt(ex * ex + ey * ey);
  }
  const float wpx = floorf(xmax - 0.5f) - ceilf(xmin - 0.5f) + 1.f;
  const float hpx = floorf(ymax - 0.5f) - ceilf(ymin - 0.5f) + 1.f;
  ok = ok && wpx >= 2.f && hpx >= 2.f;
  vis = ok;
  cx = ok ? fmaf(0.25f * (pxs[0] + pxs[1] + pxs[2] + pxs[3]) - 0.5f, c.inv_half_res, -1.f) : 0.f;
  cy = ok ? fmaf(0.25f * (pys[0] + pys[1] + pys[2] + pys[3]) - 0.5f, c.inv_half_res, -1.f) : 0.f;
  area = ok ? fmaxf(0.5f * fabsf(a2) - 0.5f * per + 1.f, 0.f) * c.inv_res2 : 0.f;
  ratio = ok ? wpx * frcp(hpx) : 0.f;
}

// yaw task: centre of the red sphere (main.py:16-23, radius 0.1 at (2,0,1); yaw.py:63 detect_red_sphere_center is
// missing from the reference) projected through the same FPV camera model; (0,0) when it is not in the image.
__device__ __forceinline__ void vision_point(const Env& e, const DevConfig& c, bool& vis, float& cx, float& cy) {
  const float x = e.qx, y = e.qy, z = e.qz, w = e.qw;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  const float r00 = 1.f - 2.f * (yy + zz), r01 = 2.f * (xy - wz), r02 = 2.f * (xz + wy);
  const float r10 = 2.f * (xy + wz), r11 = 1.f - 2.f * (xx + zz), r12 = 2.f * (yz - wx);
  const float r20 = 2.f * (xz - wy), r21 = 2.f * (yz + wx), r22 = 1.f - 2.f * (xx + yy);
  float sph = 0.f, cph = 1.f;
  if (fabsf(r20) < 0.99999f) {
    const float inv = frsqrt(r21 * r21 + r22 * r22);
    sph = r21 * inv; cph = r22 * inv;
  }
  const float sd = c.cam_sd, cd = c.cam_cd;
  const float fx = cd, fy = sd * sph, fzz = sd * cph;
  const float ux = -sd * cph, uy = sph * cph * (cd - 1.f), uz = fmaf(sph, sph, cd * cph * cph);
  const float rx = fy * uz - fzz * uy, ry = fzz * ux - fx * uz, rz = fx * uy - fy * ux;
  const float dx = c.panel[0] - e.px, dy = c.panel[1] - e.py, dz = c.panel[2] - e.pz;
  const float bx = r00 * dx + r10 * dy + r20 * dz - c.cam_off[0];
  const float by = r01 * dx + r11 * dy + r21 * dz - c.cam_off[1];
  const float bz = r02 * dx + r12 * dy + r22 * dz - c.cam_off[2];
  const float depth = fx * bx + fy * by + fzz * bz;
  const float k1 = c.inv_tan * frcp(fmaxf(depth, 1e-9f));
  const float px = fmaf((rx * bx + ry * by + rz * bz) * k1, c.half_res, c.half_res);
  const float py = fmaf(-(ux * bx + uy * by + uz * bz) * k1, c.half_res, c.half_res);
  vis = depth > c.cam_near && px >= 0.f && px <= c.res && py >= 0.f && py <= c.res;
  cx = vis ? fmaf(px, c.inv_half_res, -1.f) : 0.f;
  cy = vis ? fmaf(py, c.inv_half_res, -1.f) : 0.f;
}

// Aviary(start_pos, start_orn) + Aviary.reset() + the env bookkeeping of
// hover.py:98-107.  prev_action is deliberately left alone (hover.py:31,357).
__device__ __forceinline__ void respawn(Env& e, const DevConfig& c, uint32_t k0, uint32_t k1) {
  float px = c.start_pos[0], py = c.start_pos[1], pz = c.start_pos[2], yaw = c.start_rpy[2];
  if (c.spawn_pos_noise != 0.f || c.spawn_yaw_noise != 0.f) {
    const uint4 b = philox4x32_10(make_uint4(0u, STREAM_SPAWN, e.rng_ctr, 0u), k0, k1);
    px = fmaf(c.spawn_pos_noise, 2.f * u01(b.x) - 1.f, px);
    py = fmaf(c.spawn_pos_noise, 2.f * u01(b.y) - 1.f, py);
    pz = fmaf(c.spawn_pos_noise, 2.f * u01(b.z) - 1.f, pz);
    yaw = fmaf(c.spawn_yaw_noise, 2.f * u01(b.w) - 1.f, yaw);
  }
  pz = fmaxf(pz, c.floor_z);
  e.px = px; e.py = py; e.pz = pz;
  float sr, cr, sp, cp, sy, cy;
  sincosf(0.5f * c.start_rpy[0], &sr, &cr);
  sincosf(0.5f * c.start_rpy[1], &sp, &cp);
  sincosf(0.5f * yaw, &sy, &cy);
  e.qx = sr * cp * cy - cr * sp * sy;
  e.qy = cr * sp * cy + sr * cp * sy;
  e.qz = cr * cp * sy - sr * sp * cy;
  e.qw = cr * cp * cy + sr * sp * sy;
  e.vx = e.vy = e.vz = 0.f;
  e.wx = e.wy = e.wz = 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m) e.thr[m] = c.spawn_thr;
#pragma unroll
  for (int a = 0; a < 3; ++a) { e.pi[a] = 0.f; e.pe[a] = 0.f; e.swb[a] = 0.f; e.svb[a] = 0.f; }
  e.sqx = e.qx; e.sqy = e.qy; e.sqz = e.qz; e.sqw = e.qw; e.spx = px; e.spy = py; e.spz = pz;
  e.flags = (pz <= c.floor_z) ? F_CONTACT : 0u;
  e.step_count = 0;
  e.pcx = e.pcy = e.parea = e.pratio = 0.f;
  e.ep_ret = 0.f;
  if (c.task == 1) { e.pa[0] = e.pa[1] = e.pa[2] = e.pa[3] = 0.f; }  // yaw.py:90-92 clears the action history
}

}  // namespace qx
