/*
 * quadx_oracle.c -- plain C (float64) restatement of the QuadX hover env step,
 * threaded with pthreads (one contiguous env range per host thread).  TEST INFRASTRUCTURE ONLY: it is the CPU baseline timed
 * by bench.py (cpu_baseline / --impl reference) and a fast checker for large
 * batches; the product path never links or calls it.
 *
 * It restates, function by function, the same things as the numpy oracle
 * (oracle/quadx_model.py, oracle/vision.py, oracle/hover_oracle.py), i.e.
 *   - /root/reference/simulation/hover.py:72-113 (reset), :224-272 (obs),
 *     :274-332 (reward / termination), :334-358 (step)
 *   - PyFlyt 0.21.0 QuadX (mode 0 as hover.py:92 sets it, plus the outer loops of
 *     flight modes -1..7, cf2x.yaml:21-53) + pybullet 3.2.7 free-flight integration as
 *     restated in SURVEY.md section 9 (third-party; PARITY UNPINNED), with the
 *     parameters of cf2x.yaml:1-19 and cf2x.urdf:10-68.
 * tests/test_c_oracle.py pins it against the numpy oracle (which is pinned
 * against the reference's own hover.py) to 1e-9.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
  /* drone (oracle/quadx_model.py QuadXParams) */
  double total_thrust, thrust_coef, torque_coef, noise_ratio, tau;
  double drag_coef_xyz, drag_area_xyz, drag_coef_pqr, air_density;
  double kp[3], ki[3], kd[3], lim[3];
  double mass, inertia[3], motor_x[4], motor_y[4], torque_sign[4], motor_map[16], pwm_idle;
  double physics_hz, control_hz, gravity;
  int32_t state_stale, gyro;
  double max_coord_vel, floor_z;
  double cam_tilt_up_deg, cam_fov_deg, cam_res, cam_near, cam_offset[3], vis_margin_px, panel[12];
  /* env (oracle/hover_oracle.py HoverConfig) */
  int32_t env_step_ratio, max_steps, floor_grace_steps, reset_idle_steps;
  double agent_dt, flight_dome_size, floor_threshold, target_area, target_ratio, action_scale[3];
  double start_pos[3], start_rpy[3], spawn_throttle, spawn_pos_noise, spawn_yaw_noise;
  int32_t render, auto_reset, noise;
  /* PyFlyt QuadX.set_mode (-1..7) and the outer loops' gains (kp[w] ki[w] kd[w] lim[w]), cf2x.yaml:21-53 */
  int32_t flight_mode;
  double thrust_scale, thrust_bias;
  double att[12], vel[8], lpos[8], zpos[4], zvel[4];
} OrcConfig;

typedef struct {
  double pos[3], quat[4], vel[3], omega[3], thr[4], pid_i[3], pid_e[3];
  double cpid[18]; /* outer-loop (integral, previous error): att 0/3, vel 6/8, pos 10/12, z_vel 14/15, z_pos 16/17 */
  double s_wb[3], s_euler[3], s_vb[3], s_pos[3];
  double action[4], prev_action[4], prev_centre[2], prev_area, prev_ratio, prev_euler[3], ep_return;
  int64_t step_count;
  uint64_t rng_ctr;
  int32_t contact, terminated, truncated, oob, on_floor;
} OrcEnv;

typedef struct {
  OrcConfig c;
  int64_t n;
  uint64_t seed, env_id0;
  OrcEnv* e;
  double sum_ret;
  int64_t sum_len, n_done;
} Orc;

enum { STREAM_STEP = 0, STREAM_RESET = 1, STREAM_SPAWN = 2 };

/* ---- Philox4x32-7 (oracle/quadx_model.py philox4x32, ENV_PHILOX_ROUNDS) ---- */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 7; ++r) {
    if (r > 0) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
}
static void env_bits(const Orc* o, int64_t i, uint32_t w0, uint32_t stream, uint64_t ctr, uint32_t out[4]) {
  uint64_t id = o->env_id0 + (uint64_t)i;
  uint32_t k0 = (uint32_t)(o->seed & 0xFFFFFFFFu) ^ (uint32_t)(id & 0xFFFFFFFFu);
  uint32_t k1 = (uint32_t)(o->seed >> 32) ^ (uint32_t)(id >> 32);
  out[0] = w0; out[1] = stream; out[2] = (uint32_t)(ctr & 0xFFFFFFFFu); out[3] = (uint32_t)(ctr >> 32);
  philox(out, k0, k1);
}
static double u01(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }
static double u01_16(uint32_t x) { return ((double)x + 0.5) * (1.0 / 65536.0); }
/* oracle normal4(): one word = one Box-Muller pair (low half radius, high half angle) */
static void normals(const Orc* o, int64_t i, int sub, uint32_t stream, uint64_t ctr, double nz[4]) {
  uint32_t b[4];
  env_bits(o, i, (uint32_t)(sub >> 1), stream, ctr, b);
  for (int k = 0; k < 2; ++k) {
    uint32_t w = b[2 * (sub & 1) + k];
    double r = sqrt(-2.0 * log(u01_16(w & 0xFFFFu))), t = 2.0 * M_PI * u01_16(w >> 16) - M_PI;
    nz[2 * k] = r * cos(t);
    nz[2 * k + 1] = r * sin(t);
  }
}

/* ---- rotations (pybullet conventions) -------------------------------------- */
static void quat_to_mat(const double q[4], double R[9]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void quat_to_euler(const double q[4], double e[3]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double sarg = -2.0 * (x * z - w * y);
  if (sarg <= -0.99999) { e[0] = 0; e[1] = -0.5 * M_PI; e[2] = 2 * atan2(x, -y); }
  else if (sarg >= 0.99999) { e[0] = 0; e[1] = 0.5 * M_PI; e[2] = 2 * atan2(-x, y); }
  else {
    e[0] = atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z);
    e[1] = asin(sarg);
    e[2] = atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
  }
}
static void euler_to_quat(const double e[3], double q[4]) {
  double cr = cos(0.5 * e[0]), sr = sin(0.5 * e[0]), cp = cos(0.5 * e[1]), sp = sin(0.5 * e[1]), cy = cos(0.5 * e[2]), sy = sin(0.5 * e[2]);
  q[0] = sr * cp * cy - cr * sp * sy; q[1] = cr * sp * cy + sr * cp * sy;
  q[2] = cr * cp * sy - sr * sp * cy; q[3] = cr * cp * cy + sr * sp * sy;
}
static void matT_vec(const double R[9], const double v[3], double o[3]) {
  o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2]; o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2]; o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
static void mat_vec(const double R[9], const double v[3], double o[3]) {
  o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2]; o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2]; o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
static double sgn(double x) { return (x > 0) - (x < 0); }

/* ---- QuadX.update_state ------------------------------------------------------ */
static void snapshot(OrcEnv* e) {
  double R[9];
  quat_to_mat(e->quat, R);
  matT_vec(R, e->omega, e->s_wb);
  matT_vec(R, e->vel, e->s_vb);
  quat_to_euler(e->quat, e->s_euler);
  memcpy(e->s_pos, e->pos, sizeof(e->pos));
}

/* ---- PyFlyt PID.step: w channels, gains g = kp[w] ki[w] kd[w] lim[w], memory mi / me --------------------- */
static void pid(const double* g, int w, double T, double* mi, double* me, const double* state, const double* sp, double* out) {
  for (int a = 0; a < w; ++a) {
    double err = sp[a] - state[a], lim = g[3 * w + a];
    mi[a] = clampd(mi[a] + g[w + a] * err * T, -lim, lim);
    double d = g[2 * w + a] * (err - me[a]) / T;
    me[a] = err;
    out[a] = clampd(g[a] * err + mi[a] + d, -lim, lim);
  }
}

/* ---- QuadX.update_control: outer loops of the flight mode -> rate PID -> mix -> saturation ------------------- */
static void control(const OrcConfig* c, OrcEnv* e, const double sp[4], double pwm[4]) {
  double T = 1.0 / c->control_hz, cmd[4], a[3] = {sp[0], sp[1], sp[2]}, z = sp[3];
  int mode = c->flight_mode;
  double* m = e->cpid;
  if (mode == -1) { for (int k = 0; k < 4; ++k) pwm[k] = sp[k]; return; }
  if (mode == 1 || mode == 3) pid(c->att, 3, T, m + 0, m + 3, e->s_euler, a, a);
  else if (mode >= 4) {
    if (mode == 7) pid(c->lpos, 2, T, m + 10, m + 12, e->s_pos, a, a);
    if (mode >= 6) {
      double cy = cos(e->s_euler[2]), sy = sin(e->s_euler[2]), gx = a[0], gy = a[1];
      a[0] = cy * gx + sy * gy; a[1] = -sy * gx + cy * gy;
    }
    double ang[2];
    pid(c->vel, 2, T, m + 6, m + 8, e->s_vb, a, ang);
    a[0] = -ang[1]; a[1] = ang[0];
    if (mode == 7) pid(c->att, 3, T, m + 0, m + 3, e->s_euler, a, a);
    else {
      double g2[8] = {c->att[0], c->att[1], c->att[3], c->att[4], c->att[6], c->att[7], c->att[9], c->att[10]};
      pid(g2, 2, T, m + 0, m + 3, e->s_euler, a, a);
    }
  }
  if (mode == 2 || mode == 3 || mode == 4 || mode == 7) pid(c->zpos, 1, T, m + 16, m + 17, e->s_pos + 2, &z, &z);
  if (mode != 0) pid(c->zvel, 1, T, m + 14, m + 15, e->s_vb + 2, &z, &z);
  z = clampd(z, 0.0, 1.0);
  for (int k = 0; k < 3; ++k) {
    double err = a[k] - e->s_wb[k];
    e->pid_i[k] = clampd(e->pid_i[k] + c->ki[k] * err * T, -c->lim[k], c->lim[k]);
    double d = c->kd[k] * (err - e->pid_e[k]) / T;
    cmd[k] = clampd(c->kp[k] * err + e->pid_i[k] + d, -c->lim[k], c->lim[k]);
    e->pid_e[k] = err;
  }
  cmd[3] = z;
  double high = -1e300, low = 1e300;
  for (int k = 0; k < 4; ++k) {
    pwm[k] = 0;
    for (int j = 0; j < 4; ++j) pwm[k] += cmd[j] * c->motor_map[4 * k + j];
    if (pwm[k] > high) high = pwm[k];
  }
  if (high > 1.0) for (int k = 0; k < 4; ++k) pwm[k] /= high;
  for (int k = 0; k < 4; ++k) if (pwm[k] < low) low = pwm[k];
  if (low < c->pwm_idle) for (int k = 0; k < 4; ++k) pwm[k] = pwm[k] + (1.0 - pwm[k]) / (1.0 - low) * (c->pwm_idle - low);
}

/* ---- update_physics + update_state + stepSimulation, one sub-step ------------ */
static void substep(const OrcConfig* c, OrcEnv* e, const double pwm[4], const double* nz) {
  double h = 1.0 / c->physics_hz, max_rpm = sqrt(c->total_thrust / (4.0 * c->thrust_coef));
  double fsum = 0, tb[3] = {0, 0, 0};
  for (int m = 0; m < 4; ++m) {
    e->thr[m] = e->thr[m] + (h / c->tau) * (pwm[m] - e->thr[m]);
    if (nz) e->thr[m] = e->thr[m] + nz[m] * e->thr[m] * c->noise_ratio;
    double rpm = e->thr[m] * max_rpm, rr = fabs(rpm) * rpm, thrust = c->thrust_coef * rr;
    fsum += thrust;
    tb[0] += c->motor_y[m] * thrust;
    tb[1] -= c->motor_x[m] * thrust;
    tb[2] += c->torque_coef * rr * c->torque_sign[m];
  }
  double dc = 0.5 * c->air_density * c->drag_coef_xyz * c->drag_area_xyz, fb[3];
  for (int a = 0; a < 3; ++a) fb[a] = -sgn(e->s_vb[a]) * dc * e->s_vb[a] * e->s_vb[a];
  fb[2] += fsum;
  if (!e->contact) for (int a = 0; a < 3; ++a) tb[a] += -c->drag_coef_pqr * e->s_wb[a] * e->s_wb[a] * sgn(e->s_wb[a]);
  if (c->state_stale) snapshot(e);
  double R[9], wb[3], g[3] = {0, 0, 0}, wd[3], wdw[3], acc[3];
  quat_to_mat(e->quat, R);
  matT_vec(R, e->omega, wb);
  if (c->gyro) {
    double L[3] = {c->inertia[0] * wb[0], c->inertia[1] * wb[1], c->inertia[2] * wb[2]};
    g[0] = wb[1] * L[2] - wb[2] * L[1]; g[1] = wb[2] * L[0] - wb[0] * L[2]; g[2] = wb[0] * L[1] - wb[1] * L[0];
  }
  for (int a = 0; a < 3; ++a) wd[a] = (tb[a] - g[a]) / c->inertia[a];
  mat_vec(R, wd, wdw);
  mat_vec(R, fb, acc);
  for (int a = 0; a < 3; ++a) acc[a] /= c->mass;
  acc[2] -= c->gravity;
  for (int a = 0; a < 3; ++a) {
    e->omega[a] = clampd(e->omega[a] + h * wdw[a], -c->max_coord_vel, c->max_coord_vel);
    e->vel[a] = clampd(e->vel[a] + h * acc[a], -c->max_coord_vel, c->max_coord_vel);
    e->pos[a] = e->pos[a] + h * e->vel[a];
  }
  double ang = sqrt(e->omega[0] * e->omega[0] + e->omega[1] * e->omega[1] + e->omega[2] * e->omega[2]);
  double half = 0.5 * ang * h, k = ang < 1e-3 ? 0.5 * h - h * h * h * 0.020833333333 * ang * ang : sin(half) / ang;
  double d[4] = {e->omega[0] * k, e->omega[1] * k, e->omega[2] * k, cos(half)}, q[4], *b = e->quat;
  q[0] = d[3] * b[0] + d[0] * b[3] + d[1] * b[2] - d[2] * b[1];
  q[1] = d[3] * b[1] - d[0] * b[2] + d[1] * b[3] + d[2] * b[0];
  q[2] = d[3] * b[2] + d[0] * b[1] - d[1] * b[0] + d[2] * b[3];
  q[3] = d[3] * b[3] - d[0] * b[0] - d[1] * b[1] - d[2] * b[2];
  double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int a = 0; a < 4; ++a) e->quat[a] = q[a] / nq;
  int below = e->pos[2] < c->floor_z;
  if (below) {
    e->pos[2] = c->floor_z;
    if (e->vel[2] < 0) e->vel[2] = 0;
    e->vel[0] = e->vel[1] = 0; e->omega[0] = e->omega[1] = 0;
  }
  e->contact = below;
  if (!c->state_stale) snapshot(e);
}

static void aviary_steps(const Orc* o, int64_t i, OrcEnv* e, const double sp[4], int n_aviary, uint32_t stream, uint64_t ctr) {
  const OrcConfig* c = &o->c;
  int per = (int)(c->physics_hz / c->control_hz), sub = 0;
  double pwm[4] = {0, 0, 0, 0}, nz[4];
  int noisy = c->noise && c->noise_ratio != 0.0;
  for (int s = 0; s < n_aviary; ++s)
    for (int j = 0; j < per; ++j, ++sub) {
      if (j == 0) control(c, e, sp, pwm);
      if (noisy) normals(o, i, sub, stream, ctr, nz);
      substep(c, e, pwm, noisy ? nz : NULL);
    }
}

/* ---- analytic camera (oracle/vision.py analytic_features) ------------------- */
static void vision(const OrcConfig* c, const OrcEnv* e, int* vis, double cen[2], double* area, double* ratio) {
  double R[9], eu[3], off[3], eye[3];
  quat_to_mat(e->quat, R);
  mat_vec(R, c->cam_offset, off);
  for (int a = 0; a < 3; ++a) eye[a] = e->pos[a] + off[a];
  quat_to_euler(e->quat, eu);
  double roll = eu[0], pitch = eu[1] - c->cam_tilt_up_deg * M_PI / 180.0, yaw = eu[2];
  double cr = cos(roll), sr = sin(roll), cp = cos(pitch), sp = sin(pitch), cy = cos(yaw), sy = sin(yaw);
  double fwd[3] = {cy * cp, sy * cp, -sp};
  double right[3] = {-(cy * sp * sr - sy * cr), -(sy * sp * sr + cy * cr), -(cp * sr)};
  double up[3] = {cy * sp * cr + sy * sr, sy * sp * cr - cy * sr, cp * cr};
  double t = tan(c->cam_fov_deg * M_PI / 180.0 / 2.0), halfr = c->cam_res / 2.0, px[4], py[4];
  int ok = 1;
  for (int k = 0; k < 4; ++k) {
    double d[3] = {c->panel[3 * k] - eye[0], c->panel[3 * k + 1] - eye[1], c->panel[3 * k + 2] - eye[2]};
    double depth = d[0] * fwd[0] + d[1] * fwd[1] + d[2] * fwd[2];
    double xr = d[0] * right[0] + d[1] * right[1] + d[2] * right[2], yu = d[0] * up[0] + d[1] * up[1] + d[2] * up[2];
    if (!(depth > c->cam_near)) ok = 0;
    double safe = depth > 1e-9 ? depth : 1e-9;
    px[k] = (xr / (safe * t) + 1.0) * halfr;
    py[k] = (1.0 - yu / (safe * t)) * halfr;
  }
  double lo = c->vis_margin_px, hi = c->cam_res - c->vis_margin_px, xmin = px[0], xmax = px[0], ymin = py[0], ymax = py[0], sx = 0, sy2 = 0, a2 = 0, per = 0;
  for (int k = 0; k < 4; ++k) {
    if (!(px[k] >= lo && px[k] <= hi && py[k] >= lo && py[k] <= hi)) ok = 0;
    if (px[k] < xmin) xmin = px[k]; if (px[k] > xmax) xmax = px[k];
    if (py[k] < ymin) ymin = py[k]; if (py[k] > ymax) ymax = py[k];
    sx += px[k]; sy2 += py[k];
    int kn = (k + 1) & 3;
    a2 += px[k] * py[kn] - px[kn] * py[k];
    per += sqrt((px[kn] - px[k]) * (px[kn] - px[k]) + (py[kn] - py[k]) * (py[kn] - py[k]));
  }
  double w = floor(xmax - 0.5) - ceil(xmin - 0.5) + 1.0, hg = floor(ymax - 0.5) - ceil(ymin - 0.5) + 1.0;
  if (!(w >= 2.0 && hg >= 2.0)) ok = 0;
  double ar = 0.5 * fabs(a2) - 0.5 * per + 1.0;
  if (ar < 0) ar = 0;
  *vis = ok;
  cen[0] = ok ? (sx / 4 - 0.5) / halfr - 1.0 : 0.0;
  cen[1] = ok ? (sy2 / 4 - 0.5) / halfr - 1.0 : 0.0;
  *area = ok ? ar / (c->cam_res * c->cam_res) : 0.0;
  *ratio = ok ? (hg > 0 ? w / hg : 0.0) : 0.0;
}

/* ---- hover.py:224-272 --------------------------------------------------------- */
static double pymod(double x, double m) { return x - m * floor(x / m); }
static void compute_state(const OrcConfig* c, OrcEnv* e, double* obs) {
  int vis; double cen[2], area, ratio, q[4];
  for (int a = 0; a < 3; ++a) obs[a] = (pymod(e->s_euler[a] - e->prev_euler[a] + M_PI, 2 * M_PI) - M_PI) / c->agent_dt;
  euler_to_quat(e->s_euler, q);
  vision(c, e, &vis, cen, &area, &ratio);
  obs[3] = q[0]; obs[4] = q[1]; obs[5] = q[2]; obs[6] = q[3];
  obs[7] = cen[0]; obs[8] = cen[1]; obs[9] = e->prev_centre[0]; obs[10] = e->prev_centre[1];
  obs[11] = area; obs[12] = e->prev_area; obs[13] = vis ? 1.0 : 0.0; obs[14] = ratio; obs[15] = e->prev_ratio;
  for (int a = 0; a < 4; ++a) obs[16 + a] = e->action[a];
  e->prev_centre[0] = cen[0]; e->prev_centre[1] = cen[1]; e->prev_area = area; e->prev_ratio = ratio;
}

/* ---- hover.py:72-113 ----------------------------------------------------------- */
static void reset_env(const Orc* o, int64_t i, OrcEnv* e, double* obs) {
  const OrcConfig* c = &o->c;
  double pos[3] = {c->start_pos[0], c->start_pos[1], c->start_pos[2]}, rpy[3] = {c->start_rpy[0], c->start_rpy[1], c->start_rpy[2]};
  if (c->spawn_pos_noise != 0.0 || c->spawn_yaw_noise != 0.0) {
    uint32_t b[4];
    env_bits(o, i, 0, STREAM_SPAWN, e->rng_ctr, b);
    for (int a = 0; a < 3; ++a) pos[a] += c->spawn_pos_noise * (u01(b[a]) * 2.0 - 1.0);
    rpy[2] += c->spawn_yaw_noise * (u01(b[3]) * 2.0 - 1.0);
  }
  if (pos[2] < c->floor_z) pos[2] = c->floor_z;
  memcpy(e->pos, pos, sizeof(pos));
  euler_to_quat(rpy, e->quat);
  memset(e->vel, 0, sizeof(e->vel)); memset(e->omega, 0, sizeof(e->omega));
  for (int m = 0; m < 4; ++m) e->thr[m] = c->spawn_throttle;
  memset(e->pid_i, 0, sizeof(e->pid_i)); memset(e->pid_e, 0, sizeof(e->pid_e)); memset(e->cpid, 0, sizeof(e->cpid));
  e->contact = pos[2] <= c->floor_z;
  snapshot(e);
  e->step_count = 0; e->terminated = e->truncated = e->oob = e->on_floor = 0;
  memset(e->action, 0, sizeof(e->action));
  e->prev_centre[0] = e->prev_centre[1] = e->prev_area = e->prev_ratio = 0; e->ep_return = 0;
  double sp[4] = {0, 0, 0, 0}; /* QuadX.set_mode's preset: hold the current height (modes 2-4) / pose (mode 7) */
  if (c->flight_mode == 2 || c->flight_mode == 3 || c->flight_mode == 4) sp[3] = e->s_pos[2];
  if (c->flight_mode == 7) { sp[0] = e->s_pos[0]; sp[1] = e->s_pos[1]; sp[2] = e->s_euler[2]; sp[3] = e->s_pos[2]; }
  aviary_steps(o, i, e, sp, c->reset_idle_steps, STREAM_RESET, e->rng_ctr);
  memcpy(e->prev_euler, e->s_euler, sizeof(e->prev_euler));
  compute_state(c, e, obs);
}

/* ---- hover.py:334-358 + 274-332 + VecEnv auto-reset ------------------------------ */
static void step_env(Orc* o, int64_t i, const double* act, double* obs, double* reward, uint8_t* term, uint8_t* trunc, double* tobs,
                     double* acc_ret, int64_t* acc_len, int64_t* acc_done) {
  const OrcConfig* c = &o->c;
  OrcEnv* e = &o->e[i];
  memcpy(e->action, act, 4 * sizeof(double));
  double sp[4] = {act[0] * c->action_scale[0], act[1] * c->action_scale[1], act[2] * c->action_scale[2], act[3] * c->thrust_scale + c->thrust_bias};
  if (c->flight_mode == -1) for (int a = 0; a < 3; ++a) sp[a] = act[a] * c->thrust_scale + c->thrust_bias; /* four motor pwm commands */
  double r = -0.1;
  if (!(e->terminated || e->truncated)) aviary_steps(o, i, e, sp, c->env_step_ratio, STREAM_STEP, e->rng_ctr);
  e->rng_ctr += 1;
  compute_state(c, e, obs);
  int64_t k = e->step_count;
  if (k > c->max_steps) e->truncated = 1;
  if (sqrt(e->s_pos[0] * e->s_pos[0] + e->s_pos[1] * e->s_pos[1] + e->s_pos[2] * e->s_pos[2]) > c->flight_dome_size) { r = -100.0; e->oob = 1; e->terminated = 1; }
  if (!c->render && k > c->floor_grace_steps && e->s_pos[2] < c->floor_threshold) { r = -100.0; e->on_floor = 1; e->terminated = 1; }
  double target = obs[13] > 0.5 ? (-sqrt(obs[7] * obs[7] + obs[8] * obs[8])) + (-fabs(obs[11] - c->target_area)) + (-fabs(obs[14] - c->target_ratio)) : -2.0;
  double yr = fabs(e->s_wb[2]);
  r = r - 0.01 * yr * yr;
  r = r + (target - sqrt(e->s_euler[0] * e->s_euler[0] + e->s_euler[1] * e->s_euler[1]));
  double sm = 0;
  for (int a = 0; a < 4; ++a) sm += (e->action[a] - e->prev_action[a]) * (e->action[a] - e->prev_action[a]);
  r = r - sqrt(sm) * 0.2;
  r = r + 1.0;
  memcpy(e->prev_euler, e->s_euler, sizeof(e->prev_euler));
  e->step_count += 1;
  memcpy(e->prev_action, e->action, sizeof(e->action));
  *reward = r; *term = (uint8_t)e->terminated; *trunc = (uint8_t)e->truncated;
  e->ep_return += r;
  if (c->auto_reset && (e->terminated || e->truncated)) {
    *acc_ret += e->ep_return; *acc_len += e->step_count; *acc_done += 1;
    if (tobs) memcpy(tobs, obs, 20 * sizeof(double));
    reset_env(o, i, e, obs);
  }
}

/* ---- C API (ctypes: oracle/c_oracle.py) ------------------------------------------- */
int64_t orc_sizeof_config(void) { return (int64_t)sizeof(OrcConfig); }
int orc_max_threads(void) {
  const char* ev = getenv("ORC_THREADS");
  long n = ev ? atol(ev) : sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}
Orc* orc_create(const OrcConfig* c, int64_t n, uint64_t seed, uint64_t env_id0) {
  Orc* o = (Orc*)calloc(1, sizeof(Orc));
  o->c = *c; o->n = n; o->seed = seed; o->env_id0 = env_id0;
  o->e = (OrcEnv*)calloc((size_t)n, sizeof(OrcEnv));
  for (int64_t i = 0; i < n; ++i) o->e[i].quat[3] = 1.0;
  return o;
}
void orc_destroy(Orc* o) { if (o) { free(o->e); free(o); } }
typedef struct {
  Orc* o; int64_t lo, hi; int is_step;
  const uint8_t* mask; const double* actions; double *obs, *reward, *tobs; uint8_t *term, *trunc;
  double sr; int64_t sl, nd;
} Job;
static void* worker(void* arg) {
  Job* j = (Job*)arg;
  Orc* o = j->o;
  for (int64_t i = j->lo; i < j->hi; ++i) {
    if (j->is_step)
      step_env(o, i, j->actions + 4 * i, j->obs + 20 * i, j->reward + i, j->term + i, j->trunc + i, j->tobs ? j->tobs + 20 * i : NULL, &j->sr, &j->sl, &j->nd);
    else if (!j->mask || j->mask[i])
      reset_env(o, i, &o->e[i], j->obs + 20 * i);
  }
  return NULL;
}
static void run_jobs(Orc* o, Job proto) {
  int nt = orc_max_threads();
  if ((int64_t)nt > o->n) nt = (int)o->n;
  pthread_t th[256]; Job jobs[256];
  for (int t = 0; t < nt; ++t) {
    jobs[t] = proto; jobs[t].o = o; jobs[t].lo = o->n * t / nt; jobs[t].hi = o->n * (t + 1) / nt;
    jobs[t].sr = 0; jobs[t].sl = 0; jobs[t].nd = 0;
    if (t > 0) pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  worker(&jobs[0]);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
  for (int t = 0; t < nt; ++t) { o->sum_ret += jobs[t].sr; o->sum_len += jobs[t].sl; o->n_done += jobs[t].nd; }
}
void orc_reset(Orc* o, const uint8_t* mask, double* obs) {
  Job j; memset(&j, 0, sizeof(j)); j.is_step = 0; j.mask = mask; j.obs = obs;
  run_jobs(o, j);
}
void orc_step(Orc* o, const double* actions, double* obs, double* reward, uint8_t* term, uint8_t* trunc, double* tobs) {
  Job j; memset(&j, 0, sizeof(j)); j.is_step = 1; j.actions = actions; j.obs = obs; j.reward = reward; j.term = term; j.trunc = trunc; j.tobs = tobs;
  run_jobs(o, j);
}
void orc_stats(const Orc* o, double* sum_ret, int64_t* sum_len, int64_t* n_done) { *sum_ret = o->sum_ret; *sum_len = o->sum_len; *n_done = o->n_done; }
/* true rigid-body state for parity tests: pos3 quat4 vel3 omega3 thr4 = 17 doubles per env */
void orc_get_state(const Orc* o, double* out) {
  for (int64_t i = 0; i < o->n; ++i) {
    const OrcEnv* e = &o->e[i]; double* p = out + 17 * i;
    memcpy(p, e->pos, 24); memcpy(p + 3, e->quat, 32); memcpy(p + 7, e->vel, 24); memcpy(p + 10, e->omega, 24); memcpy(p + 13, e->thr, 32);
  }
}
