// qx_lanes.cuh -- lane-generic fp32 arithmetic for the env-step kernels.
//
// The physics / controller code of qx_model.cuh is written once over a value type T:
//   T = float   one env per thread (reset queue, small batches, step_k, flight modes, yaw)
//   T = float2  TWO envs per thread, one in each half of a 64-bit register pair, so that every plain
//               multiply / add / fma is ONE packed Blackwell instruction (FFMA2 / FMUL2 / FADD2, PTX *.f32x2)
//               for both envs.  Measured on B200 (tools/micro/ffma2_issue.cu): FFMA2 occupies the FMA pipe for
//               2.0 (r,imm,r / r,R.F32,r) to 2.5 (r,r,r) cycles, i.e. it does not raise the FLOP rate, but it halves
//               the issue slots the FP work needs -- and issue slots are what bound the scalar kernel.
// Both instantiations perform the same IEEE operations in the same order on each env (explicit fma / mul / add with
// round-to-nearest, no compiler contraction), so an env's trajectory does not depend on which one ran it.
//
// What does NOT pack: operand modifiers.  FFMA2 takes no |x| / -x modifiers (ptxas materialises them with extra FADDs),
// so |a|*b, min / max, selects, comparisons and the MUFU functions are issued per half, and subtraction is written
// fma(b, -1, a).  Neg<T> carries "a value whose negation is free": the value itself for float (the FFMA modifier), a
// negated copy for float2 (one FMUL2 serves every use).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qx {

struct m2 { bool x, y; };  // per-half predicate

#define QX_DI __device__ __forceinline__

// MUFU-only approximations (1-2 ulp): no Newton refinement, no slow paths
QX_DI float frcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
QX_DI float fsqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
QX_DI float frsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
QX_DI float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

template <class T> struct Lane;
template <> struct Lane<float> { typedef bool mask; typedef uint32_t u32; static constexpr int N = 1; };

// ---- T = float ---------------------------------------------------------------------------------------------------
QX_DI float vfma(float a, float b, float c) { return fmaf(a, b, c); }
QX_DI float vmul(float a, float b) { return __fmul_rn(a, b); }
QX_DI float vadd(float a, float b) { return __fadd_rn(a, b); }
QX_DI float vsub(float a, float b) { return __fsub_rn(a, b); }
QX_DI float vabsmul(float a, float b) { return __fmul_rn(fabsf(a), b); }  // |a| b
QX_DI float vmin(float a, float b) { return fminf(a, b); }
QX_DI float vmax(float a, float b) { return fmaxf(a, b); }
QX_DI float vabsmax(float a, float b) { return fmaxf(fabsf(a), fabsf(b)); }
QX_DI float vclamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
QX_DI float vrcp(float a) { return frcp(a); }
QX_DI float vsqrt(float a) { return fsqrt(a); }
QX_DI float vrsqrt(float a) { return frsqrt(a); }
QX_DI float vlg2(float a) { return __log2f(a); }
QX_DI float vsin(float a) { return __sinf(a); }
QX_DI float vcos(float a) { return __cosf(a); }
QX_DI bool vlt(float a, float b) { return a < b; }
QX_DI bool vgt(float a, float b) { return a > b; }
QX_DI bool vany(bool m) { return m; }
QX_DI bool mor(bool a, bool b) { return a || b; }
QX_DI float vsel(bool m, float a, float b) { return m ? a : b; }
template <class T> QX_DI T splat(float s);
template <> QX_DI float splat<float>(float s) { return s; }

// ---- T = two envs per thread ------------------------------------------------------------------------------------
// P2<true>: the halves live in an aligned register pair and plain arithmetic is packed (FFMA2 / FMUL2 / FADD2).
// P2<false>: same two-env code on scalar FFMA / FMUL / FADD (twice the FP instructions, half the FMA-pipe occupancy per
// instruction) -- kept to measure what the packing itself buys; bit-identical results.
template <bool PK>
struct P2 { float x, y; };
typedef P2<true> f2;
typedef P2<false> f2u;
template <bool PK> struct Lane<P2<PK>> { typedef m2 mask; typedef uint2 u32; static constexpr int N = 2; };
template <bool PK> QX_DI float2 asf2(P2<PK> a) { return make_float2(a.x, a.y); }
template <bool PK> QX_DI P2<PK> mk2p(float2 a) { return P2<PK>{a.x, a.y}; }
QX_DI f2 mk2(float x, float y) { return f2{x, y}; }
template <> QX_DI f2 splat<f2>(float s) { return f2{s, s}; }
template <> QX_DI f2u splat<f2u>(float s) { return f2u{s, s}; }

template <bool PK> QX_DI P2<PK> vfma(P2<PK> a, P2<PK> b, P2<PK> c) {
  if constexpr (PK) return mk2p<PK>(__ffma2_rn(asf2(a), asf2(b), asf2(c)));
  else return P2<PK>{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)};
}
template <bool PK> QX_DI P2<PK> vfma(P2<PK> a, float b, P2<PK> c) { return vfma(a, P2<PK>{b, b}, c); }
template <bool PK> QX_DI P2<PK> vfma(P2<PK> a, P2<PK> b, float c) { return vfma(a, b, P2<PK>{c, c}); }
template <bool PK> QX_DI P2<PK> vfma(P2<PK> a, float b, float c) { return vfma(a, P2<PK>{b, b}, P2<PK>{c, c}); }
template <bool PK> QX_DI P2<PK> vfma(float a, P2<PK> b, P2<PK> c) { return vfma(P2<PK>{a, a}, b, c); }
template <bool PK> QX_DI P2<PK> vmul(P2<PK> a, P2<PK> b) {
  if constexpr (PK) return mk2p<PK>(__fmul2_rn(asf2(a), asf2(b)));
  else return P2<PK>{__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)};
}
template <bool PK> QX_DI P2<PK> vmul(P2<PK> a, float b) { return vmul(a, P2<PK>{b, b}); }
template <bool PK> QX_DI P2<PK> vmul(float a, P2<PK> b) { return vmul(P2<PK>{a, a}, b); }
template <bool PK> QX_DI P2<PK> vadd(P2<PK> a, P2<PK> b) {
  if constexpr (PK) return mk2p<PK>(__fadd2_rn(asf2(a), asf2(b)));
  else return P2<PK>{__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)};
}
template <bool PK> QX_DI P2<PK> vadd(P2<PK> a, float b) { return vadd(a, P2<PK>{b, b}); }
template <bool PK> QX_DI P2<PK> vsub(P2<PK> a, P2<PK> b) {  // a - b, one rounding
  if constexpr (PK) return mk2p<PK>(__ffma2_rn(asf2(b), make_float2(-1.f, -1.f), asf2(a)));
  else return P2<PK>{__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)};
}
template <bool PK> QX_DI P2<PK> vsub(float a, P2<PK> b) { return vsub(P2<PK>{a, a}, b); }
template <bool PK> QX_DI P2<PK> vabsmul(P2<PK> a, P2<PK> b) { return P2<PK>{__fmul_rn(fabsf(a.x), b.x), __fmul_rn(fabsf(a.y), b.y)}; }
template <bool PK> QX_DI P2<PK> vabsmul(P2<PK> a, float b) { return P2<PK>{__fmul_rn(fabsf(a.x), b), __fmul_rn(fabsf(a.y), b)}; }
template <bool PK> QX_DI P2<PK> vmin(P2<PK> a, P2<PK> b) { return P2<PK>{fminf(a.x, b.x), fminf(a.y, b.y)}; }
template <bool PK> QX_DI P2<PK> vmax(P2<PK> a, P2<PK> b) { return P2<PK>{fmaxf(a.x, b.x), fmaxf(a.y, b.y)}; }
template <bool PK> QX_DI P2<PK> vabsmax(P2<PK> a, P2<PK> b) { return P2<PK>{fmaxf(fabsf(a.x), fabsf(b.x)), fmaxf(fabsf(a.y), fabsf(b.y))}; }
template <bool PK> QX_DI P2<PK> vclamp(P2<PK> v, float lo, float hi) { return P2<PK>{fminf(fmaxf(v.x, lo), hi), fminf(fmaxf(v.y, lo), hi)}; }
template <bool PK> QX_DI P2<PK> vrcp(P2<PK> a) { return P2<PK>{frcp(a.x), frcp(a.y)}; }
template <bool PK> QX_DI P2<PK> vsqrt(P2<PK> a) { return P2<PK>{fsqrt(a.x), fsqrt(a.y)}; }
template <bool PK> QX_DI P2<PK> vrsqrt(P2<PK> a) { return P2<PK>{frsqrt(a.x), frsqrt(a.y)}; }
template <bool PK> QX_DI P2<PK> vlg2(P2<PK> a) { return P2<PK>{__log2f(a.x), __log2f(a.y)}; }
template <bool PK> QX_DI P2<PK> vsin(P2<PK> a) { return P2<PK>{__sinf(a.x), __sinf(a.y)}; }
template <bool PK> QX_DI P2<PK> vcos(P2<PK> a) { return P2<PK>{__cosf(a.x), __cosf(a.y)}; }
template <bool PK> QX_DI m2 vlt(P2<PK> a, float b) { return m2{a.x < b, a.y < b}; }
template <bool PK> QX_DI m2 vgt(P2<PK> a, float b) { return m2{a.x > b, a.y > b}; }
QX_DI bool vany(m2 m) { return m.x || m.y; }
QX_DI m2 mor(m2 a, m2 b) { return m2{a.x || b.x, a.y || b.y}; }
template <bool PK> QX_DI P2<PK> vsel(m2 m, P2<PK> a, P2<PK> b) { return P2<PK>{m.x ? a.x : b.x, m.y ? a.y : b.y}; }
template <bool PK> QX_DI P2<PK> vsel(m2 m, P2<PK> a, float b) { return P2<PK>{m.x ? a.x : b, m.y ? a.y : b}; }
template <bool PK> QX_DI P2<PK> vsel(m2 m, float a, P2<PK> b) { return P2<PK>{m.x ? a : b.x, m.y ? a : b.y}; }

// ---- "negation is free" operands ------------------------------------------------------------------------------------
template <class T> struct Neg { T v; };
QX_DI Neg<float> mkneg(float a) { return Neg<float>{a}; }              // no instruction: the FFMA operand modifier
QX_DI Neg<f2u> mkneg(f2u a) { return Neg<f2u>{a}; }
QX_DI Neg<f2> mkneg(f2 a) { return Neg<f2>{vmul(a, -1.f)}; }           // one FMUL2 (exact), shared by every use
QX_DI float vfnma(float a, Neg<float> b, float c) { return fmaf(-a, b.v, c); }  // c - a b
QX_DI f2u vfnma(f2u a, Neg<f2u> b, f2u c) { return f2u{fmaf(-a.x, b.v.x, c.x), fmaf(-a.y, b.v.y, c.y)}; }
QX_DI f2 vfnma(f2 a, Neg<f2> b, f2 c) { return vfma(a, b.v, c); }
QX_DI float vfnma1(float a, Neg<float> b) { return fmaf(-a, b.v, 1.f); }           // 1 - a b
QX_DI f2u vfnma1(f2u a, Neg<f2u> b) { return f2u{fmaf(-a.x, b.v.x, 1.f), fmaf(-a.y, b.v.y, 1.f)}; }
QX_DI f2 vfnma1(f2 a, Neg<f2> b) { return vfma(a, b.v, 1.f); }
QX_DI float vfms(float a, float b, Neg<float> c) { return fmaf(a, b, -c.v); }     // a b - c
QX_DI f2u vfms(f2u a, f2u b, Neg<f2u> c) { return f2u{fmaf(a.x, b.x, -c.v.x), fmaf(a.y, b.y, -c.v.y)}; }
QX_DI f2 vfms(f2 a, f2 b, Neg<f2> c) { return vfma(a, b, c.v); }

}  // namespace qx
