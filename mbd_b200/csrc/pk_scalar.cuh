// pk_scalar.cuh — the scalar layer of the packed rollout kernel: every physics expression is written once against
// the operations below and instantiated for
//   T = float : one sample per thread  (host check build only)
//   T = f2    : TWO samples per thread in one 64-bit register pair; add / mul / fma map to the sm_100a packed
//               fp32 instructions FADD2 / FMUL2 / FFMA2 (PTX add/mul/fma.rn.f32x2): one issue slot and one
//               dependent-latency step (measured 4.6 cycles, scripts/probe/ffma2_probe.cu) do the work of two.
// Each packed operation is the IEEE round-to-nearest operation per component, so a packed rollout is bit-identical
// to two scalar rollouts (the arithmetic contract of include/mbd_fp32.h is unchanged).  Comparisons and selects are
// per component (unpack, scalar FSETP/FSEL, repack — ptxas keeps the halves in place).
//
// Two builds:  nvcc (device functions, f2 = packed register pair) and plain g++ (tests/host_pk: f2 = {float, float}
// emulation, used to check the templated physics against the CPU oracle without a GPU).
#pragma once

#include <math.h>
#include <stdint.h>

#include "mbd_fp32.h"

#if defined(__CUDACC__)
#define PK_FN __device__ __forceinline__
#define PK_MFN __device__ __forceinline__
#define PK_DEVICE 1
#else
#define PK_FN static inline
#define PK_MFN inline
#define PK_DEVICE 0
#endif

namespace mbd {
namespace pk {

// ---- f2 -------------------------------------------------------------------------------------------------
#if PK_DEVICE
struct f2 { unsigned long long v; };
PK_FN f2 mk2(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
PK_FN float lo(f2 x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); (void)b; return a; }
PK_FN float hi(f2 x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); (void)a; return b; }
PK_FN f2 mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
PK_FN f2 add(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
PK_FN f2 sub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
PK_FN f2 fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#else
struct f2 { float a, b; };
PK_FN f2 mk2(float a, float b) { f2 r; r.a = a; r.b = b; return r; }
PK_FN float lo(f2 x) { return x.a; }
PK_FN float hi(f2 x) { return x.b; }
PK_FN f2 mul(f2 a, f2 b) { return mk2(a.a * b.a, a.b * b.b); }
PK_FN f2 add(f2 a, f2 b) { return mk2(a.a + b.a, a.b + b.b); }
PK_FN f2 sub(f2 a, f2 b) { return mk2(a.a - b.a, a.b - b.b); }
PK_FN f2 fma(f2 a, f2 b, f2 c) { return mk2(fmaf(a.a, b.a, c.a), fmaf(a.b, b.b, c.b)); }
#endif
// ptxas 12.9 CONTRACTS a packed multiply whose only use is a packed add/sub into FFMA2 — even with the explicit .rn
// qualifiers, with -fmad=false on nvcc and on ptxas, across an opaque asm barrier, and when the product is written
// fma(a, b, -0) (checked on sm_100a; the scalar mul.rn.f32 + add.rn.f32 pair is never fused).  That silently changes
// rounding wherever the arithmetic contract has a separate multiply and add.  add_nf / sub_nf ("no fuse") do the
// addition on the two halves with scalar FADDs, which ptxas leaves alone; they are used wherever an operand of an
// add/sub is (or may be, through copies and selects) a packed product.  tests/test_pk_host.py::test_no_packed_contraction
// compares the packed instruction counts of the PTX and of the SASS of the physics.
#if PK_DEVICE
PK_FN f2 add_nf(f2 a, f2 b) {
  float r0, r1;
  asm("add.rn.f32 %0, %1, %2;" : "=f"(r0) : "f"(lo(a)), "f"(lo(b)));
  asm("add.rn.f32 %0, %1, %2;" : "=f"(r1) : "f"(hi(a)), "f"(hi(b)));
  return mk2(r0, r1);
}
PK_FN f2 sub_nf(f2 a, f2 b) {
  float r0, r1;
  asm("sub.rn.f32 %0, %1, %2;" : "=f"(r0) : "f"(lo(a)), "f"(lo(b)));
  asm("sub.rn.f32 %0, %1, %2;" : "=f"(r1) : "f"(hi(a)), "f"(hi(b)));
  return mk2(r0, r1);
}
#else
PK_FN f2 add_nf(f2 a, f2 b) { return add(a, b); }
PK_FN f2 sub_nf(f2 a, f2 b) { return sub(a, b); }
#endif
// unpack / negate / repack: ptxas folds it into the operand's negate modifier of FFMA2 / FADD2 / FMUL2
PK_FN f2 neg(f2 a) { return mk2(-lo(a), -hi(a)); }
PK_FN f2 abs_(f2 a) { return mk2(fabsf(lo(a)), fabsf(hi(a))); }

struct m2 { bool a, b; };
PK_FN m2 lt(f2 x, f2 y) { m2 m; m.a = lo(x) < lo(y); m.b = hi(x) < hi(y); return m; }
PK_FN m2 le(f2 x, f2 y) { m2 m; m.a = lo(x) <= lo(y); m.b = hi(x) <= hi(y); return m; }
PK_FN m2 gt(f2 x, f2 y) { m2 m; m.a = lo(x) > lo(y); m.b = hi(x) > hi(y); return m; }
PK_FN m2 ge(f2 x, f2 y) { m2 m; m.a = lo(x) >= lo(y); m.b = hi(x) >= hi(y); return m; }
PK_FN m2 eq(f2 x, f2 y) { m2 m; m.a = lo(x) == lo(y); m.b = hi(x) == hi(y); return m; }
PK_FN m2 mand(m2 p, m2 q) { m2 m; m.a = p.a && q.a; m.b = p.b && q.b; return m; }
PK_FN f2 sel(m2 m, f2 x, f2 y) { return mk2(m.a ? lo(x) : lo(y), m.b ? hi(x) : hi(y)); }

// ---- float (same names) -------------------------------------------------------------------------------
PK_FN float mul(float a, float b) { return a * b; }
PK_FN float add(float a, float b) { return a + b; }
PK_FN float sub(float a, float b) { return a - b; }
PK_FN float fma(float a, float b, float c) { return fmaf(a, b, c); }
PK_FN float add_nf(float a, float b) { return a + b; }
PK_FN float sub_nf(float a, float b) { return a - b; }
PK_FN float neg(float a) { return -a; }
PK_FN float abs_(float a) { return fabsf(a); }
PK_FN bool lt(float x, float y) { return x < y; }
PK_FN bool le(float x, float y) { return x <= y; }
PK_FN bool gt(float x, float y) { return x > y; }
PK_FN bool ge(float x, float y) { return x >= y; }
PK_FN bool eq(float x, float y) { return x == y; }
PK_FN bool mand(bool p, bool q) { return p && q; }
PK_FN float sel(bool m, float x, float y) { return m ? x : y; }

// broadcast of a compile-time / warp-uniform scalar
template <class T> struct Bc;
template <> struct Bc<float> { PK_MFN static float of(float c) { return c; } };
template <> struct Bc<f2> { PK_MFN static f2 of(float c) { return mk2(c, c); } };
template <class T> PK_FN T bc(float c) { return Bc<T>::of(c); }

// ---- correctly rounded division / reciprocal / square root (include/mbd_fp32.h: MBD_DIV / MBD_RCP / MBD_SQRT) ------
PK_FN float div_(float a, float b) { return MBD_DIV(a, b); }
PK_FN float rcp_(float x) { return MBD_RCP(x); }
PK_FN float sqrt_(float x) { return MBD_SQRT(x); }
#if PK_DEVICE
// the same MUFU seed + Newton FMAs as mbd_div_dev / mbd_rcp_dev / mbd_sqrt_dev, the FMAs packed
PK_FN f2 rcp_seed(f2 x) {
  float a, b;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(lo(x)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(hi(x)));
  return mk2(a, b);
}
PK_FN f2 rsqrt_seed(f2 x) {
  float a, b;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(lo(x)));
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(hi(x)));
  return mk2(a, b);
}
PK_FN f2 rcp_(f2 x) {
  f2 r = rcp_seed(x);
  f2 e = fma(x, r, bc<f2>(-1.0f));
  return fma(r, neg(e), r);
}
PK_FN f2 div_(f2 a, f2 b) {
  f2 r = rcp_seed(b);
  f2 e = fma(neg(b), r, bc<f2>(1.0f));
  r = fma(r, e, r);
  f2 q = mul(a, r);
  f2 rem = fma(neg(b), q, a);
  q = fma(r, rem, q);
  const f2 zero = bc<f2>(0.0f);
  return sel(eq(a, zero), sel(lt(b, zero), neg(a), a), q);
}
PK_FN f2 sqrt_(f2 x) {
  f2 r = rsqrt_seed(x);
  f2 s = mul(x, r);
  f2 h = mul(r, bc<f2>(0.5f));
  f2 e = fma(neg(s), s, x);
  f2 y = fma(e, h, s);
  return sel(eq(x, bc<f2>(0.0f)), x, y);
}
#else
PK_FN f2 rcp_(f2 x) { return mk2(1.0f / x.a, 1.0f / x.b); }
PK_FN f2 div_(f2 a, f2 b) { return mk2(a.a / b.a, a.b / b.b); }
PK_FN f2 sqrt_(f2 x) { return mk2(sqrtf(x.a), sqrtf(x.b)); }
#endif

// mbd_atan2f (include/mbd_fp32.h), operation for operation
template <class T>
PK_FN T atan2_(T y, T x) {
  const T zero = bc<T>(0.0f);
  T ax = abs_(x), ay = abs_(y);
  auto xg = gt(ax, ay);
  T mx = sel(xg, ax, ay);
  T mn = sel(xg, ay, ax);
  T t = sel(eq(mx, zero), zero, div_(mn, mx));
  T z = mul(t, t);
  T p = bc<T>(2.834064187e-03f);
  p = fma(p, z, bc<T>(-1.600502990e-02f));
  p = fma(p, z, bc<T>(4.258760810e-02f));
  p = fma(p, z, bc<T>(-7.495445758e-02f));
  p = fma(p, z, bc<T>(1.063675433e-01f));
  p = fma(p, z, bc<T>(-1.420257092e-01f));
  p = fma(p, z, bc<T>(1.999248415e-01f));
  p = fma(p, z, bc<T>(-3.333306611e-01f));
  p = fma(p, z, bc<T>(1.0f));
  T r = mul(t, p);
  r = sel(gt(ay, ax), sub_nf(bc<T>(MBD_HALF_PI_F), r), r);   // r is a product here
  r = sel(lt(x, zero), sub_nf(bc<T>(MBD_PI_F), r), r);
  r = sel(lt(y, zero), neg(r), r);
  return r;
}

template <class T> PK_FN T clamp_(T x, T lo_, T hi_) { return sel(lt(x, lo_), lo_, sel(gt(x, hi_), hi_, x)); }

}  // namespace pk
}  // namespace mbd
