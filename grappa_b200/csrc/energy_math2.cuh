// Packed-pair (f32x2) restatement of energy_math.cuh for kernel K13: every lane carries TWO conformations in one 64-bit
// register and the geometry runs on Blackwell's packed fp32 pipe (PTX add / sub / mul / fma .f32x2 -> SASS FADD2 / FMUL2 /
// FFMA2): one issue slot per two FMAs.  K13 is bound by instruction issue, not by HBM (profiles/r1_summary.md section
// 11), so halving the instruction stream is the lever; reciprocals / rsqrt stay scalar MUFU ops per half.
//
// Same arithmetic contract and reference lines as energy_math.cuh (internal_coordinates.py:150-210, energy.py:8-56),
// same degenerate-geometry conventions.  The bond angle needs theta itself: atan2 is evaluated here by a packed
// polynomial (|error| < 1e-7 rad, fitted on [0, 1] in t^2; CUDA's atan2f costs ~40 scalar instructions per value).
#pragma once
#include "energy_math.cuh"
#include "f32x2.cuh"

namespace gb {

// per half: x > 0 ? rsqrt(x) : 0   /   x > 0 ? 1 / x : 0
__device__ __forceinline__ F2 rsqrt_pos(F2 a) {
  const float x = lo(a), y = hi(a);
  return f2(x > 0.f ? inv_sqrt(x) : 0.f, y > 0.f ? inv_sqrt(y) : 0.f);
}
__device__ __forceinline__ F2 rcp2(F2 a) { return f2(recip(lo(a)), recip(hi(a))); }

struct W3 {
  F2 x, y, z;
};
__device__ __forceinline__ W3 operator-(W3 a, W3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ W3 operator*(F2 s, W3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ F2 dot(W3 a, W3 b) { return fma2(a.z, b.z, fma2(a.y, b.y, a.x * b.x)); }
// products and differences are written as mul / sub: ptxas contracts them into FFMA2 with a negated operand (an explicit
// sign flip of a packed pair costs two LOP3)
__device__ __forceinline__ W3 cross(W3 a, W3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// atan(t) = t * Q(t^2) on [0, 1] (degree-8 fit in t^2, |err| < 1e-7 in fp32)
__device__ __forceinline__ F2 atan_unit(F2 t) {
  const F2 u = t * t;
  F2 q = f2(2.402711134e-03f);
  q = fma2(q, u, f2(-1.413175225e-02f));
  q = fma2(q, u, f2(3.921329169e-02f));
  q = fma2(q, u, f2(-7.169886420e-02f));
  q = fma2(q, u, f2(1.045559775e-01f));
  q = fma2(q, u, f2(-1.414435524e-01f));
  q = fma2(q, u, f2(1.998228608e-01f));
  q = fma2(q, u, f2(-3.333222689e-01f));
  q = fma2(q, u, f2(9.999997641e-01f));
  return q * t;
}
__device__ __forceinline__ float atan2_half(float y, float x, float ay, float ax, float a) {
  a = ay > ax ? 1.5707963267948966f - a : a;
  a = x < 0.f ? 3.14159265358979323846f - a : a;
  return copysignf(a, y);
}
// atan2(y, x), both halves, any quadrant (reflections of atan on [0, 1]); atan2(0, 0) = 0 like torch / libm
__device__ __forceinline__ F2 atan2_2(F2 y, F2 x) {
  const float y0 = lo(y), y1 = hi(y), x0 = lo(x), x1 = hi(x);
  const float ay0 = fabsf(y0), ay1 = fabsf(y1), ax0 = fabsf(x0), ax1 = fabsf(x1);
  const float mx0 = fmaxf(ay0, ax0), mx1 = fmaxf(ay1, ax1);
  const F2 t = f2(fminf(ay0, ax0), fminf(ay1, ax1)) * f2(mx0 > 0.f ? recip(mx0) : 0.f, mx1 > 0.f ? recip(mx1) : 0.f);
  const F2 a = atan_unit(t);
  return f2(atan2_half(y0, x0, ay0, ax0, lo(a)), atan2_half(y1, x1, ay1, ax1, hi(a)));
}

// ---- bond: r, dr/dx0 (energy_math.cuh::bond_geom) -----------------------------------------------------------------
struct BondGeom2 {
  F2 r;
  W3 d0;
};
__device__ __forceinline__ BondGeom2 bond_geom2(W3 x0, W3 x1) {
  const W3 d = x0 - x1;
  const F2 r2 = dot(d, d);
  const F2 ir = rsqrt_pos(r2);
  return {r2 * ir, ir * d};
}

// ---- angle: theta, dtheta/dx0, dtheta/dx2 (energy_math.cuh::angle_geom) -------------------------------------------
struct AngleGeom2 {
  F2 theta;
  W3 d0, d2;
};
__device__ __forceinline__ AngleGeom2 angle_geom2(W3 x0, W3 x1, W3 x2) {
  const W3 a = x0 - x1, b = x2 - x1;
  const W3 n = cross(a, b);
  const F2 s2 = dot(n, n);
  const F2 is = rsqrt_pos(s2);            // collinear: s = 0, d0 = d2 = 0
  const F2 s = s2 * is;
  const F2 c = dot(a, b);
  const F2 a2 = dot(a, a), b2 = dot(b, b);
  const F2 ia2 = f2(lo(a2) > 0.f ? recip(lo(a2)) : 0.f, hi(a2) > 0.f ? recip(hi(a2)) : 0.f);
  const F2 ib2 = f2(lo(b2) > 0.f ? recip(lo(b2)) : 0.f, hi(b2) > 0.f ? recip(hi(b2)) : 0.f);
  AngleGeom2 g;
  g.theta = atan2_2(s, c);
  const F2 ca = c * ia2, cb = c * ib2;
  g.d0 = is * W3{ca * a.x - b.x, ca * a.y - b.y, ca * a.z - b.z};
  g.d2 = is * W3{cb * b.x - a.x, cb * b.y - a.y, cb * b.z - a.z};
  return g;
}

// ---- torsion: cos / sin of the dihedral and dphi/dx_i (energy_math.cuh::torsion_geom) -------------------------------
// Sign-free form: with Hn = x2 - x3 = -H,  B = G x Hn (= H x G),  sphi = (A.Hn) |G| / (|A||B|),  and
//     dphi/dx0 = p0,  dphi/dx3 = -p3,  dphi/dx1 = -(p0 + u),  dphi/dx2 = p3 + u
// with p0 = |G|/|A|^2 A,  p3 = |G|/|B|^2 B,  u = (F.G)/(|G||A|^2) A + (Hn.G)/(|G||B|^2) B  -- the caller applies the signs
// as add / subtract, so no packed sign flips are issued.
struct TorsionGeom2 {
  F2 cphi, sphi;
  W3 p0, p3, u;
};
__device__ __forceinline__ TorsionGeom2 torsion_geom2(W3 x0, W3 x1, W3 x2, W3 x3) {
  const W3 F = x0 - x1, G = x1 - x2, Hn = x2 - x3;
  const W3 A = cross(F, G), B = cross(G, Hn);
  const F2 A2 = dot(A, A), B2 = dot(B, B), G2 = dot(G, G);
  const F2 iG = rsqrt_pos(G2);
  const F2 gl = G2 * iG;
  const F2 AB2 = A2 * B2;
  const bool ok0 = lo(AB2) > 0.f, ok1 = hi(AB2) > 0.f;   // false: three collinear atoms -> phi = 0, zero derivative
  const F2 iA2 = f2(ok0 ? recip(lo(A2)) : 0.f, ok1 ? recip(hi(A2)) : 0.f);
  const F2 iB2 = f2(ok0 ? recip(lo(B2)) : 0.f, ok1 ? recip(hi(B2)) : 0.f);
  const F2 iAB = f2(ok0 ? inv_sqrt(lo(AB2)) : 0.f, ok1 ? inv_sqrt(hi(AB2)) : 0.f);
  TorsionGeom2 t;
  const F2 cr = dot(A, B) * iAB;
  t.cphi = f2(ok0 ? lo(cr) : 1.f, ok1 ? hi(cr) : 1.f);
  t.sphi = dot(A, Hn) * (gl * iAB);
  const F2 ua = dot(F, G) * (iG * iA2), ub = dot(Hn, G) * (iG * iB2);
  t.p0 = (gl * iA2) * A;
  t.p3 = (gl * iB2) * B;
  t.u = {fma2(ua, A.x, ub * B.x), fma2(ua, A.y, ub * B.y), fma2(ua, A.z, ub * B.z)};
  return t;
}

// E = sum k_n cos(n phi), dE/dphi = -sum n k_n sin(n phi); k (and n k) are uniform over the pair
template <int NPER>
__device__ __forceinline__ void torsion_series2(const float* k, F2 cphi, F2 sphi, F2& e, F2& dedphi) {
  F2 cn = cphi, sn = sphi;
  e = f2(0.f);
  dedphi = f2(0.f);
#pragma unroll
  for (int n = 1; n <= NPER; ++n) {
    e = fma2(f2(k[n - 1]), cn, e);
    dedphi = fma2(f2(-float(n) * k[n - 1]), sn, dedphi);
    if (n < NPER) {
      const F2 c2 = cn * cphi - sn * sphi;
      sn = fma2(sn, cphi, cn * sphi);
      cn = c2;
    }
  }
}

}  // namespace gb
