// Per-tuple MM geometry, energy and analytic derivatives -- shared by the forward (K13) and
// backward (K14) kernels and, compiled as plain host C++, by the CPU-side math self-test.
//
// Arithmetic contract (reference src/grappa/models/internal_coordinates.py:150-210 and
// models/energy.py:8-56; derivation in SURVEY.md appendix A.5/A.6):
//   bond     r   = |x0 - x1|                                E = 1/2 k (r - eq)^2
//   angle    th  = atan2(|a x b|, a.b), a = x0-x1, b = x2-x1 E = 1/2 k (th - eq)^2
//   torsion  phi = atan2((n1 x n2).r21/|r21|, n1.n2),  n1 = r01 x r21, n2 = r21 x r23
//                                                           E = sum_{n=1..n_per} k_n cos(n phi)
// The reference adds randn*1e-5 noise to the three torsion displacement vectors
// (internal_coordinates.py:194-196); this implementation is noise-free (SURVEY.md 8c hazard 1).
// Torsion energies and forces never evaluate a trigonometric function: cos(n phi), sin(n phi)
// come from cos(phi), sin(phi) by the angle-addition recurrence; atan2f is only used when the
// caller asks for the angle itself ('x' field).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif

namespace gb {

struct V3 {
  float x, y, z;
};

GB_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
GB_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
GB_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
GB_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
GB_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GB_HD V3 cross(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
GB_HD float inv_sqrt(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU.RSQ (2 ulp); geometry never produces denormals
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}
// reciprocal: one MUFU.RCP on the device (<= 1 ulp), exact division on the host build
GB_HD float recip(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}

#define GB_MAX_PERIODICITY 6

// ---- bond ---------------------------------------------------------------------------------------
struct BondGeom {
  float r;      // bond length
  V3 d0;        // dr/dx0 (= -dr/dx1)
};
// Degenerate geometry follows the reference's autograd conventions instead of producing inf * 0 = NaN:
//   * coincident bond atoms: r = 0, zero derivative (torch.norm's subgradient at 0 is 0);
//   * exactly collinear angle (nitriles, alkynes, CO2-type molecules stored on an axis): theta = atan2(0, c) in {0, pi},
//     zero derivative (internal_coordinates.py:155-170: atan2's partials vanish at s = 0 and norm's backward is 0);
//   * torsion with three collinear atoms: the reference hides the singularity behind randn * 1e-5 noise
//     (internal_coordinates.py:194-196); the noise-free definition here is phi = atan2(0, 0) = 0 with zero derivative.
GB_HD BondGeom bond_geom(V3 x0, V3 x1) {
  V3 d = x0 - x1;
  float r2 = dot(d, d);
  float ir = r2 > 0.f ? inv_sqrt(r2) : 0.f;
  BondGeom g;
  g.r = r2 * ir;
  g.d0 = ir * d;
  return g;
}

// ---- angle --------------------------------------------------------------------------------------
struct AngleGeom {
  float theta;
  V3 d0, d2;    // dtheta/dx0, dtheta/dx2 ; dtheta/dx1 = -(d0 + d2)
};
GB_HD AngleGeom angle_geom(V3 x0, V3 x1, V3 x2) {
  V3 a = x0 - x1, b = x2 - x1;
  V3 n = cross(a, b);
  float s2 = dot(n, n);
  float is = s2 > 0.f ? inv_sqrt(s2) : 0.f;   // collinear: s = 0, d0 = d2 = 0
  float s = s2 * is;              // |a||b| sin(theta)
  float c = dot(a, b);            // |a||b| cos(theta)
  float a2 = dot(a, a), b2 = dot(b, b);
  float ia2 = a2 > 0.f ? recip(a2) : 0.f, ib2 = b2 > 0.f ? recip(b2) : 0.f;
  AngleGeom g;
  g.theta = atan2f(s, c);
  // dtheta/da = (c/|a|^2 a - b) / s ,  dtheta/db = (c/|b|^2 b - a) / s
  g.d0 = is * ((c * ia2) * a - b);
  g.d2 = is * ((c * ib2) * b - a);
  return g;
}

// ---- torsion ------------------------------------------------------------------------------------
struct TorsionGeom {
  float cphi, sphi;     // cos / sin of the dihedral (sign convention of the reference)
  V3 d0, d1, d2, d3;    // dphi/dx_i
};
GB_HD TorsionGeom torsion_geom(V3 x0, V3 x1, V3 x2, V3 x3) {
  V3 F = x0 - x1, G = x1 - x2, H = x3 - x2;
  V3 A = cross(F, G), B = cross(H, G);
  float A2 = dot(A, A), B2 = dot(B, B), G2 = dot(G, G);
  float iG = G2 > 0.f ? inv_sqrt(G2) : 0.f;
  float gl = G2 * iG;                      // |G|
  float AB2 = A2 * B2;
  const bool ok = AB2 > 0.f;               // false: three collinear atoms -> phi = 0, zero derivative
  float iA2 = ok ? recip(A2) : 0.f, iB2 = ok ? recip(B2) : 0.f;
  float iAB = ok ? inv_sqrt(AB2) : 0.f;
  TorsionGeom t;
  t.cphi = ok ? dot(A, B) * iAB : 1.f;
  // (A x B).G = ((F x G) x (H x G)).G = -[F,G,H] |G|^2 = -(A.H) |G|^2   (vector quadruple product)
  t.sphi = -dot(A, H) * gl * iAB;
  float fg = dot(F, G) * iG, hg = dot(H, G) * iG;
  t.d0 = (gl * iA2) * A;
  t.d3 = (-gl * iB2) * B;
  V3 u = (fg * iA2) * A - (hg * iB2) * B;
  t.d1 = v3(-t.d0.x - u.x, -t.d0.y - u.y, -t.d0.z - u.z);
  t.d2 = v3(-t.d3.x + u.x, -t.d3.y + u.y, -t.d3.z + u.z);
  return t;
}

// E = sum k_n cos(n phi), dE/dphi = -sum n k_n sin(n phi); cn/sn receive cos/sin(n phi) if non-null.
template <int NPER>
GB_HD void torsion_series(const float* k, float cphi, float sphi, float& e, float& dedphi,
                          float* cn_out, float* sn_out) {
  float cn = cphi, sn = sphi;
  e = 0.f;
  dedphi = 0.f;
#pragma unroll
  for (int n = 1; n <= NPER; ++n) {
    e += k[n - 1] * cn;
    dedphi -= float(n) * k[n - 1] * sn;
    if (cn_out) { cn_out[n - 1] = cn; sn_out[n - 1] = sn; }
    float c2 = cn * cphi - sn * sphi;
    sn = sn * cphi + cn * sphi;
    cn = c2;
  }
}

GB_HD void torsion_series_dyn(int nper, const float* k, float cphi, float sphi, float& e, float& dedphi,
                              float* cn_out, float* sn_out) {
  float cn = cphi, sn = sphi;
  e = 0.f;
  dedphi = 0.f;
  for (int n = 1; n <= nper; ++n) {
    e += k[n - 1] * cn;
    dedphi -= float(n) * k[n - 1] * sn;
    if (cn_out) { cn_out[n - 1] = cn; sn_out[n - 1] = sn; }
    float c2 = cn * cphi - sn * sphi;
    sn = sn * cphi + cn * sphi;
    cn = c2;
  }
}

}  // namespace gb
