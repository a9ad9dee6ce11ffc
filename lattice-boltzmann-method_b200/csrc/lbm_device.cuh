// lbm_device.cuh — D2Q9 constants and per-node arithmetic shared by every kernel.
//
// Direction order, velocities and weights are the reference's solver::c / solver::E
// (src/solver.cpp:12-21): 0,(1,0),(0,1),(-1,0),(0,-1),(1,1),(-1,1),(-1,-1),(1,-1).
// Direction 1 moves along axis 0 (x, "rows"); axis 1 (y, "columns") is the contiguous one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lbm
{

constexpr int Q = 9;

// Kept as functions of a compile-time-unrolled index so they fold to immediates.
__host__ __device__ __forceinline__ constexpr double W(int q)
{
  return q == 0 ? 4.0 / 9.0 : (q < 5 ? 1.0 / 9.0 : 1.0 / 36.0);
}
__host__ __device__ __forceinline__ constexpr int CX(int q)
{
  return (q == 1 || q == 5 || q == 8) ? 1 : ((q == 3 || q == 6 || q == 7) ? -1 : 0);
}
__host__ __device__ __forceinline__ constexpr int CY(int q)
{
  return (q == 2 || q == 5 || q == 6) ? 1 : ((q == 4 || q == 7 || q == 8) ? -1 : 0);
}
__host__ __device__ __forceinline__ constexpr int OPP(int q)
{
  return q == 0 ? 0 : (q == 1 ? 3 : (q == 2 ? 4 : (q == 3 ? 1 : (q == 4 ? 2 : (q == 5 ? 7 : (q == 6 ? 8 : (q == 7 ? 5 : 6)))))));
}

enum Mode
{
  MODE_LOCAL = 0,     // source buffer holds post-stream populations (first step after lbm_set_f)
  MODE_PULL = 1,      // source buffer holds post-collision populations: pull + BC, then collide
  MODE_PULL_ONLY = 2  // pull + BC only, write post-stream populations in the reference's AoS layout
};

// EQ_KBC selects the entropic central-moment collision of ulbm::d2q9::kbc (src/ulbm.cpp) and its
// product-form equilibrium; it is the third "equilibrium kind" of the single-phase kernel family.
enum EqKind { EQ_COMP = 0, EQ_INCOMP = 1, EQ_KBC = 2 };
// FORCE_REGION: the FORCE_IBM data path (a force field on a rectangle) with the source-term constants read from the
// parameters instead of folded in (lbm_set_force_region); internal, selected when the field is a fixed one
enum ForceKind { FORCE_NONE = 0, FORCE_UNIFORM = 1, FORCE_IBM = 2, FORCE_REGION = 3 };

// Everything a step kernel needs to know about one slab.
struct SlabGeom
{
  int Xl;          // rows owned by this slab
  int Y;           // columns
  int pitch;       // row pitch in doubles (multiple of 16 => 128-byte aligned rows)
  int xg0;         // global index of local row 0
  long long plane; // (Xl + 2) * pitch : one population plane incl. the two ghost rows
};

__device__ __forceinline__ long long node_off(const SlabGeom& g, int x, int y)
{
  return (long long)(x + 1) * g.pitch + y;  // ghost row below row 0 sits at storage row 0
}

struct BgkParams
{
  double omega;    // fluid relaxation (params::lattice::omega)
  double inv_omega;  // 1 / omega (the KBC collision's 1 / s2)
  double omega_g;  // ADE lattice relaxation
  double Fg0, Fg1; // uniform force (test/gravity_test.cpp:85)
  double w_s;      // settling velocity (test/rectangle_sedimentation_test.cpp:89)
  // immersed-boundary force field on the ROI (test/cylinder_test.cpp:110-127), global coordinates
  int roi_r0, roi_r1, roi_c0, roi_c1;
  const double* Fx;
  const double* Fy;
  double ics2, ics4;  // FORCE_REGION: constants of the source term as the driver names them, 3 and 9 in
                      // decompose_domain_loop.cpp:68-69 (1/3, 1/9 of gravity_test / cylinder_test are compile-time)
  // LBM_MODEL_KBC, first step after an import: rho {Xl,Y}, u {Xl,Y,2} supplied by the caller (the drivers' m0 / m1
  // members, test/ulbm_poiseuille.cpp:93) instead of the moments of the imported populations; nullptr otherwise
  const double* mom_in_rho;
  const double* mom_in_u;
  int Y;           // row length of mom_in_* / snap_*
  // MODE_PULL_ONLY: when set, the pass writes rho {Xl,Y} and u {Xl,Y,2} of the post-stream state (what the next
  // iteration of the driver would compute: calc_rho, calc_u / calc_incomp_u, += Fg) instead of the populations
  double* snap_rho;
  double* snap_u;
};

// rho = sum_q f, (jx, jy) = sum_q f c_q in the q order of the reference's reductions
// (solver::calc_rho / calc_incomp_u, src/solver.cpp:23-31).
__host__ __device__ __forceinline__ void moments(const double (&f)[9], double& rho, double& jx, double& jy)
{
  rho = ((((((((f[0] + f[1]) + f[2]) + f[3]) + f[4]) + f[5]) + f[6]) + f[7]) + f[8]);
  jx = (((((f[1] - f[3]) + f[5]) - f[6]) - f[7]) + f[8]);
  jy = (((((f[2] - f[4]) + f[5]) + f[6]) - f[7]) - f[8]);
}

// rho, u of a post-stream node with the conventions of the model's driver (lbm_get_moments)
template <int EQ, int FORCE>
__host__ __device__ __forceinline__ void snapshot_moments(const double (&f)[9], const BgkParams& p, double& rho, double& ux, double& uy)
{
  double jx, jy;
  moments(f, rho, jx, jy);
  ux = EQ == EQ_INCOMP ? jx : jx / rho;
  uy = EQ == EQ_INCOMP ? jy : jy / rho;
  if constexpr (FORCE == FORCE_UNIFORM)
  {
    ux += p.Fg0;
    uy += p.Fg1;
  }
}

// solver::equilibrium (src/solver.cpp:51-62)
__host__ __device__ __forceinline__ double feq_comp(int q, double rho, double ux, double uy, double uu)
{
  const double cu = (double)CX(q) * ux + (double)CY(q) * uy;
  const double A = 1.0 + 3.0 * cu + 4.5 * (cu * cu) - 1.5 * uu;
  return (rho * A) * W(q);
}

// solver::incomp_equilibrium (src/solver.cpp:39-49)
__host__ __device__ __forceinline__ double feq_incomp(int q, double rho, double ux, double uy)
{
  const double cu = (double)CX(q) * ux + (double)CY(q) * uy;
  return (rho + 3.0 * cu) * W(q);
}

__host__ __device__ __forceinline__ double feq_kbc_q(int q, double rho, double ux, double uy);

template <int EQ>
__host__ __device__ __forceinline__ double feq_any(int q, double rho, double ux, double uy, double uu)
{
  if constexpr (EQ == EQ_COMP) return feq_comp(q, rho, ux, uy, uu);
  else if constexpr (EQ == EQ_KBC) return feq_kbc_q(q, rho, ux, uy);
  else return feq_incomp(q, rho, ux, uy);
}

// The nine polynomial factors of kbc::eval_equilibrium / eval_iequilibrium (src/ulbm.cpp:230-240,250-258);
// ux2, uy2 are arguments because the reference keeps them as members that only collide() refreshes.
__host__ __device__ __forceinline__ void kbc_eq_coef(double ux, double uy, double ux2, double uy2, double (&e)[9])
{
  constexpr double cs2 = 1.0 / 3.0, cs4 = 1.0 / 9.0;
  e[0] = 2.0 * cs2 * (0.5 * ux2 + 0.5 * uy2 - 1.0) + cs4 + ux2 * uy2 - ux2 - uy2 + 1.0;
  e[1] = 0.5 * (-cs2 * (ux2 + uy2 + ux - 1.0) - cs4 - ux2 * uy2 + ux2 - uy2 * ux + ux);
  e[2] = 0.5 * (-cs2 * (ux2 + uy2 + uy - 1.0) - cs4 - ux2 * uy2 - ux2 * uy + uy2 + uy);
  e[3] = 0.5 * (-cs2 * (ux2 + uy2 - ux - 1.0) - cs4 - ux2 * uy2 + ux2 + uy2 * ux - ux);
  e[4] = 0.5 * (-cs2 * (ux2 + uy2 - uy - 1.0) - cs4 - ux2 * uy2 + ux2 * uy + uy2 - uy);
  e[5] = 0.25 * (cs2 * (ux2 + uy2 + ux + uy) + cs4 + ux2 * uy2 + ux2 * uy + uy2 * ux + ux * uy);
  e[6] = 0.25 * (cs2 * (ux2 + uy2 - ux + uy) + cs4 + ux2 * uy2 + ux2 * uy - uy2 * ux - ux * uy);
  e[7] = 0.25 * (cs2 * (ux2 + uy2 - ux - uy) + cs4 + ux2 * uy2 - ux2 * uy - uy2 * ux + ux * uy);
  e[8] = 0.25 * (cs2 * (ux2 + uy2 + ux - uy) + cs4 + ux2 * uy2 - ux2 * uy + uy2 * ux - ux * uy);
}

// f_equi as test/ulbm_poiseuille.cpp:117 hands it to its pressure rule: kbc.iequi_f.pow(-1), i.e. the
// reciprocal of the stored reciprocal 1 / (e_q m0)
__host__ __device__ __forceinline__ double feq_kbc_q(int q, double rho, double ux, double uy)
{
  double e[9];
  kbc_eq_coef(ux, uy, ux * ux, uy * uy, e);
  double r = 0.0;
#pragma unroll
  for (int k = 0; k < 9; k++)
    if (k == q) r = 1.0 / (1.0 / (e[k] * rho));
  return r;
}

// M^-1 of kbc::collide() step 4 (src/ulbm.cpp:114-122; the sign flip of :123 is left to the caller)
__host__ __device__ __forceinline__ void kbc_minv(const double (&g)[9], double (&c)[9])
{
  // the halves and quarters are exact scalings, taken once; each output is then a sum of a pair term and a shared term
  const double h1 = 0.5 * g[1], h2 = 0.5 * g[2], h6 = 0.5 * g[6], h7 = 0.5 * g[7], h8 = 0.5 * g[8];
  const double q3 = 0.25 * g[3], q4 = 0.25 * g[4];
  const double a8 = (q3 + q4) - h8, b8 = (q3 - q4) - h8;
  const double d17 = h1 - h7, d26 = h2 - h6;
  c[0] = g[0] - g[3] + g[8];
  c[1] = a8 + d17;
  c[3] = a8 - d17;
  c[2] = b8 + d26;
  c[4] = b8 - d26;
  const double p58 = 0.25 * (g[5] + g[8]), m58 = 0.25 * (g[8] - g[5]);
  const double p67 = 0.25 * (g[6] + g[7]), m67 = 0.25 * (g[6] - g[7]);
  c[5] = p58 + p67;
  c[7] = p58 - p67;
  c[6] = m58 + m67;
  c[8] = m58 - m67;
}

// kbc::collide() of one node (src/ulbm.cpp:91-126: eval_central_momenta :264-320, eval_gamma :138-148,
// eval_delta_s :157-189, eval_delta_h :191-224, eval_iequilibrium :226-244, then S (cT - cT_eq), N^-1, -M^-1,
// + adve_f).  In: post-stream f and the iteration's m0, u.  Out: post-collision f.
//
// The reference spells delta_s, delta_h and 1/f_eq out as 27 polynomials and sums nine weighted products per
// central moment (~630 fp64 instructions per node: fp64-bound at half the HBM roofline).  Here they are
// evaluated through their structure, to rounding-level differences (T = M^-1 N^-1, the transform of steps 3-4):
//   central moments  from the raw moments of opposite-direction pairs, shifted by u (binomial expansion)
//   delta_s = T (-m0, 0, 0, cT3 - 2 cs2 m0, cT4, cT5, 0, 0, -cs4 m0)
//   delta_h = T (-m0, 0, 0, -2 cs2 m0, 0, 0, cT6, cT7, cT8 - cs4 m0) + the reference's `ux2+uy` slips in entries
//             5..8 (src/ulbm.cpp:211-223 adds ux2 and uy where the expansion has ux2*uy), kept as written
//   collision = s2 T (0,0,0, cT3 - 2 cs2 m0, cT4, cT5, 0,0,0) + gamma s2 T (0,..,0, cT6, cT7, cT8 - cs4 m0)
//             [+ T (cT0 - m0, cT1, cT2, 0,..) when the caller supplied m0, u: zero to rounding otherwise]
//   so N^-1 runs once on the shear part P, once on the higher-order part H and on the two equilibrium-only
//   vectors; delta_s, delta_h and the collision are combinations of those, and M^-1 runs three times;
//   f_eq,q = m0 phi_x(c_qx) phi_y(c_qy), phi(0) = 1 - cs2 - u^2, phi(+-1) = (cs2 + u^2 +- u) / 2 (product form of
//             :230-238); in gamma's quotient m0 and the common denominator cancel: no reciprocal at all.
// given = false: m0, ux, uy are OUTPUTS — the populations' own moments (kbc.m0 = adve_f.sum, kbc.m1 = adve_f c^T / m0,
// test/ulbm_double_shear_flow.cpp:143-146), taken from the pair sums the central moments need anyway.
__host__ __device__ __forceinline__ void kbc_collide(double (&f)[9], double s2, double is2, double& m0, double& ux, double& uy, bool given)
{
  constexpr double cs2 = 1.0 / 3.0, cs4 = 1.0 / 9.0;  // is2 = 1 / s2, taken once on the host
  // ---- raw moments m_ab = sum f cx^a cy^b from pair sums / differences, then central moments
  const double A1 = f[1] + f[3], D1 = f[1] - f[3], A2 = f[2] + f[4], D2 = f[2] - f[4];
  const double A57 = f[5] + f[7], D57 = f[5] - f[7], A68 = f[6] + f[8], D68 = f[6] - f[8];
  const double m22 = A57 + A68, m11 = A57 - A68, m21 = D57 + D68, m12 = D57 - D68;
  const double m20 = A1 + m22, m02 = A2 + m22;
  const double m10 = D1 + m12, m01 = D2 + m21;
  const double r00 = ((f[0] + A1) + A2) + m22;
  if (!given)
  {
    const double ir = 1.0 / r00;
    m0 = r00;
    ux = m10 * ir;
    uy = m01 * ir;
  }
  const double ux2 = ux * ux, uy2 = uy * uy, uxy = ux * uy;
  // shift to the node's velocity one axis at a time (binomial transform along x, then along y): 16 fused
  // multiply-adds for the eight central moments instead of the ~40 operations of the expanded polynomials.
  // x_ab = sum f (cx - ux)^a cy^b.  With the populations' own m0, u the first-order moments are zero by definition.
  const double x10 = given ? m10 - ux * r00 : 0.0;
  const double x11 = m11 - ux * m01, x12 = m12 - ux * m02;
  const double x20 = (m20 - ux * m10) - ux * x10, x21 = (m21 - ux * m11) - ux * x11, x22 = (m22 - ux * m12) - ux * x12;
  const double k10 = x10;
  const double k01 = given ? m01 - uy * r00 : 0.0;
  const double k11 = x11 - uy * x10, k21 = x21 - uy * x20;
  const double k02 = (m02 - uy * m01) - uy * k01, k12 = (x12 - uy * x11) - uy * k11, k22 = (x22 - uy * x21) - uy * k21;
  const double k20 = x20;
  const double K3 = (k20 + k02) - 2.0 * cs2 * m0, C4 = k20 - k02, C5 = k11, C6 = k21, C7 = k12, K8 = k22 - cs4 * m0;

  // ---- N^-1 (src/ulbm.cpp:104-112) on P = (0,0,0,K3,C4,C5,0,0,0) and H = (0,...,0,C6,C7,K8)
  double gP[9], gH[9];
  gP[0] = gP[1] = gP[2] = 0.0;
  gP[3] = K3; gP[4] = C4; gP[5] = C5;
  gP[6] = 0.5 * (K3 + C4) * uy + 2.0 * C5 * ux;
  gP[7] = 0.5 * (K3 - C4) * ux + 2.0 * C5 * uy;
  gP[8] = (0.5 * K3 * (ux2 + uy2) - 0.5 * C4 * (ux2 - uy2)) + 4.0 * C5 * uxy;
  gH[0] = gH[1] = gH[2] = gH[3] = gH[4] = gH[5] = 0.0;
  gH[6] = C6; gH[7] = C7;
  gH[8] = (2.0 * C6 * uy + 2.0 * C7 * ux) + K8;
  // equilibrium-only vectors E0 = (-m0,0,..,0,-cs4 m0) (in delta_s) and E1 = (-m0,0,0,-2 cs2 m0,0,..,0) (in delta_h)
  const double n0 = -m0;
  double gE[9];  // N^-1 (-m0, 0, .., 0): common to both
  gE[0] = n0; gE[1] = n0 * ux; gE[2] = n0 * uy; gE[3] = n0 * (ux2 + uy2); gE[4] = n0 * (ux2 - uy2); gE[5] = n0 * uxy;
  gE[6] = gE[1] * uxy; gE[7] = gE[2] * uxy; gE[8] = n0 * ux2 * uy2;
  const double e3 = -2.0 * cs2 * m0;  // v3 of E1: adds e3 to g3, 0.5 e3 uy to g6, 0.5 e3 ux to g7, 0.5 e3 (ux2+uy2) to g8
  double gs[9], gh[9], ds[9], dh[9];
#pragma unroll
  for (int k = 0; k < 9; k++)
  {
    gs[k] = gP[k] + gE[k];
    gh[k] = gH[k] + gE[k];
  }
  gs[8] += -cs4 * m0;
  gh[3] += e3;
  gh[6] += 0.5 * e3 * uy;
  gh[7] += 0.5 * e3 * ux;
  gh[8] += 0.5 * e3 * (ux2 + uy2);
  kbc_minv(gs, ds);
  kbc_minv(gh, dh);
  {
    // entries 5..8 as the reference writes them: (.. + ux2 + uy ..) / (.. - ux2 + uy ..) where the expansion has +- ux2*uy
    const double pp = ux2 * uy;
    const double e56 = -0.25 * m0 * ((ux2 + uy) - pp), e78 = -0.25 * m0 * ((uy - ux2) + pp);
    dh[5] += e56;
    dh[6] += e56;
    dh[7] += e78;
    dh[8] += e78;
  }
  // ---- gamma (src/ulbm.cpp:138-148).  1 / f_eq,q = 1 / (m0 phi_x(c_qx) phi_y(c_qy)) weighs both sums of the quotient, so m0
  // and the common factor 1 / (prod_a phi_x(a) prod_b phi_y(b)) cancel: the weight of q becomes the product of the OTHER
  // two phi_x times the OTHER two phi_y, and the six reciprocals (an fp64 reciprocal is a dozen instructions) are gone.
  const double ax = cs2 + ux2, ay = cs2 + uy2;
  const double fx0 = 1.0 - ax, fx1 = 0.5 * (ax + ux), fx2 = 0.5 * (ax - ux);  // phi_x of c_x = 0, +1, -1
  const double fy0 = 1.0 - ay, fy1 = 0.5 * (ay + uy), fy2 = 0.5 * (ay - uy);
  const double rx[3] = {fx1 * fx2, fx0 * fx2, fx0 * fx1};
  const double ry[3] = {fy1 * fy2, fy0 * fy2, fy0 * fy1};
  double num = 0.0, den = 0.0;
#pragma unroll
  for (int q = 0; q < 9; q++)
  {
    const double ie = rx[CX(q) == 0 ? 0 : (CX(q) > 0 ? 1 : 2)] * ry[CY(q) == 0 ? 0 : (CY(q) > 0 ? 1 : 2)];
    const double w = dh[q] * ie;
    num += ds[q] * w;
    den += dh[q] * w;
  }
  const double gamma = is2 - (1.0 - is2) * num / den;
  const double gs2 = gamma * s2;
  // ---- collision: S (cT - cT_eq) through N^-1 and M^-1 (steps 1-5)
  double g[9], c[9];
#pragma unroll
  for (int k = 0; k < 9; k++) g[k] = s2 * gP[k] + gs2 * gH[k];
  if (given)
  {
    // cT0 - m0, cT1, cT2 (S = 1 on them) are not rounding noise when the caller's m0, u differ from the populations' own
    const double r0 = r00 - m0, r1 = k10, r2 = k01;
    g[0] += r0;
    g[1] += r0 * ux + r1;
    g[2] += r0 * uy + r2;
    g[3] += (r0 * (ux2 + uy2) + 2.0 * r1 * ux) + 2.0 * r2 * uy;
    g[4] += (r0 * (ux2 - uy2) + 2.0 * r1 * ux) - 2.0 * r2 * uy;
    g[5] += (r0 * uxy + r1 * uy) + r2 * ux;
    g[6] += (r0 * ux2 * uy + 2.0 * r1 * uxy) + r2 * ux2;
    g[7] += (r0 * ux * uy2 + r1 * uy2) + 2.0 * r2 * uxy;
    g[8] += (r0 * ux2 * uy2 + 2.0 * r1 * ux * uy2) + 2.0 * r2 * ux2 * uy;
  }
  kbc_minv(g, c);
#pragma unroll
  for (int q = 0; q < 9; q++) f[q] = f[q] - c[q];
}

// Anti-bounce-back constant (test/cylinder_test.cpp:135): (2 + 9 (c.uw)^2 - 3 uw.uw) w
__host__ __device__ __forceinline__ double abb_term(int q, double uwx, double uwy)
{
  const double cu = (double)CX(q) * uwx + (double)CY(q) * uwy;
  return (2.0 + 9.0 * (cu * cu) - 3.0 * (uwx * uwx + uwy * uwy)) * W(q);
}

// One BGK collision in registers.  In: post-stream f.  Out: post-collision f, and the reference's
// `rho` / `u` variables of this iteration (u includes the += Fg shift of the gravity driver).
//   FORCE_NONE    solver::collision                      (src/solver.cpp:65-74)
//   FORCE_UNIFORM test/gravity_test.cpp:139-160          (ics2 = 1/3, ics4 = 1/9 as named there)
//   FORCE_IBM     test/cylinder_test.cpp:100-127         (source only where the ROI force is given)
//   EQ_KBC        ulbm::d2q9::kbc::collide               (src/ulbm.cpp:91-126), omega = s2; `given` = the
//                 caller's m0, u for the first step after an import (passed in through rho, ux, uy)
template <int EQ, int FORCE>
__host__ __device__ __forceinline__ void bgk_collide(double (&f)[9], const BgkParams& p, bool in_roi, double Fx, double Fy,
                                            double& rho, double& ux, double& uy, bool given = false)
{
  double jx, jy;
  if constexpr (EQ == EQ_KBC)
  {
    kbc_collide(f, p.omega, p.inv_omega, rho, ux, uy, given);  // given = false: rho, ux, uy come back as the populations' moments
    return;
  }
  moments(f, rho, jx, jy);
  if constexpr (EQ == EQ_COMP)
  {
    ux = jx / rho;
    uy = jy / rho;
  }
  else
  {
    ux = jx;
    uy = jy;
  }
  if constexpr (FORCE == FORCE_UNIFORM)
  {
    ux += p.Fg0;
    uy += p.Fg1;
  }
  const double uu = ux * ux + uy * uy;
  const double omega = p.omega;
  if constexpr (FORCE == FORCE_NONE)
  {
#pragma unroll
    for (int q = 0; q < 9; q++) f[q] = (1.0 - omega) * f[q] + omega * feq_any<EQ>(q, rho, ux, uy, uu);
  }
  else
  {
    // compile-time constants on the drivers' standard path: as run-time parameters they cost the cylinder workload 3 %
    const double ics2 = FORCE == FORCE_REGION ? p.ics2 : 1.0 / 3.0, ics4 = FORCE == FORCE_REGION ? p.ics4 : 1.0 / 9.0;
    double fx = Fx, fy = Fy;
    if constexpr (FORCE == FORCE_UNIFORM)
    {
      fx = p.Fg0;
      fy = p.Fg1;
    }
    const double uF = ux * fx + uy * fy;
    const double pref = 1.0 - 0.5 * omega;
#pragma unroll
    for (int q = 0; q < 9; q++)
    {
      const double fe = feq_any<EQ>(q, rho, ux, uy, uu);
      double v = f[q] + (-omega * (f[q] - fe));
      if (FORCE == FORCE_UNIFORM || in_roi)
      {
        const double cu = (double)CX(q) * ux + (double)CY(q) * uy;
        const double cF = (double)CX(q) * fx + (double)CY(q) * fy;
        v += (pref * ((ics2 + ics4 * cu) * cF - ics2 * uF)) * W(q);
      }
      f[q] = v;
    }
  }
}

// Second lattice of the sedimentation driver (test/rectangle_sedimentation_test.cpp:125,131):
// g_coll = (1-w) g + w * equilibrium(u + w_s, C), with the scalar w_s added to both components.
__host__ __device__ __forceinline__ void ade_collide(double (&g)[9], double omega_g, double ux, double uy, double w_s, double& C)
{
  double jx, jy;
  moments(g, C, jx, jy);
  const double ax = ux + w_s, ay = uy + w_s;
  const double aa = ax * ax + ay * ay;
#pragma unroll
  for (int q = 0; q < 9; q++) g[q] = (1.0 - omega_g) * g[q] + omega_g * feq_comp(q, C, ax, ay, aa);
}

}  // namespace lbm
