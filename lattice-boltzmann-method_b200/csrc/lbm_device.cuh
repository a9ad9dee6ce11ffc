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
enum ForceKind { FORCE_NONE = 0, FORCE_UNIFORM = 1, FORCE_IBM = 2 };

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
  double omega_g;  // ADE lattice relaxation
  double Fg0, Fg1; // uniform force (test/gravity_test.cpp:85)
  double w_s;      // settling velocity (test/rectangle_sedimentation_test.cpp:89)
  // immersed-boundary force field on the ROI (test/cylinder_test.cpp:110-127), global coordinates
  int roi_r0, roi_r1, roi_c0, roi_c1;
  const double* Fx;
  const double* Fy;
  // LBM_MODEL_KBC, first step after an import: rho {Xl,Y}, u {Xl,Y,2} supplied by the caller (the drivers' m0 / m1
  // members, test/ulbm_poiseuille.cpp:93) instead of the moments of the imported populations; nullptr otherwise
  const double* mom_in_rho;
  const double* mom_in_u;
  int Y;           // row length of mom_in_*
};

// rho = sum_q f, (jx, jy) = sum_q f c_q in the q order of the reference's reductions
// (solver::calc_rho / calc_incomp_u, src/solver.cpp:23-31).
__device__ __forceinline__ void moments(const double (&f)[9], double& rho, double& jx, double& jy)
{
  rho = ((((((((f[0] + f[1]) + f[2]) + f[3]) + f[4]) + f[5]) + f[6]) + f[7]) + f[8]);
  jx = (((((f[1] - f[3]) + f[5]) - f[6]) - f[7]) + f[8]);
  jy = (((((f[2] - f[4]) + f[5]) + f[6]) - f[7]) - f[8]);
}

// solver::equilibrium (src/solver.cpp:51-62)
__device__ __forceinline__ double feq_comp(int q, double rho, double ux, double uy, double uu)
{
  const double cu = (double)CX(q) * ux + (double)CY(q) * uy;
  const double A = 1.0 + 3.0 * cu + 4.5 * (cu * cu) - 1.5 * uu;
  return (rho * A) * W(q);
}

// solver::incomp_equilibrium (src/solver.cpp:39-49)
__device__ __forceinline__ double feq_incomp(int q, double rho, double ux, double uy)
{
  const double cu = (double)CX(q) * ux + (double)CY(q) * uy;
  return (rho + 3.0 * cu) * W(q);
}

__device__ __forceinline__ double feq_kbc_q(int q, double rho, double ux, double uy);

template <int EQ>
__device__ __forceinline__ double feq_any(int q, double rho, double ux, double uy, double uu)
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
__device__ __forceinline__ double feq_kbc_q(int q, double rho, double ux, double uy)
{
  double e[9];
  kbc_eq_coef(ux, uy, ux * ux, uy * uy, e);
  double r = 0.0;
#pragma unroll
  for (int k = 0; k < 9; k++)
    if (k == q) r = 1.0 / (1.0 / (e[k] * rho));
  return r;
}

// kbc::collide() of one node (src/ulbm.cpp:91-126: eval_central_momenta :264-320, eval_gamma :138-148,
// eval_delta_s :157-189, eval_delta_h :191-224 with its `ux2+uy` terms as written, eval_iequilibrium :226-244,
// then S (cT - cT_eq), N^-1, -M^-1, + adve_f).  In: post-stream f and the iteration's m0, u.  Out: post-collision f.
__device__ __forceinline__ void kbc_collide(double (&f)[9], double s2, double m0, double ux, double uy)
{
  constexpr double cs2 = 1.0 / 3.0, cs4 = 1.0 / 9.0;
  const double is2 = 1.0 / s2;
  double cT[9];
#pragma unroll
  for (int k = 0; k < 9; k++) cT[k] = 0.0;
#pragma unroll
  for (int q = 0; q < 9; q++)
  {
    const double cmx = CX(q) == 0 ? -ux : (CX(q) > 0 ? 1.0 - ux : -1.0 - ux);
    const double cmy = CY(q) == 0 ? -uy : (CY(q) > 0 ? 1.0 - uy : -1.0 - uy);
    const double cmx2 = cmx * cmx, cmy2 = cmy * cmy;
    cT[0] += f[q];
    cT[1] += f[q] * cmx;
    cT[2] += f[q] * cmy;
    cT[3] += f[q] * (cmx2 + cmy2);
    cT[4] += f[q] * (cmx2 - cmy2);
    cT[5] += f[q] * cmx * cmy;
    cT[6] += f[q] * cmx2 * cmy;
    cT[7] += f[q] * cmx * cmy2;
    cT[8] += f[q] * cmx2 * cmy2;
  }
  const double ux2 = ux * ux, uy2 = uy * uy;
  const double C3 = cT[3], C4 = cT[4], C5 = cT[5], C6 = cT[6], C7 = cT[7], C8 = cT[8];
  const double K3 = C3 - 2.0 * cs2 * m0;
  double ds[9], dh[9], e[9];
  ds[0] = -0.5 * C4 * (ux2 - uy2) + 4.0 * C5 * ux * uy - cs4 * m0 - m0 * (ux2 * uy2 - ux2 - uy2 + 1) + K3 * (0.5 * ux2 + 0.5 * uy2 - 1.0);
  ds[1] = 0.25 * C4 * (ux2 - uy2 + ux + 1) - C5 * uy * (2.0 * ux + 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 + uy2 * ux - ux) - 0.25 * K3 * (ux2 + uy2 + ux - 1.0);
  ds[2] = -0.25 * C4 * (-ux2 + uy2 + uy + 1) - C5 * ux * (2.0 * uy + 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - uy2 + ux2 * uy - uy) - 0.25 * K3 * (ux2 + uy2 + uy - 1.0);
  ds[3] = 0.25 * C4 * (ux2 - uy2 - ux + 1) - C5 * uy * (2.0 * ux - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 - uy2 * ux + ux) - 0.25 * K3 * (ux2 + uy2 - ux - 1.0);
  ds[4] = 0.25 * C4 * (ux2 - uy2 + uy - 1) - C5 * ux * (2.0 * uy - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - uy2 - ux2 * uy + uy) - 0.25 * K3 * (ux2 + uy2 - uy - 1.0);
  ds[5] = -0.125 * C4 * (ux2 - uy2 + ux - uy) + C5 * (ux * uy + 0.5 * ux + 0.5 * uy + 0.25) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 + ux2 * uy + uy2 * ux + ux * uy) + 0.125 * K3 * (ux2 + uy2 + ux + uy);
  ds[6] = 0.125 * C4 * (-ux2 + uy2 + ux + uy) + C5 * (ux * uy + 0.5 * ux - 0.5 * uy - 0.25) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 + ux2 * uy - uy2 * ux - ux * uy) + 0.125 * K3 * (ux2 + uy2 - ux + uy);
  ds[7] = -0.125 * C4 * (ux2 - uy2 - ux + uy) + C5 * (ux * uy - 0.5 * ux - 0.5 * uy + 0.25) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 - ux2 * uy - uy2 * ux + ux * uy) + 0.125 * K3 * (ux2 + uy2 - ux - uy);
  ds[8] = -0.125 * C4 * (ux2 - uy2 + ux + uy) + C5 * (ux * uy - 0.5 * ux + 0.5 * uy - 0.25) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 - ux2 * uy + uy2 * ux - ux * uy) + 0.125 * K3 * (ux2 + uy2 + ux - uy);
  dh[0] = 2.0 * C6 * uy + 2.0 * C7 * ux + C8 - 2.0 * cs2 * m0 * (0.5 * ux2 + 0.5 * uy2 - 1.0) - cs4 * m0 - m0 * (ux2 * uy2 - ux2 - uy2 + 1.0);
  dh[1] = -C6 * uy - C7 * (ux + 0.5) - 0.5 * C8 + 0.5 * cs2 * m0 * (ux2 + uy2 + ux - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 + uy2 * ux - ux);
  dh[2] = -C6 * (uy + 0.5) - C7 * ux - 0.5 * C8 + 0.5 * cs2 * m0 * (ux2 + uy2 + uy - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 + ux2 * uy - uy2 - uy);
  dh[3] = -C6 * uy - C7 * (ux - 0.5) - 0.5 * C8 + 0.5 * cs2 * m0 * (ux2 + uy2 - ux - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 - uy2 * ux + ux);
  dh[4] = -C6 * (uy - 0.5) - C7 * ux - 0.5 * C8 + 0.5 * cs2 * m0 * (ux2 + uy2 - uy - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 * uy - uy2 + uy);
  // src/ulbm.cpp:211-223 as written: `ux2+uy`, not `ux2*uy`
  dh[5] = C6 * (0.5 * uy + 0.25) + C7 * (0.5 * ux + 0.25) + 0.25 * C8 - 0.25 * cs2 * m0 * (ux2 + uy2 + ux + uy) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 + ux2 + uy + uy2 * ux + ux * uy);
  dh[6] = C6 * (0.5 * uy + 0.25) + C7 * (0.5 * ux - 0.25) + 0.25 * C8 - 0.25 * cs2 * m0 * (ux2 + uy2 - ux + uy) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 + ux2 + uy - uy2 * ux - ux * uy);
  dh[7] = C6 * (0.5 * uy - 0.25) + C7 * (0.5 * ux - 0.25) + 0.25 * C8 - 0.25 * cs2 * m0 * (ux2 + uy2 - ux - uy) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 - ux2 + uy - uy2 * ux + ux * uy);
  dh[8] = C6 * (0.5 * uy - 0.25) + C7 * (0.5 * ux + 0.25) + 0.25 * C8 - 0.25 * cs2 * m0 * (ux2 + uy2 + ux - uy) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 - ux2 + uy + uy2 * ux - ux * uy);
  kbc_eq_coef(ux, uy, ux2, uy2, e);
  double num = 0.0, den = 0.0;
#pragma unroll
  for (int q = 0; q < 9; q++)
  {
    const double ie = 1.0 / (e[q] * m0);
    num += ds[q] * dh[q] * ie;
    den += dh[q] * dh[q] * ie;
  }
  const double gamma = is2 - (1.0 - is2) * num / den;
  const double gs2 = gamma * s2;
  cT[0] += -m0;
  cT[3] += -2.0 * cs2 * m0;
  cT[8] += -cs4 * m0;
  cT[3] *= s2; cT[4] *= s2; cT[5] *= s2;
  cT[6] *= gs2; cT[7] *= gs2; cT[8] *= gs2;
  double g[9];
  g[0] = cT[0];
  g[1] = cT[0] * ux + cT[1];
  g[2] = cT[0] * uy + cT[2];
  g[3] = cT[0] * (ux2 + uy2) + 2.0 * cT[1] * ux + 2.0 * cT[2] * uy + cT[3];
  g[4] = cT[0] * (ux2 - uy2) + 2.0 * cT[1] * ux - 2.0 * cT[2] * uy + cT[4];
  g[5] = cT[0] * ux * uy + cT[1] * uy + cT[2] * ux + cT[5];
  g[6] = cT[0] * ux2 * uy + 2.0 * cT[1] * ux * uy + cT[2] * ux2 + 0.5 * cT[3] * uy + 0.5 * cT[4] * uy + 2.0 * cT[5] * ux + cT[6];
  g[7] = cT[0] * ux * uy2 + cT[1] * uy2 + 2.0 * cT[2] * ux * uy + 0.5 * cT[3] * ux - 0.5 * cT[4] * ux + 2.0 * cT[5] * uy + cT[7];
  g[8] = cT[0] * ux2 * uy2 + 2.0 * cT[1] * ux * uy2 + 2.0 * cT[2] * ux2 * uy + 0.5 * cT[3] * (ux2 + uy2) - 0.5 * cT[4] * (ux2 - uy2) + 4.0 * cT[5] * ux * uy + 2.0 * cT[6] * uy + 2.0 * cT[7] * ux + cT[8];
  double c[9];
  c[0] = g[0] - g[3] + g[8];
  c[1] = 0.5 * g[1] + 0.25 * g[3] + 0.25 * g[4] - 0.5 * g[7] - 0.5 * g[8];
  c[2] = 0.5 * g[2] + 0.25 * g[3] - 0.25 * g[4] - 0.5 * g[6] - 0.5 * g[8];
  c[3] = -0.5 * g[1] + 0.25 * g[3] + 0.25 * g[4] + 0.5 * g[7] - 0.5 * g[8];
  c[4] = -0.5 * g[2] + 0.25 * g[3] - 0.25 * g[4] + 0.5 * g[6] - 0.5 * g[8];
  c[5] = 0.25 * (g[5] + g[6] + g[7] + g[8]);
  c[6] = 0.25 * (-g[5] + g[6] - g[7] + g[8]);
  c[7] = 0.25 * (g[5] - g[6] - g[7] + g[8]);
  c[8] = 0.25 * (-g[5] - g[6] + g[7] + g[8]);
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
__device__ __forceinline__ void bgk_collide(double (&f)[9], const BgkParams& p, bool in_roi, double Fx, double Fy,
                                            double& rho, double& ux, double& uy, bool given = false)
{
  double jx, jy;
  if constexpr (EQ == EQ_KBC)
  {
    if (!given)
    {
      moments(f, rho, jx, jy);
      ux = jx / rho;  // kbc.m1 = adve_f c^T / m0 (test/ulbm_double_shear_flow.cpp:145-146)
      uy = jy / rho;
    }
    kbc_collide(f, p.omega, rho, ux, uy);
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
    constexpr double ics2 = 1.0 / 3.0, ics4 = 1.0 / 9.0;
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
__device__ __forceinline__ void ade_collide(double (&g)[9], double omega_g, double ux, double uy, double w_s, double& C)
{
  double jx, jy;
  moments(g, C, jx, jy);
  const double ax = ux + w_s, ay = uy + w_s;
  const double aa = ax * ax + ay * ay;
#pragma unroll
  for (int q = 0; q < 9; q++) g[q] = (1.0 - omega_g) * g[q] + omega_g * feq_comp(q, C, ax, ay, aa);
}

}  // namespace lbm
