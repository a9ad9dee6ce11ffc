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

enum EqKind { EQ_COMP = 0, EQ_INCOMP = 1 };
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

template <int EQ>
__device__ __forceinline__ double feq_any(int q, double rho, double ux, double uy, double uu)
{
  if constexpr (EQ == EQ_COMP) return feq_comp(q, rho, ux, uy, uu);
  else return feq_incomp(q, rho, ux, uy);
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
template <int EQ, int FORCE>
__device__ __forceinline__ void bgk_collide(double (&f)[9], const BgkParams& p, bool in_roi, double Fx, double Fy,
                                            double& rho, double& ux, double& uy)
{
  double jx, jy;
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
