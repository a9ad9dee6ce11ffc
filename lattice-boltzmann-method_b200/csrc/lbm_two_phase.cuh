// lbm_two_phase.cuh — per-node arithmetic of the two colour-gradient models.
//
//   MODEL_MRTCG : test/mrtcg_rayleigh_taylor.cpp:431-464 (and mrtcg_static_droplet.cpp, same operators)
//   MODEL_RK    : test/rk_static_droplet_test.cpp:157-262,544-609
//
// Two lattices (red = 0, blue = 1) in the same SoA layout as the single-phase family, plus five
// scalar moment planes (rho_r, rho_b, u_x, u_y, phase) padded by two replicated cells on every
// side — the padding of the reference's replicate-mode Conv2d (src/differential.cpp:8-9) — so the
// 5x5 / 3x3 finite differences never clamp an index.
#pragma once
#include "../../include/lbm_b200.h"
#include "lbm_device.cuh"

namespace lbm
{

// TP_CSF = test/mrt_rayleigh_taylor.cpp: the MRTCG model with the perturbation replaced by a continuum-surface-force
// term (curvature from a second pass of the 5x5 differences over the interface normal) and the interfacial
// tension carried into the next step's velocity
enum TpModel { TP_MRTCG = 0, TP_RK = 1, TP_CSF = 2 };

struct TpParams
{
  double r_rho0, b_rho0, r_beta, b_beta, r_A, b_A;
  double r_irho0, b_irho0;    // 1 / rho0_k
  double w2sum;               // TP_CSF: A_r (1 - rlx_r / 2) + A_b (1 - rlx_b / 2)   (mrt_rayleigh_taylor.cpp:512-513)
  double r_phi[3], b_phi[3];  // by |c|^2 class: q = 0, q = 1..4, q = 5..8  (src/colour.cpp:56-64)
  double r_eta[3], b_eta[3];  // src/colour.cpp:49-54
  double cr, cb;              // (1.8 alpha_k - 0.8)   (mrtcg_rayleigh_taylor.cpp:328-329)
  double sigma, Fg0, Fg1;
  int add_force;
  // relaxation blend (mrtcg_rayleigh_taylor.cpp:34-100 in omega; rk_static_droplet_test.cpp:287-359 in tau)
  double delta, r_val, b_val, s1, s2, s3, t2, t3;
};

// Per-fluid constants of class colour (src/colour.cpp:37,49-64) by |c|^2 class
inline void tp_fill_colour(const lbm_colour_desc& c, double (&phi)[3], double (&eta)[3], double& cs2)
{
  cs2 = 3.0 * (1.0 - c.alpha) / 5.0;  // src/colour.cpp:37
  phi[0] = c.alpha;
  phi[1] = 0.2 * (1.0 - c.alpha);
  phi[2] = 0.05 * (1.0 - c.alpha);
  for (int k = 0; k < 3; k++) eta[k] = 1.0 + 0.5 * (3.0 * cs2 - 1.0) * (3.0 * (double)k - 4.0);
}

// Everything the kernels need from an lbm_config, in host fp64 with the drivers' formula order.  A header function so
// that the CPU test of the collision arithmetic (tests/host_kernels) derives its constants with the product's own code.
inline void tp_fill_params(const lbm_config& c, TpModel model, TpParams& p)
{
  double r_cs2, b_cs2;
  tp_fill_colour(c.red, p.r_phi, p.r_eta, r_cs2);
  tp_fill_colour(c.blue, p.b_phi, p.b_eta, b_cs2);
  p.r_rho0 = c.red.rho_0; p.b_rho0 = c.blue.rho_0;
  p.r_irho0 = 1.0 / c.red.rho_0; p.b_irho0 = 1.0 / c.blue.rho_0;
  p.r_beta = c.red.beta; p.b_beta = c.blue.beta;
  p.r_A = c.red.A; p.b_A = c.blue.A;
  p.cr = 1.8 * c.red.alpha - 0.8;
  p.cb = 1.8 * c.blue.alpha - 0.8;
  p.sigma = c.sigma;
  p.Fg0 = c.Fg[0]; p.Fg1 = c.Fg[1];
  p.add_force = c.add_force;
  p.delta = c.delta;
  p.w2sum = 0.0;
  if (model != TP_RK)
  {
    // relaxation_function{r, b, delta}: omegas from nu and the colour's own cs2 (mrtcg_rayleigh_taylor.cpp:57-66)
    p.r_val = 1.0 / (0.5 + c.red.nu / r_cs2);
    p.b_val = 1.0 / (0.5 + c.blue.nu / b_cs2);
    // TP_CSF: omega2_k = A_k (1 - rlx_k / 2) eta with colour::rlx = 1 / (0.5 + nu / cs2) (src/colour.cpp:38-39)
    p.w2sum = c.red.A * (1.0 - 0.5 * p.r_val) + c.blue.A * (1.0 - 0.5 * p.b_val);
  }
  else
  {
    // colour::init_omega with cs2 = 1/3, blended in tau space (rk_static_droplet_test.cpp:264-265,320-323)
    const double cs2 = 1.0 / 3.0;
    const double r_om = 1.0 / (0.5 + c.red.nu / cs2), b_om = 1.0 / (0.5 + c.blue.nu / cs2);
    p.r_val = 1.0 / r_om;
    p.b_val = 1.0 / b_om;
  }
  if (model == TP_CSF) p.add_force = 1;  // mrt_rayleigh_taylor.cpp:527-531
  p.s1 = 2.0 * p.r_val * p.b_val / (p.r_val + p.b_val);
  p.s2 = 2.0 * (p.r_val - p.s1) / p.delta;
  p.s3 = -p.s2 / (2.0 * p.delta);
  p.t2 = 2.0 * (p.s1 - p.b_val) / p.delta;
  p.t3 = p.t2 / (2.0 * p.delta);
}

struct MomGeom
{
  int pm;            // pitch of a moment plane (>= Y + 4, multiple of 16)
  long long mplane;  // (Xl + 4) * pm
};

enum MomPlane { M_RR = 0, M_RB = 1, M_UX = 2, M_UY = 3, M_PH = 4, M_COUNT = 5 };

__device__ __forceinline__ long long mom_off(const MomGeom& m, int x, int y) { return (long long)(x + 2) * m.pm + (y + 2); }

__host__ __device__ __forceinline__ constexpr int QCLASS(int q) { return q == 0 ? 0 : (q < 5 ? 1 : 2); }

// B of the perturbation operator (mrtcg_rayleigh_taylor.cpp:158-163)
__host__ __device__ __forceinline__ constexpr double BQ(int q) { return q == 0 ? -4.0 / 27.0 : (q < 5 ? 2.0 / 27.0 : 5.0 / 108.0); }

// 5x5 isotropic weights xi[|a|][|b|] / 5040 (src/differential.hpp:9-18)
__host__ __device__ __forceinline__ constexpr double XI5(int a, int b)
{
  const int aa = a < 0 ? -a : a, bb = b < 0 ? -b : b;
  return (aa == 0 ? (bb == 0 ? 0.0 : (bb == 1 ? 960.0 : 84.0))
                  : (aa == 1 ? (bb == 0 ? 960.0 : (bb == 1 ? 448.0 : 32.0)) : (bb == 0 ? 84.0 : (bb == 1 ? 32.0 : 1.0)))) *
         (1.0 / 5040.0);
}

// M (Lallemand-Luo ordering rho,e,eps,jx,qx,jy,qy,pxx,pxy) and 36 M^-1 (mrtcg_rayleigh_taylor.cpp:130-156)
__host__ __device__ __forceinline__ constexpr double MM(int a, int q)
{
  switch (a)
  {
    case 0: return 1.0;
    case 1: return q == 0 ? -4.0 : (q < 5 ? -1.0 : 2.0);
    case 2: return q == 0 ? 4.0 : (q < 5 ? -2.0 : 1.0);
    case 3: return (double)CX(q);
    case 4: return q < 5 ? -2.0 * (double)CX(q) : (double)CX(q);
    case 5: return (double)CY(q);
    case 6: return q < 5 ? -2.0 * (double)CY(q) : (double)CY(q);
    case 7: return (q == 1 || q == 3) ? 1.0 : ((q == 2 || q == 4) ? -1.0 : 0.0);
    default: return (q == 5 || q == 7) ? 1.0 : ((q == 6 || q == 8) ? -1.0 : 0.0);
  }
}
__host__ __device__ __forceinline__ constexpr double MI36(int q, int a)
{
  switch (a)
  {
    case 0: return 4.0;
    case 1: return q == 0 ? -4.0 : (q < 5 ? -1.0 : 2.0);
    case 2: return q == 0 ? 4.0 : (q < 5 ? -2.0 : 1.0);
    case 3: return 6.0 * (double)CX(q);
    case 4: return q < 5 ? -6.0 * (double)CX(q) : 3.0 * (double)CX(q);
    case 5: return 6.0 * (double)CY(q);
    case 6: return q < 5 ? -6.0 * (double)CY(q) : 3.0 * (double)CY(q);
    case 7: return (q == 1 || q == 3) ? 9.0 : ((q == 2 || q == 4) ? -9.0 : 0.0);
    default: return (q == 5 || q == 7) ? 9.0 : ((q == 6 || q == 8) ? -9.0 : 0.0);
  }
}

// relaxation_function::eval; a NaN phase matches no branch (the reference then keeps the previous
// value, zero on the first step)
__host__ __device__ __forceinline__ double relax_eval(const TpParams& p, double psi)
{
  double s = 0.0;
  if (psi > p.delta) s = p.r_val;
  if (p.delta >= psi && psi > 0.0) s = p.s1 + p.s2 * psi + p.s3 * psi * psi;
  if (0.0 >= psi && psi >= -p.delta) s = p.s1 + p.t2 * psi + p.t3 * psi * psi;
  if (psi < -p.delta) s = p.b_val;
  return s;
}

// Divisions: fp64 division costs ~12 issue slots on the fp64 pipe, and the reference's formulas hold
// ~45 of them per node.  Every quotient by a per-node scalar (rho, |grad|, rho0_k) is therefore taken as
// a product with ONE reciprocal of that scalar; the result differs from the reference's quotient by
// <= 1 ulp per operation, far inside the parity tolerance (1e-12 relative after one step).

// eval_phase_field (mrtcg_rayleigh_taylor.cpp:212-225)
__host__ __device__ __forceinline__ double phase_of(const TpParams& p, double rr, double rb)
{
  // The two products are rounded on their own (no contraction into a - b / a + b): in the bulk of one fluid the other
  // density is ~1e-22 of rounding residue, and the phase must come out as EXACTLY +-1 there, as it does in the
  // reference — the models divide by 1e-20 + |grad(phase)|, so a 1-ulp ripple on the plateau would turn into an O(1)
  // interface "normal".  (Observed: fma(3, RN(1/3), -+1e-22) straddles a rounding midpoint and gave 1 + 2^-52.)
#ifdef __CUDA_ARCH__
  const double a = __dmul_rn(rr, p.r_irho0), b = __dmul_rn(rb, p.b_irho0);
  return __dsub_rn(a, b) / __dadd_rn(a, b);
#else  // host build (tests/host_kernels): nothing contracts there
  const double a = rr * p.r_irho0, b = rb * p.b_irho0;
  return (a - b) / (a + b);
#endif
}

// Moments of a freshly streamed node: what the drivers compute at the END of an iteration
// (mrtcg_rayleigh_taylor.cpp:472-477 ; rk_static_droplet_test.cpp:602-609).
// fsx, fsy: TP_CSF only, the interfacial tension of the step that produced this state (mrt_rayleigh_taylor.cpp:543-544)
template <int MODEL>
__host__ __device__ __forceinline__ void tp_moments(const TpParams& p, const double (&fr)[9], const double (&fb)[9], double& rr,
                                           double& rb, double& ux, double& uy, double& ph, double fsx = 0.0, double fsy = 0.0)
{
  double jx, jy, dummy;
  moments(fr, rr, jx, jy);
  moments(fb, rb, jx, jy);
  const double rho = rr + rb;
  double t[9];
#pragma unroll
  for (int q = 0; q < 9; q++) t[q] = fr[q] + fb[q];
  moments(t, dummy, jx, jy);
  const double inv_rho = 1.0 / rho;
  ux = jx * inv_rho;
  uy = jy * inv_rho;
  if constexpr (MODEL == TP_MRTCG)
  {
    ux = ux + (0.5 * p.Fg0) * inv_rho;
    uy = uy + (0.5 * p.Fg1) * inv_rho;
  }
  if constexpr (MODEL == TP_CSF)
  {
    ux = ux + (0.5 * (p.Fg0 + fsx)) * inv_rho;
    uy = uy + (0.5 * (p.Fg1 + fsy)) * inv_rho;
  }
  ph = phase_of(p, rr, rb);
}

// eval_equilibrium of one colour
template <int MODEL>
__host__ __device__ __forceinline__ double tp_feq(int q, double rho_k, const double (&phi)[3], const double (&eta)[3], double ux,
                                         double uy, double uu)
{
  const double ue = (double)CX(q) * ux + (double)CY(q) * uy;
  if constexpr (MODEL != TP_RK)  // mrtcg_rayleigh_taylor.cpp:233-247 (note 9 and 3, not 4.5 and 1.5)
    return rho_k * (phi[QCLASS(q)] + W(q) * ((3.0 * ue) * eta[QCLASS(q)] + 9.0 * (ue * ue) - 3.0 * uu));
  else  // rk_static_droplet_test.cpp:183-199
    return rho_k * (phi[QCLASS(q)] + ((3.0 * ue + 4.5 * (ue * ue)) - 1.5 * uu) * W(q));
}

// Stencil results a collision needs
struct TpStencil
{
  double gx, gy;               // grad(phase): MRTCG axis-0 / axis-1 derivatives; RK the driver's swapped pair
  // MRTCG only: d/dx of Q_x and d/dy of Q_y, Q = Q_r + Q_b = (c_r rho_r + c_b rho_b) u.  update_C (:320-336) wants
  // these per colour, but only the colour SUM of C enters the update (see tp_collide) and the differences are
  // linear, so the two colours' momentum fields are added before they are differentiated.
  double DxQx, DyQy;
  double Fsx, Fsy;  // TP_CSF only: interfacial tension -sigma/2 K grad(phase) of this step (input of the collision)
};

// One two-colour collision in registers: fr/fb in = post-stream, out = post-collision.
template <int MODEL>
__host__ __device__ __forceinline__ void tp_collide(const TpParams& p, double (&fr)[9], double (&fb)[9], double rr, double rb,
                                           double ux, double uy, double ph, const TpStencil& st)
{
  const double uu = ux * ux + uy * uy;
  const double gn = sqrt(st.gx * st.gx + st.gy * st.gy);
  if constexpr (MODEL == TP_RK)
  {
    // relax = 1/tau(phase) (:587-588); omega1 = relax (feq - f) (:255-262); omega2 Reis (:239-245); col = f + (w1 + w2)
    const double relax = 1.0 / relax_eval(p, ph);
    const double gn2 = gn * gn;
    const double ign2 = 1.0 / (1e-20 + gn2);
#pragma unroll
    for (int q = 0; q < 9; q++)
    {
      const double fe = st.gx * (double)CX(q) + st.gy * (double)CY(q);
      const double core = ((fe * fe) * ign2) * W(q) - BQ(q);
      const double o1r = relax * (tp_feq<TP_RK>(q, rr, p.r_phi, p.r_eta, ux, uy, uu) - fr[q]);
      const double o1b = relax * (tp_feq<TP_RK>(q, rb, p.b_phi, p.b_eta, ux, uy, uu) - fb[q]);
      fr[q] = fr[q] + (o1r + ((0.5 * p.r_A) * gn) * core);
      fb[q] = fb[q] + (o1b + ((0.5 * p.b_A) * gn) * core);
    }
  }
  else
  {
    // mrtcg_rayleigh_taylor.cpp:431-464.  Two restructurings, both exact up to rounding:
    //  (1) only the SUM of the two colours' MRT operators enters the update (total = sum_k (f_k + omega1_k
    //      + omega2_k), :455, then recoloured, :456-457) and M^-1 S M is linear, so the transform runs once on
    //      d = (feq_r + feq_b) - (f_r + f_b) with C = C_r + C_b instead of once per colour;
    //  (2) D2Q9 directions come in opposite pairs (1,3) (2,4) (5,7) (6,8): every per-direction term is
    //      even (u.c)^2-like or odd (u.c)-like in c, so each pair shares one even and one odd part, and the
    //      M / M^-1 rows act on pair sums and differences.
    const double rho = rr + rb;
    const double s_nu = relax_eval(p, ph);
    const double mix[3] = {rr * p.r_phi[0] + rb * p.b_phi[0], rr * p.r_phi[1] + rb * p.b_phi[1], rr * p.r_phi[2] + rb * p.b_phi[2]};
    const double emix1 = rr * p.r_eta[1] + rb * p.b_eta[1], emix2 = rr * p.r_eta[2] + rb * p.b_eta[2];
    const double rinv = 1.0 / (1e-20 + gn);
    const double inv_rho = 1.0 / rho;
    const double wr = rr * inv_rho, wb = rb * inv_rho;            // rho_k / rho
    const double k0 = (wr * wb) * rinv;                           // rho_r rho_b / (rho^2 (1e-20 + |grad|))
    const double A2 = (4.5 * p.sigma * s_nu) * (0.5 * gn) * 2.0;  // both colours: A = 4.5 sigma s_nu, xi = 0.5 |grad| (..)
    const double uF = ux * p.Fg0 + uy * p.Fg1;
    // TP_CSF: omega2_k = A_k (1 - rlx_k / 2) eta, eta_q = w_q (3 (c_q - u) + 9 (u.c_q) c_q) . Fs   (eval_eta :366-385)
    const double uFs = ux * st.Fsx + uy * st.Fsy;
    const double Fse[4] = {st.Fsx, st.Fsy, st.Fsx + st.Fsy, st.Fsy - st.Fsx};
    const double pref = 1.0 - 0.5 * s_nu;
    constexpr double ISQ2 = 0.7071067811865476;
    constexpr double W1 = 1.0 / 9.0, W2 = 1.0 / 36.0;
    // c_q . v for the pair representatives q = 1, 2, 5, 6 (the opposite direction has the opposite sign)
    const double ue[4] = {ux, uy, ux + uy, uy - ux};
    const double ge[4] = {st.gx, st.gy, st.gx + st.gy, st.gy - st.gx};
    const double Fe[4] = {p.Fg0, p.Fg1, p.Fg0 + p.Fg1, p.Fg1 - p.Fg0};
    constexpr int QA[4] = {1, 2, 5, 6}, QB[4] = {3, 4, 7, 8};

    // ---- d = feq_sum - f_sum, as pair sums s and differences a  (eval_equilibrium :233-247, note 9 and 3)
    const double fs0 = fr[0] + fb[0];
    const double d0 = (mix[0] + (4.0 / 9.0) * (rho * (-3.0 * uu))) - fs0;
    double fsA[4], fsB[4], s[4], a[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
      const double Wc = k < 2 ? W1 : W2, mx = k < 2 ? mix[1] : mix[2], em = k < 2 ? emix1 : emix2;
      const double even = mx + Wc * (rho * (9.0 * (ue[k] * ue[k]) - 3.0 * uu));
      const double odd = Wc * ((3.0 * ue[k]) * em);
      fsA[k] = fr[QA[k]] + fb[QA[k]];
      fsB[k] = fr[QB[k]] + fb[QB[k]];
      const double dA = (even + odd) - fsA[k], dB = (even - odd) - fsB[k];
      s[k] = dA + dB;
      a[k] = dA - dB;
    }
    // ---- m = S M d + C on the six non-conserved moments (M :130-142; S = diag(0,1.25,1.14,0,1.6,0,1.6,s_nu,s_nu)
    //      :384-387 + update_S; C: update_C :320-336, both colours)
    const double SA = s[0] + s[1], SD = s[2] + s[3];
    const double DQ1 = st.DxQx + st.DyQy;
    const double DQ7 = st.DxQx - st.DyQy;
    const double m_e = 1.25 * ((2.0 * SD - SA) - 4.0 * d0) + (3.0 * (1.0 - 0.5 * 1.25)) * DQ1;
    const double m_eps = 1.14 * ((SD - 2.0 * SA) + 4.0 * d0);
    const double m_qx = 1.6 * ((a[2] - a[3]) - 2.0 * a[0]);
    const double m_qy = 1.6 * ((a[2] + a[3]) - 2.0 * a[1]);
    const double m_pxx = s_nu * (s[0] - s[1]) + pref * DQ7;
    const double m_pxy = s_nu * (s[2] - s[3]);
    // ---- omega1 = M^-1 m (36 M^-1 :144-156): even and odd part per pair
    const double o1_0 = (1.0 / 9.0) * (m_eps - m_e);
    const double E_ax = (-1.0 / 36.0) * (m_e + 2.0 * m_eps), E_dg = (1.0 / 36.0) * (2.0 * m_e + m_eps);
    const double o1E[4] = {E_ax + 0.25 * m_pxx, E_ax - 0.25 * m_pxx, E_dg + 0.25 * m_pxy, E_dg - 0.25 * m_pxy};
    const double o1O[4] = {(-1.0 / 6.0) * m_qx, (-1.0 / 6.0) * m_qy, (1.0 / 12.0) * (m_qx + m_qy), (1.0 / 12.0) * (m_qy - m_qx)};

    // ---- perturbation, recolouring, force (eval_xi / eval_per_operator / eval_kappa / eval_rec_operator, :455-464)
    {
      // q = 0: grad . c = 0, B_0 = -4/27  |  CSF: eta_0 = w_0 (-3 u.Fs)
      const double total = (fs0 + o1_0) + (MODEL == TP_CSF ? p.w2sum * ((4.0 / 9.0) * (-3.0 * uFs)) : A2 * (4.0 / 27.0));
      double nr = wr * total, nb = wb * total;
      if (p.add_force)
      {
        const double src = (pref * (-3.0 * uF)) * (4.0 / 9.0);
        nr += src;
        nb += src;
      }
      fr[0] = nr;
      fb[0] = nb;
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
      const double Wc = k < 2 ? W1 : W2, Bc = k < 2 ? 2.0 / 27.0 : 5.0 / 108.0, mx = k < 2 ? mix[1] : mix[2];
      double o2e, o2o, kap;  // omega2 summed over the colours: part even in c, part odd in c; kappa is odd in c
      if constexpr (MODEL == TP_CSF)
      {
        o2e = (p.w2sum * Wc) * (9.0 * (ue[k] * Fse[k]) - 3.0 * uFs);
        o2o = (p.w2sum * Wc) * (3.0 * Fse[k]);
        kap = (k0 * ge[k]) * mx;  // eval_kappa with E, not unit_E (mrt_rayleigh_taylor.cpp:317)
      }
      else
      {
        const double t = ge[k] * rinv;
        o2e = A2 * (Wc * (t * t) - Bc);
        o2o = 0.0;
        kap = (k0 * (k < 2 ? ge[k] : ge[k] * ISQ2)) * mx;
      }
      const double totA = (fsA[k] + (o1E[k] + o1O[k])) + (o2e + o2o);
      const double totB = (fsB[k] + (o1E[k] - o1O[k])) + (o2e - o2o);
      const double kr = p.r_beta * kap, kb = p.b_beta * kap;
      double nrA = wr * totA + kr, nbA = wb * totA + kb;
      double nrB = wr * totB - kr, nbB = wb * totB - kb;
      if (p.add_force)  // (1 - s_nu/2) ((3 + 9 u.c) F.c - 3 u.F) w   (ics2 = 3, ics4 = 9)
      {
        const double ev = (pref * Wc) * (9.0 * (ue[k] * Fe[k]) - 3.0 * uF), od = (pref * Wc) * (3.0 * Fe[k]);
        nrA += ev + od;
        nbA += ev + od;
        nrB += ev - od;
        nbB += ev - od;
      }
      fr[QA[k]] = nrA;
      fb[QA[k]] = nbA;
      fr[QB[k]] = nrB;
      fb[QB[k]] = nbB;
    }
  }
}

}  // namespace lbm
