// TEST INFRASTRUCTURE — the product's collision arithmetic (csrc/lbm_device.cuh) compiled for the HOST, so that the CPU
// suite can hold it against the oracle without a GPU: the functions are __host__ __device__, the same source the
// kernels inline.  (Host code has no fused multiply-add: results differ from the device's in the last bits only.)
#include "../../lattice-boltzmann-method_b200/csrc/lbm_device.cuh"
#include "../../lattice-boltzmann-method_b200/csrc/lbm_two_phase.cuh"

extern "C"
{

// kbc::collide() of N nodes in place; given = 0: m0, u are the populations' own moments (computed here like
// kbc_collide does itself), given = 1: the caller's m0 {N}, u {N,2}
void host_kbc_collide(double* f, const double* m0, const double* u, int given, long N, double s2)
{
  for (long n = 0; n < N; n++)
  {
    double v[9];
    for (int q = 0; q < 9; q++) v[q] = f[9 * n + q];
    double rho = 0.0, ux = 0.0, uy = 0.0;
    if (given)
    {
      rho = m0[n]; ux = u[2 * n]; uy = u[2 * n + 1];
    }
    lbm::kbc_collide(v, s2, 1.0 / s2, rho, ux, uy, given != 0);
    for (int q = 0; q < 9; q++) f[9 * n + q] = v[q];
  }
}

// ---- single-phase family: bgk_collide<EQ, FORCE> of N nodes in place.  eq: 0 compressible, 1 incompressible;
// force: 0 none, 1 uniform (Fg added to u, gravity_test), 2 field F {N,2} with the 1/3, 1/9 constants (cylinder_test),
// 3 field F with run-time ics2, ics4 (decompose_domain_loop).  rho {N}, u {N,2}: the iteration's moments, out.
void host_bgk_collide(int eq, int force, double* f, long N, double omega, double Fg0, double Fg1, const double* F, double ics2,
                      double ics4, double* rho, double* u)
{
  lbm::BgkParams p = {};
  p.omega = omega; p.inv_omega = 1.0 / omega; p.Fg0 = Fg0; p.Fg1 = Fg1; p.ics2 = ics2; p.ics4 = ics4;
  for (long n = 0; n < N; n++)
  {
    double v[9];
    for (int q = 0; q < 9; q++) v[q] = f[9 * n + q];
    const double Fx = F ? F[2 * n] : 0.0, Fy = F ? F[2 * n + 1] : 0.0;
    double r, ux, uy;
#define HK_CASE(E, FO) if (eq == E && force == FO) lbm::bgk_collide<E, FO>(v, p, true, Fx, Fy, r, ux, uy);
    HK_CASE(lbm::EQ_COMP, lbm::FORCE_NONE) HK_CASE(lbm::EQ_COMP, lbm::FORCE_UNIFORM) HK_CASE(lbm::EQ_COMP, lbm::FORCE_IBM)
    HK_CASE(lbm::EQ_COMP, lbm::FORCE_REGION) HK_CASE(lbm::EQ_INCOMP, lbm::FORCE_NONE) HK_CASE(lbm::EQ_INCOMP, lbm::FORCE_UNIFORM)
    HK_CASE(lbm::EQ_INCOMP, lbm::FORCE_IBM) HK_CASE(lbm::EQ_INCOMP, lbm::FORCE_REGION)
#undef HK_CASE
    for (int q = 0; q < 9; q++) f[9 * n + q] = v[q];
    rho[n] = r; u[2 * n] = ux; u[2 * n + 1] = uy;
  }
}

// ---- two-phase models (csrc/lbm_two_phase.cuh).  model: 0 = MRT colour gradient, 1 = Rothman-Keller, 2 = MRT + CSF.
// Constants from the product's own tp_fill_params.

// moments of freshly streamed populations: rr, rb {N}, u {N,2}, ph {N};  Fs {N,2} (model 2) or null
void host_tp_moments(int model, const lbm_config* cfg, const double* fr, const double* fb, const double* Fs, long N,
                     double* rr, double* rb, double* u, double* ph)
{
  lbm::TpParams p;
  lbm::tp_fill_params(*cfg, (lbm::TpModel)model, p);
  for (long n = 0; n < N; n++)
  {
    double a[9], b[9];
    for (int q = 0; q < 9; q++) { a[q] = fr[9 * n + q]; b[q] = fb[9 * n + q]; }
    const double fsx = Fs ? Fs[2 * n] : 0.0, fsy = Fs ? Fs[2 * n + 1] : 0.0;
    if (model == 0) lbm::tp_moments<lbm::TP_MRTCG>(p, a, b, rr[n], rb[n], u[2 * n], u[2 * n + 1], ph[n]);
    else if (model == 1) lbm::tp_moments<lbm::TP_RK>(p, a, b, rr[n], rb[n], u[2 * n], u[2 * n + 1], ph[n]);
    else lbm::tp_moments<lbm::TP_CSF>(p, a, b, rr[n], rb[n], u[2 * n], u[2 * n + 1], ph[n], fsx, fsy);
  }
}

void host_tp_phase(int model, const lbm_config* cfg, const double* rr, const double* rb, long N, double* ph)
{
  lbm::TpParams p;
  lbm::tp_fill_params(*cfg, (lbm::TpModel)model, p);
  for (long n = 0; n < N; n++) ph[n] = lbm::phase_of(p, rr[n], rb[n]);
}

// collision of N nodes in place (fr, fb: post-stream in, post-collision out); st4 {N,4} = grad_x, grad_y of the phase
// field and d/dx Q_x, d/dy Q_y of the colour-summed momentum field; Fs {N,2} (model 2) or null
void host_tp_collide(int model, const lbm_config* cfg, double* fr, double* fb, const double* rr, const double* rb,
                     const double* u, const double* ph, const double* st4, const double* Fs, long N)
{
  lbm::TpParams p;
  lbm::tp_fill_params(*cfg, (lbm::TpModel)model, p);
  for (long n = 0; n < N; n++)
  {
    double a[9], b[9];
    for (int q = 0; q < 9; q++) { a[q] = fr[9 * n + q]; b[q] = fb[9 * n + q]; }
    lbm::TpStencil st;
    st.gx = st4[4 * n]; st.gy = st4[4 * n + 1]; st.DxQx = st4[4 * n + 2]; st.DyQy = st4[4 * n + 3];
    st.Fsx = Fs ? Fs[2 * n] : 0.0; st.Fsy = Fs ? Fs[2 * n + 1] : 0.0;
    if (model == 0) lbm::tp_collide<lbm::TP_MRTCG>(p, a, b, rr[n], rb[n], u[2 * n], u[2 * n + 1], ph[n], st);
    else if (model == 1) lbm::tp_collide<lbm::TP_RK>(p, a, b, rr[n], rb[n], u[2 * n], u[2 * n + 1], ph[n], st);
    else lbm::tp_collide<lbm::TP_CSF>(p, a, b, rr[n], rb[n], u[2 * n], u[2 * n + 1], ph[n], st);
    for (int q = 0; q < 9; q++) { fr[9 * n + q] = a[q]; fb[9 * n + q] = b[q]; }
  }
}

// the constants themselves, for a direct look: {cr, cb, r_val, b_val, s1, s2, s3, t2, t3, w2sum}
void host_tp_constants(int model, const lbm_config* cfg, double* out)
{
  lbm::TpParams p;
  lbm::tp_fill_params(*cfg, (lbm::TpModel)model, p);
  const double v[10] = {p.cr, p.cb, p.r_val, p.b_val, p.s1, p.s2, p.s3, p.t2, p.t3, p.w2sum};
  for (int k = 0; k < 10; k++) out[k] = v[k];
}

}  // extern "C"
