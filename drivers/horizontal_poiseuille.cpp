// Mirror of test/horizontal_poiseuille_test.cpp (driver 10): 21x21 incompressible BGK channel,
// pressure-periodic rows, half-way bounce-back columns, 8301 steps, convergence check every 100
// steps, final L2 error against the analytic parabola with the reference's own criterion.
#include "common.hpp"

int main()
{
  // :50-66
  const int T = 8301;
  const int H = 21, W = 21;
  const double tau = std::sqrt(3.0 / 16.0) + 0.5;
  const double omega = 1.0 / tau;
  const double u_max = 1.030985714E-1;
  const double nu = (2.0 * tau - 1.0) / 6.0;
  const double p_grad = 8.0 * nu * u_max / (W * W);
  const double rho_outlet = 1.0;
  const double rho_inlet = 3.0 * (H - 1) * p_grad + rho_outlet;
  std::cout << "T=" << T << "\nH=" << H << "; W=" << W << "\nomega=" << omega << "\nnu=" << nu << "\nRe=" << W * u_max / nu
            << "\ngrad(p)=" << p_grad << "\nrho_inlet=" << rho_inlet << std::endl;

  lbm_config cfg;
  lbm_config_default(&cfg);
  cfg.model = LBM_MODEL_BGK;
  cfg.X = H; cfg.Y = W; cfg.x0 = 0; cfg.x1 = H;
  cfg.omega = omega;
  cfg.equilibrium = LBM_EQ_INCOMPRESSIBLE;
  lbm_domain* d = nullptr;
  DRV_CHECK(lbm_create(&cfg, &d));
  DRV_CHECK(lbm_preset_poiseuille(d, rho_inlet, rho_outlet));

  // :78-91  f_adve = incomp_equilibrium(u = 0, rho = 1)
  const size_t N = (size_t)H * W;
  std::vector<double> u(2 * N, 0.0), rho(N, 1.0), f(9 * N), old_u(2 * N, 1.0);
  DRV_CHECK(lbm_init_equilibrium(d, 0, LBM_EQ_INCOMPRESSIBLE, rho.data(), u.data()));

  drv::Series fs(H, W, T, 9), ux(H, W, T), uy(H, W, T), ps(H, W, T);
  const int t_interval = 100;
  const double tolerance = 1e-12;
  std::cout << "main loop starts" << std::endl;
  for (int t = 0; t < T; t++)
  {
    // :104-110 snapshots: f_adve of this iteration, u/rho of the previous one
    DRV_CHECK(lbm_get_f(d, 0, f.data()));
    fs.put(t, f, 9, 0);
    ux.put(t, u, 2, 0);
    uy.put(t, u, 2, 1);
    ps.put(t, rho, 1, 0, 1.0 / 3.0);
    if (t % t_interval == 1)  // :113-126
    {
      double m = 0.0, mo = 0.0;
      for (size_t n = 0; n < N; n++) { m += u[2 * n]; mo += old_u[2 * n]; }
      const double diff = std::fabs((m / N) / (mo / N) - 1.0);
      if (diff < tolerance) { std::cout << "last t=" << t << std::endl; break; }
      old_u = u;
    }
    // :130-152  moments of f_adve(t) are what iteration t computes; then one fused step
    DRV_CHECK(lbm_get_moments(d, 0, rho.data(), u.data()));
    DRV_CHECK(lbm_step(d, 1));
  }
  DRV_CHECK(lbm_synchronize(d));
  std::cout << "saving results into files" << std::endl;
  ux.save("hpt-ux.pt"); uy.save("hpt-uy.pt"); fs.save("hpt-fs.pt"); ps.save("hpt-ps.pt");

  // :163-175
  double den = 0.0;
  std::vector<double> ua(W);
  for (int j = 0; j < W; j++)
  {
    const double y = (double)(j + 1) - 0.5;
    ua[j] = -4.0 * u_max / (W * W) * y * (y - W);
    den += ua[j] * ua[j];
  }
  den = 1.0 / std::sqrt(den);
  double sum = 0.0;
  for (int r = 1; r < H - 1; r++)
  {
    double e = 0.0;
    for (int j = 0; j < W; j++) { const double dlt = u[2 * ((size_t)r * W + j)] - ua[j]; e += dlt * dlt; }
    sum += std::sqrt(e) * den;
  }
  const double l2 = (1.0 / H) * sum;
  std::cout << "L2=" << l2 << std::endl;
  lbm_destroy(d);
  if (!(l2 <= 1e-11)) { std::cerr << "Large L2 error\n"; return 3; }
  return 0;
}
