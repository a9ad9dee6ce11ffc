// Mirror of test/rk_static_droplet_test.cpp (driver 17): Rothman-Keller colour-gradient droplet.
//   rk_static_droplet [L [steps]]     defaults L = 101 (#define L, :8), radius 25 (:9), 2000 steps (:519)
// Adds the Laplace-law evaluation the reference leaves to offline tooling: pressure jump
// p_in - p_out with p = sum_k rho_k * (3/5)(1 - alpha_k), printed next to 1/R.
#include "common.hpp"

static double sigmoid(double x) { return 1.0 / (1.0 + std::exp(-x)); }

int main(int argc, char* argv[])
{
  const int L = argc > 1 ? std::atoi(argv[1]) : 101;
  const int T = argc > 2 ? std::atoi(argv[2]) : 2000;
  const double Radius = L == 101 ? 25.0 : L / 4.0;
  lbm_config cfg;
  lbm_config_default(&cfg);
  cfg.model = LBM_MODEL_RK;
  cfg.X = L; cfg.Y = L; cfg.x1 = L;
  cfg.red = {1.2, 1.0 / 3.0, 1e-4, 0.16, +0.7};   // :504
  cfg.blue = {1.0, 0.2, 1e-4, 0.14, -0.7};        // :506
  cfg.delta = 0.98;                               // :517
  lbm_domain* d = nullptr;
  DRV_CHECK(lbm_create(&cfg, &d));
  DRV_CHECK(lbm_preset_rk(d));
  const size_t N = (size_t)L * L;
  std::vector<double> rr(N), rb(N), u(2 * N, 0.0);  // the driver's 1e-15 gaussian noise in u is not reproduced
  const double Cc = L / 2.0, factor = 2.0;
  for (int r = 0; r < L; r++)
    for (int c = 0; c < L; c++)
    {
      const double s = std::sqrt((r - Cc) * (r - Cc) + (c - Cc) * (c - Cc));  // init_rho (:363-396)
      rr[(size_t)r * L + c] = 1.2 * (1.0 - sigmoid(factor * (s - Radius)));
      rb[(size_t)r * L + c] = 1.0 * sigmoid(factor * (s - Radius));
    }
  DRV_CHECK(lbm_init_two_phase(d, rr.data(), rb.data(), u.data()));
  drv::Series uxs(L, L, T), uys(L, L, T), rhos(L, L, T), rhons(L, L, T);
  std::vector<double> rho(N), ph(N);
  std::cout << "main loop" << std::endl;
  for (int t = 0; t < T; t++)
  {
    DRV_CHECK(lbm_get_phase(d, ph.data(), nullptr, nullptr));  // rhons[t] = phase field at the start of iteration t (:547-548)
    rhons.put(t, ph, 1, 0);
    DRV_CHECK(lbm_step(d, 1));
    DRV_CHECK(lbm_get_moments(d, 0, rho.data(), u.data()));    // :611-613
    uxs.put(t, u, 2, 0); uys.put(t, u, 2, 1); rhos.put(t, rho, 1, 0);
  }
  std::cout << "saving results" << std::endl;
  uxs.save("rk-static-droplet-ux.pt"); uys.save("rk-static-droplet-uy.pt");
  rhos.save("rk-static-droplet-rho.pt"); rhons.save("rk-static-droplet-rhon.pt");
  // Laplace law
  DRV_CHECK(lbm_get_phase(d, ph.data(), rr.data(), rb.data()));
  auto pressure = [&](int r, int c) {
    const size_t n = (size_t)r * L + c;
    return rr[n] * 0.6 * (1.0 - 1.0 / 3.0) + rb[n] * 0.6 * (1.0 - 0.2);
  };
  const double p_in = pressure(L / 2, L / 2), p_out = pressure(2, 2);
  std::cout << "Laplace: p_in - p_out = " << p_in - p_out << " ; 1/R = " << 1.0 / Radius << std::endl;
  lbm_destroy(d);
  return 0;
}
