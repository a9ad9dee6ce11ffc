// Mirrors of the two remaining 21x21 / 51x51 channel drivers, selected by argv[1]:
//   channel_variants specular   test/specular_boundary_test.cpp (driver 13): compressible eq., specular columns
//   channel_variants gravity    test/gravity_test.cpp (driver 14): incompressible eq., Guo forcing Fg = (-0.0003, 0)
// Both save fs / ux / uy / ps of every step like the reference ({H,W,9,T}, {H,W,T}).
#include <cstring>

#include "common.hpp"

int main(int argc, char* argv[])
{
  if (argc < 2) { std::cerr << "usage: channel_variants specular|gravity [steps]\n"; return 1; }
  const bool grav = std::strcmp(argv[1], "gravity") == 0;
  const int T = argc > 2 ? std::atoi(argv[2]) : 10000;                      // both drivers: T = 10000
  const int H = grav ? 21 : 51, W = H;
  const double tau = std::sqrt(3.0 / 16.0) + 0.5, omega = 1.0 / tau, u_max = 0.1;
  const double nu = (2.0 * tau - 1.0) / 6.0, p_grad = 8.0 * nu * u_max / (W * W);
  const double rho_outlet = 1.0;
  const double rho_inlet = grav ? rho_outlet : 3.0 * (H - 1) * p_grad + rho_outlet;  // gravity_test.cpp:76
  lbm_config cfg;
  lbm_config_default(&cfg);
  cfg.X = H; cfg.Y = W; cfg.x1 = H; cfg.omega = omega;
  cfg.equilibrium = grav ? LBM_EQ_INCOMPRESSIBLE : LBM_EQ_COMPRESSIBLE;
  if (grav) { cfg.force = LBM_FORCE_UNIFORM; cfg.Fg[0] = -0.0003; cfg.Fg[1] = 0.0; }
  lbm_domain* d = nullptr;
  DRV_CHECK(lbm_create(&cfg, &d));
  if (grav) DRV_CHECK(lbm_preset_poiseuille(d, rho_inlet, rho_outlet));
  else DRV_CHECK(lbm_preset_specular_channel(d, rho_inlet, rho_outlet));
  const size_t N = (size_t)H * W;
  std::vector<double> u(2 * N, 0.0), rho(N, 1.0), f(9 * N);
  DRV_CHECK(lbm_init_equilibrium(d, 0, LBM_EQ_INCOMPRESSIBLE, rho.data(), u.data()));
  drv::Series fs(H, W, T, 9), ux(H, W, T), uy(H, W, T), ps(H, W, T);
  for (int t = 0; t < T; t++)
  {
    DRV_CHECK(lbm_get_f(d, 0, f.data()));
    fs.put(t, f, 9, 0); ux.put(t, u, 2, 0); uy.put(t, u, 2, 1); ps.put(t, rho, 1, 0, 1.0 / 3.0);
    DRV_CHECK(lbm_get_moments(d, 0, rho.data(), u.data()));
    DRV_CHECK(lbm_step(d, 1));
  }
  const std::string pre = grav ? "gt-" : "sbt-";
  ux.save(pre + "ux.pt"); uy.save(pre + "uy.pt"); fs.save(pre + "fs.pt"); ps.save(pre + "ps.pt");
  lbm_destroy(d);
  return 0;
}
