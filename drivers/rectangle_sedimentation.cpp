// Mirror of test/rectangle_sedimentation_test.cpp (driver 15): fluid BGK lattice + advection-diffusion
// lattice, ABB inlet / extrapolated ABB outlet, specular top, no-slip bottom and rectangle.
//   rectangle_sedimentation <parameters.toml>
#include "common.hpp"

int main(int argc, char* argv[])
{
  if (argc < 2) { std::cerr << "usage: rectangle_sedimentation <parameters.toml>\n"; return 1; }
  lbm_params p;
  DRV_CHECK(lbm_params_from_toml(argv[1], 1, &p));
  drv::print_params(p);
  const int R23 = -151, C28 = 200, C38 = 250;       // :73-75
  const double w_s = 3e-3, scalar_C_w = 1e-3;       // :89-90
  std::cout << R23 << "\n" << C28 << "\n" << C38 << "\nC_w=" << scalar_C_w << std::endl;
  if (p.X <= 151 || p.Y <= C38) { std::cerr << "grid too small for the hard-coded rectangle\n"; return 1; }

  lbm_config cfg;
  lbm_config_default(&cfg);
  cfg.model = LBM_MODEL_BGK_ADE;
  cfg.X = p.X; cfg.Y = p.Y; cfg.x1 = p.X;
  cfg.omega = p.omega; cfg.omega_g = p.omega / 1.0;  // Sc = 1 (:131)
  cfg.equilibrium = LBM_EQ_COMPRESSIBLE;
  cfg.w_s = w_s;
  lbm_domain* d = nullptr;
  DRV_CHECK(lbm_create(&cfg, &d));
  std::vector<double> C_w(p.X, 0.0);
  for (int x = p.X - 50; x < p.X; x++) C_w[x] = scalar_C_w;  // :93
  DRV_CHECK(lbm_preset_sedimentation(d, p.u, C_w.data(), R23, C28, C38));

  // :84-103  u = (0, lp.u), C = C_w on column 0, g = equilibrium(u, C), f = incomp_equilibrium(u, 1)
  const size_t N = (size_t)p.X * p.Y;
  std::vector<double> u(2 * N, 0.0), rho(N, 1.0), C(N, 0.0);
  for (size_t n = 0; n < N; n++) u[2 * n + 1] = p.u;
  for (int x = 0; x < p.X; x++) C[(size_t)x * p.Y] = C_w[x];
  DRV_CHECK(lbm_init_equilibrium(d, 1, LBM_EQ_COMPRESSIBLE, C.data(), u.data()));
  DRV_CHECK(lbm_init_equilibrium(d, 0, LBM_EQ_INCOMPRESSIBLE, rho.data(), u.data()));

  drv::Series ux(p.X, p.Y, p.total_snapshots), uy(p.X, p.Y, p.total_snapshots), ps(p.X, p.Y, p.total_snapshots),
      cs(p.X, p.Y, p.total_snapshots);
  int i = 0;
  std::cout << "main loop\n";
  for (int t = 0; t < p.total_steps; t++)
  {
    if (t % p.snapshot_steps == 0)
    {
      // u, rho, C at the start of iteration t are the moments of the current post-stream state (:199-201,237)
      DRV_CHECK(lbm_get_moments(d, 0, rho.data(), u.data()));
      DRV_CHECK(lbm_get_moments(d, 1, C.data(), nullptr));
      std::cout << t << "; t=" << t * p.dt << " s\t\t\r" << std::flush;
      ux.put(i, u, 2, 0); uy.put(i, u, 2, 1); ps.put(i, rho, 1, 0, 1.0 / 3.0); cs.put(i, C, 1, 0);
      ++i;
    }
    DRV_CHECK(lbm_step(d, 1));
  }
  DRV_CHECK(lbm_synchronize(d));
  std::cout << "\nSaving results" << std::endl;
  const std::string pre = p.file_prefix;
  ux.save(pre + "-ux.pt"); uy.save(pre + "-uy.pt"); ps.save(pre + "-ps.pt"); cs.save(pre + "-cs.pt");
  lbm_destroy(d);
  return 0;
}
