// Mirror of test/decompose_domain.cpp (driver 19): two 21x21 sub-domains A over B along axis 0, each an
// lbm_domain slab, advanced in lock step; the "bind" block is the ghost-row exchange of the linked slabs.
#include "common.hpp"

int main()
{
  const int T = 500, H = 21, W = 21;
  const double tau = std::sqrt(3.0 / 16.0) + 0.5, omega = 1.0 / tau, u_max = 1.030985714E-1;
  const double nu = (2.0 * tau - 1.0) / 6.0, p_grad = 8.0 * nu * u_max / (W * W);
  const double rho_outlet = 1.0, rho_inlet = 3.0 * (H - 1) * p_grad + rho_outlet;
  lbm_domain* dom[2] = {nullptr, nullptr};
  for (int k = 0; k < 2; k++)
  {
    lbm_config cfg;
    lbm_config_default(&cfg);
    cfg.X = 2 * H; cfg.Y = W; cfg.x0 = k * H; cfg.x1 = (k + 1) * H;
    cfg.omega = omega;
    cfg.equilibrium = LBM_EQ_COMPRESSIBLE;
    DRV_CHECK(lbm_create(&cfg, &dom[k]));
    DRV_CHECK(lbm_preset_poiseuille(dom[k], rho_inlet, rho_outlet));  // cross-domain pressure rows (:50-73) + walls
  }
  // advect wraps inside each domain (:155-156); the bind joins A's last row and B's first (:181-187)
  DRV_CHECK(lbm_link_neighbours(dom[0], dom[0], dom[1]));
  DRV_CHECK(lbm_link_neighbours(dom[1], dom[0], dom[1]));
  const size_t N = (size_t)H * W;
  std::vector<double> u(2 * N, 0.0), rho(N, 1.0), f(9 * N);
  for (int k = 0; k < 2; k++) DRV_CHECK(lbm_init_equilibrium(dom[k], 0, LBM_EQ_COMPRESSIBLE, rho.data(), u.data()));
  drv::Series fsA(H, W, T, 9), fsB(H, W, T, 9);
  for (int t = 0; t < T; t++)
  {
    DRV_CHECK(lbm_get_f(dom[0], 0, f.data())); fsA.put(t, f, 9, 0);
    DRV_CHECK(lbm_get_f(dom[1], 0, f.data())); fsB.put(t, f, 9, 0);
    DRV_CHECK(lbm_step_group(dom, 2, 1));
  }
  fsA.save("A-domain-decomp-hpt-fs.pt"); fsB.save("B-domain-decomp-hpt-fs.pt");
  lbm_destroy(dom[0]); lbm_destroy(dom[1]);
  return 0;
}
