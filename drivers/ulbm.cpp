// Mirrors of the two ulbm::d2q9::kbc drivers (entropic central-moment collision, src/ulbm.cpp), selected by argv[1]:
//   ulbm double-shear-flow [H [T]]   test/ulbm_double_shear_flow.cpp   defaults 128 x 128, T = 10000, snapshot every 10 steps
//   ulbm poiseuille        [H [T]]   test/ulbm_poiseuille.cpp          defaults 128 x 128, T = 300000, snapshot every 100 steps
// The snapshot stacks hold m1 / m0 at the START of iteration t, like the reference's, and are saved under its file names.
#include <cstring>

#include "common.hpp"

int main(int argc, char* argv[])
{
  if (argc < 2) { std::cerr << "usage: ulbm double-shear-flow|poiseuille [H [T]]\n"; return 1; }
  const bool shear = std::strcmp(argv[1], "double-shear-flow") == 0;
  const int H = argc > 2 ? std::atoi(argv[2]) : 128, W = H;
  const int T = argc > 3 ? std::atoi(argv[3]) : (shear ? 10000 : 300000);
  const int period = shear ? 10 : 100;
  const double nu = shear ? 1.70766666E-4 : 1E-4;             // :77 / :72
  const double omega = 1.0 / (0.5 + 3.0 * nu);
  const double u_max = shear ? 0.02 : 0.05;
  std::cout << "T: " << T << "\nH=" << H << "; W=" << W << "\nnu: " << nu << "\nomega: " << omega << "\ntau: " << 1.0 / omega
            << "\nu_max: " << u_max << "\nRe: " << W * u_max / nu << std::endl;

  lbm_config cfg;
  lbm_config_default(&cfg);
  cfg.model = LBM_MODEL_KBC;
  cfg.X = H; cfg.Y = W; cfg.x1 = H;
  cfg.omega = omega;                                          // kbc{H, W, omega}: s2
  lbm_domain* d = nullptr;
  DRV_CHECK(lbm_create(&cfg, &d));
  const size_t N = (size_t)H * W;
  std::vector<double> rho(N, 1.0), u(2 * N, 0.0);             // kbc.m0.fill_(1.0); m1 = 0
  if (shear)
  {
    DRV_CHECK(lbm_preset_periodic(d));                        // the "periodic boundary conditions" block repeats advect's wrap (:118-142)
    const double alpha = 80.0, delta = 0.05;                  // set_initial_conditions (:44-67)
    for (int r = 0; r < H; r++)
      for (int c = 0; c < W; c++)
      {
        u[2 * ((size_t)r * W + c)] = u_max * std::tanh(alpha * (0.25 * H - std::abs(c - 0.5 * H)));
        u[2 * ((size_t)r * W + c) + 1] = u_max * delta * std::sin(6.2832 * (r + 0.25 * H) / H);
      }
    DRV_CHECK(lbm_init_equilibrium(d, 0, LBM_EQ_KBC_FRESH, rho.data(), u.data()));  // kbc.eval_equilibrium(kbc.adve_f) (:97)
  }
  else
  {
    const double p_grad = 8.0 * nu * u_max / (W * W), rho_outlet = 1.0, rho_inlet = 3.0 * (H - 1) * p_grad + rho_outlet;
    std::cout << "grad(p)=" << p_grad << "\nrho_inlet=" << rho_inlet << std::endl;
    DRV_CHECK(lbm_preset_poiseuille(d, rho_inlet, rho_outlet));  // pressure rows on coll_f (:117), bounce-back columns (:121-127)
    std::vector<double> zeros(9 * N, 0.0);
    DRV_CHECK(lbm_set_f(d, 0, zeros.data()));                 // adve_f is never initialised: zeros (src/ulbm.cpp:45)
  }
  DRV_CHECK(lbm_set_moments(d, rho.data(), u.data()));        // the first collide() reads the members m0, m1

  const int Ts = T / period;
  drv::Series ux(H, W, Ts), uy(H, W, Ts), rhos(H, W, Ts);
  std::cout << "main loop starts" << std::endl;
  for (int t = 0; t < T; t += period)
  {
    const int ts = t / period;
    if (t > 0) DRV_CHECK(lbm_get_moments(d, 0, rho.data(), u.data()));
    std::cout << t << "\t\r" << std::flush;
    ux.put(ts, u, 2, 0); uy.put(ts, u, 2, 1); rhos.put(ts, rho, 1, 0);
    DRV_CHECK(lbm_step(d, std::min(period, T - t)));
  }
  DRV_CHECK(lbm_synchronize(d));
  std::cout << "\nsaving results into files" << std::endl;
  if (shear) { ux.save("ulbm-double-shear-flow-ux.pt"); uy.save("ulbm-double-shear-flow-uy.pt"); rhos.save("ulbm-double-shear-flow-rho.pt"); }
  else { ux.save("ulbm-poiseuillehpt-ux.pt"); uy.save("ulbm-poiseuillehpt-uy.pt"); rhos.save("ulbm-poiseuillehpt-rho.pt"); }
  lbm_destroy(d);
  return 0;
}
