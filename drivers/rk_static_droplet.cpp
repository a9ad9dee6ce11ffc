// Mirror of test/rk_static_droplet_test.cpp (driver 17): Rothman-Keller colour-gradient droplet.
//   rk_static_droplet [L [steps [full]]]   defaults L = 101 (#define L, :8), radius 25 (:9), 2000 steps (:519), full = 1
// full = 1 writes all nineteen files of the reference (:617-635: populations, normal, curvature, interfacial tension,
// kappa, 1/tau, the red colour's omega1/2/3 — 8.8 GB of host memory at 101 x 101 x 2000, as in the reference);
// full = 0 keeps the four state files (ux, uy, rho, rhon).
// Adds the Laplace-law evaluation the reference leaves to offline tooling: pressure jump
// p_in - p_out with p = sum_k rho_k * (3/5)(1 - alpha_k), printed next to 1/R.
#include "common.hpp"

static double sigmoid(double x) { return 1.0 / (1.0 + std::exp(-x)); }

int main(int argc, char* argv[])
{
  const int L = argc > 1 ? std::atoi(argv[1]) : 101;
  const int T = argc > 2 ? std::atoi(argv[2]) : 2000;
  const bool full = argc > 3 ? std::atoi(argv[3]) != 0 : true;
  const double Radius = L == 101 ? 25.0 : L / 4.0;
  const double sigma = 5e-3;  // :499
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
  const int Td = full ? T : 0;  // the diagnostic stacks (:521-540)
  drv::Series nxs(L, L, Td), nys(L, L, Td), Ks(L, L, Td), Fsxs(L, L, Td), Fsys(L, L, Td), norms(L, L, Td), gradxs(L, L, Td),
      gradys(L, L, Td), rparams(L, L, Td);
  drv::Series r_fs(L, L, Td, 9), b_fs(L, L, Td, 9), kappas(L, L, Td, 9), omega1s(L, L, Td, 9), omega2s(L, L, Td, 9), omega3s(L, L, Td, 9);
  std::vector<double> rho(N), ph(N), grad(full ? 2 * N : 0), norm(full ? N : 0), nn(full ? 2 * N : 0), K(full ? N : 0),
      Fs(full ? 2 * N : 0), kappa(full ? 9 * N : 0), rp(full ? N : 0), o1(full ? 9 * N : 0), o2(full ? 9 * N : 0), o3(full ? 9 * N : 0),
      f(full ? 9 * N : 0);
  std::cout << "main loop" << std::endl;
  for (int t = 0; t < T; t++)
  {
    if (full)
    {
      // everything the reference evaluates before the step (:546-589), from the state at the top of the iteration
      lbm_rk_diag dg{};
      dg.phase = ph.data(); dg.grad = grad.data(); dg.norm = norm.data(); dg.n = nn.data(); dg.K = K.data(); dg.Fs = Fs.data();
      dg.kappa = kappa.data(); dg.rparams = rp.data(); dg.omega1 = o1.data(); dg.omega2 = o2.data(); dg.omega3 = o3.data();
      DRV_CHECK(lbm_rk_diagnostics(d, sigma, &dg));
      norms.put(t, norm, 1, 0); gradxs.put(t, grad, 2, 0); gradys.put(t, grad, 2, 1); nxs.put(t, nn, 2, 0); nys.put(t, nn, 2, 1);
      Ks.put(t, K, 1, 0); Fsxs.put(t, Fs, 2, 0); Fsys.put(t, Fs, 2, 1); kappas.put(t, kappa, 9, 0); rparams.put(t, rp, 1, 0);
      omega1s.put(t, o1, 9, 0); omega2s.put(t, o2, 9, 0); omega3s.put(t, o3, 9, 0);
    }
    else
      DRV_CHECK(lbm_get_phase(d, ph.data(), nullptr, nullptr));
    rhons.put(t, ph, 1, 0);  // rhons[t] = phase field at the start of iteration t (:547-548)
    DRV_CHECK(lbm_step(d, 1));
    if (full)
    {
      DRV_CHECK(lbm_get_f(d, 0, f.data())); r_fs.put(t, f, 9, 0);  // adv_f after the step (:592, :597)
      DRV_CHECK(lbm_get_f(d, 1, f.data())); b_fs.put(t, f, 9, 0);
    }
    DRV_CHECK(lbm_get_moments(d, 0, rho.data(), u.data()));    // :611-613
    uxs.put(t, u, 2, 0); uys.put(t, u, 2, 1); rhos.put(t, rho, 1, 0);
  }
  std::cout << "saving results" << std::endl;
  uxs.save("rk-static-droplet-ux.pt"); uys.save("rk-static-droplet-uy.pt");
  rhos.save("rk-static-droplet-rho.pt"); rhons.save("rk-static-droplet-rhon.pt");
  if (full)
  {
    r_fs.save("rk-static-droplet-r-fs.pt"); b_fs.save("rk-static-droplet-b-fs.pt");
    nxs.save("rk-static-droplet-nx.pt"); nys.save("rk-static-droplet-ny.pt"); Ks.save("rk-static-droplet-ks.pt");
    norms.save("rk-static-droplet-norms.pt"); Fsxs.save("rk-static-droplet-fx.pt"); Fsys.save("rk-static-droplet-fy.pt");
    gradxs.save("rk-static-droplet-gradx.pt"); gradys.save("rk-static-droplet-grady.pt"); rparams.save("rk-static-droplet-rparams.pt");
    kappas.save("rk-static-droplet-kappas.pt"); omega1s.save("rk-static-droplet-omegas1.pt");
    omega2s.save("rk-static-droplet-omegas2.pt"); omega3s.save("rk-static-droplet-omegas3.pt");
  }
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
