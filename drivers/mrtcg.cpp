// Mirrors of the three MRT colour-gradient drivers, selected by argv[1]:
//   mrtcg rayleigh-taylor <file.toml>   test/mrtcg_rayleigh_taylor.cpp (driver 16; needs [general])
//   mrtcg static-droplet  <file.toml>   test/mrtcg_static_droplet.cpp  (driver 18; sigma = 0.1, Fg = (0,-6.25e-6), source not added)
//   mrtcg csf             <file.toml>   test/mrt_rayleigh_taylor.cpp   (continuum-surface-force variant; also saves gradx / grady)
#include <cstring>

#include "common.hpp"

static double sigmoid(double x) { return 1.0 / (1.0 + std::exp(-x)); }

int main(int argc, char* argv[])
{
  if (argc < 3) { std::cerr << "usage: mrtcg rayleigh-taylor|static-droplet|csf <file.toml>\n"; return 1; }
  const bool csf = std::strcmp(argv[1], "csf") == 0;
  const bool rt = csf || std::strcmp(argv[1], "rayleigh-taylor") == 0;
  lbm_two_phase_params tp;
  DRV_CHECK(lbm_two_phase_from_toml(argv[2], rt ? 1 : 0, &tp));
  lbm_colour red, blue;
  DRV_CHECK(lbm_colour_from_toml(argv[2], "red", &red));
  DRV_CHECK(lbm_colour_from_toml(argv[2], "blue", &blue));
  const int R = tp.rows, C = tp.columns;
  std::cout << "DOMAIN parameters:\nR=" << R << "\nC=" << C << "\nT=" << tp.time_steps << "\nnr_snapshots=" << tp.nr_snapshots
            << "\nperiod_snapshots=" << tp.period_snapshots << std::endl;

  lbm_config cfg;
  lbm_config_default(&cfg);
  cfg.model = csf ? LBM_MODEL_MRT_CSF : LBM_MODEL_MRTCG;
  cfg.X = R; cfg.Y = C; cfg.x1 = R;
  cfg.red = {red.rho_0, red.alpha, red.A, red.nu, red.beta};
  cfg.blue = {blue.rho_0, blue.alpha, blue.A, blue.nu, blue.beta};
  cfg.delta = 0.1;  // relaxation_function{r, b, 0.1}
  if (rt) { cfg.sigma = tp.sigma; cfg.Fg[0] = tp.gravity_magnitude; cfg.Fg[1] = 0.0; cfg.add_force = 1; }
  else { cfg.sigma = 0.1; cfg.Fg[0] = 0.0; cfg.Fg[1] = -6.25e-6; cfg.add_force = 0; }
  lbm_domain* d = nullptr;
  DRV_CHECK(lbm_create(&cfg, &d));
  DRV_CHECK(lbm_preset_mrtcg(d));

  // initial densities: init_rho_cosine (RT :182-210) / init_rho_droplet (droplet :182-204)
  const size_t N = (size_t)R * C;
  std::vector<double> rr(N), rb(N), u(2 * N, 0.0);
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++)
    {
      const size_t n = (size_t)r * C + c;
      if (rt)
      {
        const double s = R / 2.0 - 0.1 * C * std::cos(2.0 * 3.141592 * c / C);
        rr[n] = red.rho_0 * (r < s ? 1.0 : 0.0);
        rb[n] = blue.rho_0 * (r >= s ? 1.0 : 0.0);
      }
      else
      {
        const double ctr = R / 2.0, s = std::sqrt((r - ctr) * (r - ctr) + (c - ctr) * (c - ctr));
        rr[n] = red.rho_0 * (1.0 - sigmoid(1.0 * (s - 25.0)));
        rb[n] = blue.rho_0 * sigmoid(1.0 * (s - 25.0));
        const double rho = rr[n] + rb[n];  // droplet driver shifts the initial u (:457)
        u[2 * n] = 0.0 + 0.5 * cfg.Fg[0] / rho;
        u[2 * n + 1] = 0.0 + 0.5 * cfg.Fg[1] / rho;
      }
    }
  DRV_CHECK(lbm_init_two_phase(d, rr.data(), rb.data(), u.data()));

  const int S = tp.nr_snapshots;
  drv::Series rhos(R, C, S), uxs(R, C, S), uys(R, C, S), phases(R, C, S);
  drv::Series gradx(R, C, csf ? S : 0), grady(R, C, csf ? S : 0);
  std::vector<double> rho(N), ph(N, 0.0), Fs(csf ? 2 * N : 0, 0.0);
  std::cout << "main loop" << std::endl;
  for (int t = 0; t < tp.time_steps; t++)
  {
    if (t % tp.period_snapshots == 0)
    {
      const int k = t / tp.period_snapshots;
      DRV_CHECK(lbm_get_moments(d, 0, rho.data(), u.data()));
      rhos.put(k, rho, 1, 0); uxs.put(k, u, 2, 0); uys.put(k, u, 2, 1);
      phases.put(k, ph, 1, 0);  // the driver stores the phase field of the PREVIOUS iteration (:421)
      if (csf) { gradx.put(k, Fs, 2, 0); grady.put(k, Fs, 2, 1); }  // ... and its interfacial tension (mrt_rayleigh_taylor.cpp:485-486)
    }
    if ((t + 1) % tp.period_snapshots == 0) DRV_CHECK(lbm_get_phase(d, ph.data(), nullptr, nullptr));
    DRV_CHECK(lbm_step(d, 1));
    if (csf && (t + 1) % tp.period_snapshots == 0) DRV_CHECK(lbm_get_interfacial_tension(d, Fs.data()));
  }
  DRV_CHECK(lbm_synchronize(d));
  std::cout << "save snapshots" << std::endl;
  const std::string pre = rt ? std::string(tp.name) + "-mrtcg-rayleigh-taylor-" : std::string("mrtcg-static-droplet-");
  rhos.save(pre + "rhos.pt"); uxs.save(pre + "uxs.pt"); uys.save(pre + "uys.pt"); phases.save(pre + "phases.pt");
  if (csf) { gradx.save(pre + "gradx.pt"); grady.save(pre + "grady.pt"); }
  lbm_destroy(d);
  return 0;
}
