// Mirror of test/cylinder_test.cpp (driver 11; driver 12 free_stream with --no-body):
//   cylinder <parameters.toml> <boundary.toml> [go|a...]      flow past an immersed-boundary cylinder
//   cylinder <parameters.toml> --free-stream [go|a...]        test/free_stream_test.cpp (no body, incompressible eq.)
// argv[3] present skips the y/n prompt; starting with 'a' aborts after printing the parameters (:79-82).
#include <cstring>

#include "common.hpp"

int main(int argc, char* argv[])
{
  if (argc < 3) { std::cerr << "usage: cylinder <parameters.toml> <boundary.toml | --free-stream> [go]\n"; return 1; }
  const bool free_stream = std::strcmp(argv[2], "--free-stream") == 0;
  lbm_params p;
  DRV_CHECK(lbm_params_from_toml(argv[1], 1, &p));
  drv::print_params(p);

  std::vector<double> mx, my;
  if (!free_stream)
  {
    int n = 0;
    DRV_CHECK(lbm_markers_from_toml(argv[2], "cylinder-a", nullptr, nullptr, &n));
    mx.resize(n); my.resize(n);
    DRV_CHECK(lbm_markers_from_toml(argv[2], "cylinder-a", mx.data(), my.data(), &n));
  }
  if (argc < 4) { if (!drv::continue_execution()) return 0; }
  if (argc >= 4 && argv[3][0] == 'a') return 0;

  lbm_config cfg;
  lbm_config_default(&cfg);
  cfg.model = LBM_MODEL_BGK;
  cfg.X = p.X; cfg.Y = p.Y; cfg.x0 = 0; cfg.x1 = p.X;
  cfg.omega = p.omega;
  cfg.equilibrium = free_stream ? LBM_EQ_INCOMPRESSIBLE : LBM_EQ_COMPRESSIBLE;
  cfg.force = free_stream ? LBM_FORCE_NONE : LBM_FORCE_IBM;
  lbm_domain* d = nullptr;
  DRV_CHECK(lbm_create(&cfg, &d));
  // free_stream_test.cpp hard-codes u_w = (0.1, 0) (:70-71); cylinder_test.cpp uses (lp.u, 0) (:74-76)
  const double uwx = free_stream ? 0.1 : p.u;
  DRV_CHECK(lbm_preset_free_stream(d, uwx, 0.0));
  long roi[4] = {0, 0, 0, 0};
  if (!free_stream)
  {
    DRV_CHECK(lbm_ibm_set_markers(d, mx.data(), my.data(), (int)mx.size(), 5));
    DRV_CHECK(lbm_ibm_get_roi(d, roi));
    std::cout << "markers.size=" << mx.size() << "\nrows=" << roi[0] << ":" << roi[1] << "\ncols=" << roi[2] << ":" << roi[3] << std::endl;
  }

  const size_t N = (size_t)p.X * p.Y;
  std::vector<double> u(2 * N, 0.0), rho(N, 1.0);
  for (size_t n = 0; n < N; n++) u[2 * n] = uwx;
  DRV_CHECK(lbm_init_equilibrium(d, 0, LBM_EQ_INCOMPRESSIBLE, rho.data(), u.data()));  // :86

  const long RR = roi[1] - roi[0], RC = roi[3] - roi[2];
  drv::Series ux(p.X, p.Y, p.total_snapshots), uy(p.X, p.Y, p.total_snapshots), ps(p.X, p.Y, p.total_snapshots);
  std::vector<double> F((size_t)std::max(1L, RR * RC * 2), 0.0), Fs_series((size_t)2 * p.total_snapshots, 0.0);
  drv::Series forces(std::max(1L, RR), std::max(1L, RC), p.total_snapshots, 2);
  double Fs[2] = {0.0, 0.0};
  int i = 0;
  for (int t = 0; t < p.total_steps; t++)
  {
    if (t % p.snapshot_steps == 0)  // sp.snapshot(t), :90-99
    {
      std::cout << t << "; t=" << t * p.dt << " s\t\r" << std::flush;
      ux.put(i, u, 2, 0); uy.put(i, u, 2, 1); ps.put(i, rho, 1, 0, 1.0 / 3.0);
      if (!free_stream && i < p.total_snapshots)
      {
        forces.put(i, F, 2, 0);
        Fs_series[i] = Fs[0];
        Fs_series[p.total_snapshots + i] = Fs[1];
      }
      ++i;
    }
    const bool want = ((t + 1) % p.snapshot_steps == 0) || t + 1 == p.total_steps;
    if (want) DRV_CHECK(lbm_get_moments(d, 0, rho.data(), u.data()));  // u, rho of iteration t (:101-104)
    DRV_CHECK(lbm_step(d, 1));
    if (want && !free_stream)
    {
      DRV_CHECK(lbm_ibm_get_force(d, F.data()));  // F of iteration t (:110); F_s = sum over the ROI (:112)
      Fs[0] = Fs[1] = 0.0;
      for (long n = 0; n < RR * RC; n++) { Fs[0] += F[2 * n]; Fs[1] += F[2 * n + 1]; }
    }
  }
  DRV_CHECK(lbm_synchronize(d));
  std::cout << "\nSaving results" << std::endl;
  const std::string pre = std::string(p.file_prefix) + (free_stream ? "fst-" : "ct-");
  ux.save(pre + "ux.pt"); uy.save(pre + "uy.pt"); ps.save(pre + "ps.pt");
  if (!free_stream)
  {
    drv::save_array(pre + "Fs.pt", Fs_series, {2, (long)p.total_snapshots});
    forces.save(pre + "F.pt");
  }
  lbm_destroy(d);
  return 0;
}
