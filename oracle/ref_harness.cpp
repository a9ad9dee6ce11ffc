// TEST INFRASTRUCTURE — C-callable harness over the UNMODIFIED reference library.
//
// Links the reference's own src/*.cpp objects (compiled in place from
// /root/reference, CPU libtorch; see Makefile) and exposes their public
// functions to ctypes with plain pointers in the reference's AoS layout
// ({X,Y,9} row-major, fp64).  Used by tests/ to pin oracle/lbm_oracle.c and to
// generate tests/golden/*.  No arithmetic of its own: every number comes out of
// a reference function.
#include <torch/torch.h>
#include <toml++/toml.hpp>

#include <chrono>
#include <cstring>
#include <iostream>
#include <string>

#include "src/colour.hpp"
#include "src/differential.hpp"
#include "src/domain.hpp"
#include "src/ibm.hpp"
#include "src/params.hpp"
#include "src/solver.hpp"
#include "src/ulbm.hpp"

namespace
{
thread_local std::string g_err;

torch::Tensor wrap(const double* p, std::initializer_list<int64_t> shape)
{
  return torch::from_blob(const_cast<double*>(p), shape, torch::TensorOptions().dtype(torch::kDouble)).clone();
}

void copy_out(const torch::Tensor& t, double* out)
{
  auto c = t.contiguous().to(torch::kDouble);
  std::memcpy(out, c.data_ptr<double>(), sizeof(double) * c.numel());
}

// Sets the default dtype like every reference main() does (e.g. test/cylinder_test.cpp:41) and
// mutes the reference's std::cout chatter for the duration of a call.
struct dtype_guard
{
  std::ios_base::iostate saved;
  dtype_guard() : saved(std::cout.rdstate())
  {
    torch::set_default_dtype(caffe2::scalarTypeToTypeMeta(torch::kDouble));
    std::cout.setstate(std::ios_base::failbit);
  }
  ~dtype_guard() { std::cout.clear(saved); }
};
}  // namespace

#define REF_TRY try { dtype_guard guard__;
#define REF_CATCH \
  } catch (const std::exception& e) { g_err = e.what(); return 1; } \
  return 0;

extern "C"
{

const char* ref_last_error() { return g_err.c_str(); }

void ref_set_num_threads(int n) { at::set_num_threads(n); }
int ref_get_num_threads() { return at::get_num_threads(); }

// solver::E, solver::c (src/solver.cpp:12-21)
int ref_constants(double* w9, double* c18)
{
  REF_TRY
  copy_out(solver::E, w9);
  copy_out(solver::c, c18);
  REF_CATCH
}

// src/solver.cpp:23-26
int ref_calc_rho(const double* f, int X, int Y, double* rho)
{
  REF_TRY
  auto tf = wrap(f, {X, Y, 9});
  auto r = torch::zeros({X, Y, 1});
  solver::calc_rho(r, tf);
  copy_out(r, rho);
  REF_CATCH
}

// src/solver.cpp:34-37
int ref_calc_u(const double* f, const double* rho, int X, int Y, double* u)
{
  REF_TRY
  auto tf = wrap(f, {X, Y, 9});
  auto tr = wrap(rho, {X, Y, 1});
  auto tu = torch::zeros({X, Y, 2});
  solver::calc_u(tu, tf, tr);
  copy_out(tu, u);
  REF_CATCH
}

// src/solver.cpp:28-31
int ref_calc_incomp_u(const double* f, int X, int Y, double* u)
{
  REF_TRY
  auto tf = wrap(f, {X, Y, 9});
  auto tu = torch::zeros({X, Y, 2});
  solver::calc_incomp_u(tu, tf);
  copy_out(tu, u);
  REF_CATCH
}

// src/solver.cpp:51-62
int ref_equilibrium(const double* u, const double* rho, int X, int Y, double* feq)
{
  REF_TRY
  auto tu = wrap(u, {X, Y, 2});
  auto tr = wrap(rho, {X, Y, 1});
  auto fe = torch::zeros({X, Y, 9});
  solver::equilibrium(fe, tu, tr);
  copy_out(fe, feq);
  REF_CATCH
}

// src/solver.cpp:39-49
int ref_incomp_equilibrium(const double* u, const double* rho, int X, int Y, double* feq)
{
  REF_TRY
  auto tu = wrap(u, {X, Y, 2});
  auto tr = wrap(rho, {X, Y, 1});
  auto fe = torch::zeros({X, Y, 9});
  solver::incomp_equilibrium(fe, tu, tr);
  copy_out(fe, feq);
  REF_CATCH
}

// src/solver.cpp:65-74
int ref_collision(const double* f, const double* feq, double omega, int X, int Y, double* fcoll)
{
  REF_TRY
  auto tf = wrap(f, {X, Y, 9});
  auto te = wrap(feq, {X, Y, 9});
  auto tc = torch::zeros({X, Y, 9});
  solver::collision(tc, tf, te, omega);
  copy_out(tc, fcoll);
  REF_CATCH
}

// src/solver.cpp:76-131
int ref_advect(const double* f, int X, int Y, double* g)
{
  REF_TRY
  auto tf = wrap(f, {X, Y, 9});
  auto tg = torch::zeros({X, Y, 9});
  solver::advect(tg, tf);
  copy_out(tg, g);
  REF_CATCH
}

// src/differential.cpp:23-39 ; out_x, out_y are {R,C}
int ref_differential(const double* psi, int R, int C, double* out_x, double* out_y)
{
  REF_TRY
  static differential D{};
  auto tp = wrap(psi, {R, C});
  copy_out(D.x(tp), out_x);
  copy_out(D.y(tp), out_y);
  REF_CATCH
}

// src/params.cpp:7-120.  out[] = {fp.nu, fp.u, fp.l, fp.rho_0, fp.Re,
//   lp.tau, lp.omega, lp.Re, lp.nu, lp.l, lp.dx, lp.dt, lp.T, lp.u, lp.X, lp.Y,
//   sp.stop_time, sp.snapshot_period, sp.total_steps, sp.snapshot_steps, sp.total_snapshots}
// (the simulation block is filled only when with_simulation != 0).
int ref_params(const char* toml_path, int with_simulation, double* out)
{
  REF_TRY
  toml::table tbl = toml::parse_file(toml_path);
  const params::flow fp{tbl};
  const params::lattice lp{tbl, fp};
  int k = 0;
  out[k++] = fp.nu; out[k++] = fp.u; out[k++] = fp.l; out[k++] = fp.rho_0; out[k++] = fp.Re;
  out[k++] = lp.tau; out[k++] = lp.omega; out[k++] = lp.Re; out[k++] = lp.nu; out[k++] = lp.l;
  out[k++] = lp.dx; out[k++] = lp.dt; out[k++] = lp.T; out[k++] = lp.u; out[k++] = lp.X; out[k++] = lp.Y;
  if (with_simulation)
  {
    const params::simulation sp{tbl, lp};
    out[k++] = sp.stop_time; out[k++] = sp.snapshot_period; out[k++] = sp.total_steps;
    out[k++] = sp.snapshot_steps; out[k++] = sp.total_snapshots;
  }
  REF_CATCH
}

// src/colour.cpp:11-64.  out[] = {rho_0, alpha, A, nu, mu, beta, cs2, ics2, rlx, phi[9], eta[9]}
int ref_colour(const char* toml_path, const char* table_name, double* out)
{
  REF_TRY
  toml::table tbl = toml::parse_file(toml_path);
  const torch::Tensor E = solver::c;
  colour k{tbl[table_name], 2, 2, E};
  int n = 0;
  out[n++] = k.rho_0; out[n++] = k.alpha; out[n++] = k.A; out[n++] = k.nu; out[n++] = k.mu;
  out[n++] = k.beta; out[n++] = k.cs2; out[n++] = k.ics2; out[n++] = k.rlx;
  copy_out(k.phi, out + n); n += 9;
  copy_out(k.eta.index({0, 0}), out + n);
  REF_CATCH
}

// src/ibm.cpp:60-190.  roi[4] = {row_start,row_stop,col_start,col_stop};
// F_out must hold (row_stop-row_start)*(col_stop-col_start)*2 doubles; pass
// F_out = nullptr to query the ROI only.
int ref_ibm_force(const char* toml_path, const char* name, const double* u, const double* rho,
                  int X, int Y, long* roi, double* F_out)
{
  REF_TRY
  toml::table tbl = toml::parse_file(toml_path);
  ibm ib{tbl, name, torch::kCPU};
  roi[0] = ib.rows.start().maybe_as_int().value();
  roi[1] = ib.rows.stop().maybe_as_int().value();
  roi[2] = ib.cols.start().maybe_as_int().value();
  roi[3] = ib.cols.stop().maybe_as_int().value();
  if (F_out)
  {
    auto tu = wrap(u, {X, Y, 2});
    auto tr = wrap(rho, {X, Y, 1});
    auto F = ib.eulerian_force_density(tu, tr);
    copy_out(F, F_out);
  }
  REF_CATCH
}

// Times the loop body of test/cylinder_test.cpp:100-163 (the statements below are that loop,
// calling the reference's own solver:: / ibm functions; snapshots and prompts left out) on a
// {X,Y} grid: `warmup` untimed + `steps` timed iterations.  Used by bench.py as the CPU baseline
// ("kind": "reference").  markers_toml holds `[cylinder-a] x=[..] y=[..]`.
int ref_cylinder_loop(int X, int Y, double omega, double u_lb, const char* markers_toml, int warmup, int steps,
                      double* seconds_per_step, double* checksum)
{
  REF_TRY
  using torch::Tensor;
  using torch::indexing::Slice;
  using torch::indexing::Ellipsis;
  using solver::E;
  using solver::c;
  toml::table tbl_boundary = toml::parse_file(markers_toml);
  const torch::Device dev = torch::kCPU;
  Tensor f_equi = torch::zeros({X, Y, 9}, dev);
  Tensor f_coll = torch::zeros_like(f_equi, dev);
  Tensor f_adve = torch::zeros_like(f_equi, dev);
  Tensor u = torch::zeros({X, Y, 2}, dev);
  Tensor rho = torch::ones({X, Y, 1}, dev);
  ibm ib{tbl_boundary, "cylinder-a", dev};
  Tensor F, S;
  const double ics2 = 1.0 / 3.0, ics4 = 1.0 / 9.0;
  Tensor equi_populations = torch::zeros_like(f_equi);
  Tensor u_w = torch::zeros({Y, 2}, dev);
  u_w.index({Slice(), 0}) = u_lb;
  u.index({Ellipsis, 0}) = u_lb;
  Tensor abb_bc = torch::zeros({Y, 1}, dev);
  solver::incomp_equilibrium(f_adve, u, rho);
  std::chrono::steady_clock::time_point t0;
  for (int t = 0; t < warmup + steps; t++)
  {
    if (t == warmup) t0 = std::chrono::steady_clock::now();
    solver::calc_rho(rho, f_adve);
    solver::calc_u(u, f_adve, rho);
    solver::equilibrium(f_equi, u, rho);
    equi_populations.copy_(-omega * (f_adve - f_equi));
    F = ib.eulerian_force_density(u, rho);
    auto u_roi = u.index({ib.rows, ib.cols, Slice()});
    S = ((1 - 0.5 * omega) * ((ics2 + ics4 * u_roi.matmul(c)) * F.matmul(c) - ics2 * (u_roi * F).sum(2).unsqueeze(2)) * E).clone().detach();
    f_coll.copy_(f_adve + equi_populations);
    f_coll.index({ib.rows, ib.cols, Slice()}) += S;
    solver::advect(f_adve, f_coll);
    const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    for (int row : {0, -1})
    {
      abb_bc = ((2.0 + 9.0 * torch::pow(u_w.matmul(c), 2.0) - 3.0 * u_w.mul(u_w).sum(1).unsqueeze(1)) * E).squeeze(0).clone().detach();
      for (int q : {1, 2, 3, 4, 5, 6, 7, 8})
        f_adve.index({row, Slice(), opp[q]}) = (-f_coll.index({row, Slice(), q}) + abb_bc.index({Slice(), q})).clone().detach();
    }
    f_adve.index({Slice(), -1, 4}) = f_coll.index({Slice(), -1, 2}).clone().detach();
    f_adve.index({Slice(), -1, 7}) = f_coll.index({Slice(), -1, 6}).clone().detach();
    f_adve.index({Slice(), -1, 8}) = f_coll.index({Slice(), -1, 5}).clone().detach();
    f_adve.index({Slice(), 0, 2}) = f_coll.index({Slice(), 0, 4}).clone().detach();
    f_adve.index({Slice(), 0, 5}) = f_coll.index({Slice(), 0, 8}).clone().detach();
    f_adve.index({Slice(), 0, 6}) = f_coll.index({Slice(), 0, 7}).clone().detach();
  }
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  *seconds_per_step = dt / steps;
  *checksum = f_adve.sum().item<double>();
  REF_CATCH
}

// The time loop of test/horizontal_poiseuille_test.cpp:100-153 (moments, incompressible equilibrium, solver::collision,
// the pressure-periodic rows of :25-45, solver::advect, bounce-back columns; snapshots and the convergence check left
// out) on an {H,W} grid, every operator the reference's own.  CPU baseline of bench.py's poiseuille workload.
int ref_poiseuille_loop(int H, int W, double omega, double rho_inlet, double rho_outlet, int warmup, int steps,
                        double* seconds_per_step, double* checksum)
{
  REF_TRY
  using torch::Tensor;
  using torch::indexing::Slice;
  using torch::indexing::Ellipsis;
  const torch::Device dev = torch::kCPU;
  Tensor f_equi = torch::zeros({H, W, 9}, dev);
  Tensor f_coll = torch::zeros_like(f_equi);
  Tensor f_adve = torch::zeros_like(f_equi);
  Tensor u = torch::zeros({H, W, 2}, dev);
  Tensor rho = torch::ones({H, W, 1}, dev);
  Tensor temp_equi = torch::zeros({1, W, 9}, dev);
  Tensor temp_rho = torch::ones({1, W, 1}, dev);
  solver::incomp_equilibrium(f_adve, u, rho);
  std::chrono::steady_clock::time_point t0;
  for (int t = 0; t < warmup + steps; t++)
  {
    if (t == warmup) t0 = std::chrono::steady_clock::now();
    solver::calc_rho(rho, f_adve);
    solver::calc_incomp_u(u, f_adve);
    solver::incomp_equilibrium(f_equi, u, rho);
    solver::collision(f_coll, f_adve, f_equi, omega);
    solver::incomp_equilibrium(temp_equi, u.index({-2, Ellipsis}).unsqueeze(0), rho_inlet * temp_rho);
    f_coll.index({0, Ellipsis}) = (temp_equi + f_coll.index({-2, Ellipsis}) - f_equi.index({-2, Ellipsis})).squeeze(0).clone().detach();
    solver::incomp_equilibrium(temp_equi, u.index({1, Ellipsis}).unsqueeze(0), rho_outlet * temp_rho);
    f_coll.index({-1, Ellipsis}) = (temp_equi + f_coll.index({1, Ellipsis}) - f_equi.index({1, Ellipsis})).squeeze(0).clone().detach();
    solver::advect(f_adve, f_coll);
    f_adve.index({Slice(), -1, 4}) = f_coll.index({Slice(), -1, 2}).clone().detach();
    f_adve.index({Slice(), -1, 7}) = f_coll.index({Slice(), -1, 5}).clone().detach();
    f_adve.index({Slice(), -1, 8}) = f_coll.index({Slice(), -1, 6}).clone().detach();
    f_adve.index({Slice(), 0, 2}) = f_coll.index({Slice(), 0, 4}).clone().detach();
    f_adve.index({Slice(), 0, 5}) = f_coll.index({Slice(), 0, 7}).clone().detach();
    f_adve.index({Slice(), 0, 6}) = f_coll.index({Slice(), 0, 8}).clone().detach();
  }
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  *seconds_per_step = dt / steps;
  *checksum = f_adve.sum().item<double>();
  REF_CATCH
}

// struct domain (src/domain.cpp:3-12): shapes of the five buffers, 15 ints.
int ref_domain_shapes(int R, int C, long* shapes15)
{
  REF_TRY
  domain d{R, C};
  const torch::Tensor* ts[5] = {&d.adve_f, &d.equi_f, &d.coll_f, &d.m_0, &d.m_1};
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 3; j++) shapes15[3 * i + j] = ts[i]->size(j);
  REF_CATCH
}

// ulbm::d2q9::kbc (src/ulbm.hpp:11-90) stepped like its two drivers do.
//   bc = 0  test/ulbm_double_shear_flow.cpp:118-142  fully periodic (the "bind to self" assignments repeat advect's wrap)
//   bc = 1  test/ulbm_poiseuille.cpp:116-143         pressure-periodic rows on coll_f, bounce-back columns
// m0 {R,C}, m1 {R,C,2}, adve {R,C,9} in/out.  The few tensor expressions of the drivers' loop bodies
// (moments, the pressure rule) are repeated here verbatim; collide() and advect() are the reference's.
int ref_kbc_run(int R, int C, double s2, double* m0, double* m1, double* adve, int steps, int bc, double rho_in, double rho_out)
{
  REF_TRY
  using torch::indexing::Ellipsis;
  using torch::indexing::None;
  using torch::indexing::Slice;
  ulbm::d2q9::kbc kbc{R, C, s2};
  kbc.m0 = wrap(m0, {R, C});
  kbc.m1 = wrap(m1, {R, C, 2});
  kbc.adve_f = wrap(adve, {R, C, 9});
  const torch::Tensor c = solver::c;
  for (int t = 0; t < steps; t++)
  {
    kbc.collide();
    if (bc == 1)
    {
      // periodic_boundary_condition(kbc.coll_f, kbc.iequi_f.pow(-1), kbc.m1, kbc.m0, rho_inlet, rho_outlet)  (ulbm_poiseuille.cpp:39-60,117)
      torch::Tensor f_equi = kbc.iequi_f.pow(-1);
      torch::Tensor temp_equi = torch::zeros({1, C, 9});
      torch::Tensor temp_rho = torch::ones({1, C, 1});
      solver::incomp_equilibrium(temp_equi, kbc.m1.index({-2, Ellipsis}).unsqueeze(0), rho_in * temp_rho);
      kbc.coll_f.index({0, Ellipsis}) = (temp_equi + kbc.coll_f.index({-2, Ellipsis}) - f_equi.index({-2, Ellipsis})).squeeze(0).clone().detach();
      solver::incomp_equilibrium(temp_equi, kbc.m1.index({1, Ellipsis}).unsqueeze(0), rho_out * temp_rho);
      kbc.coll_f.index({-1, Ellipsis}) = (temp_equi + kbc.coll_f.index({1, Ellipsis}) - f_equi.index({1, Ellipsis})).squeeze(0).clone().detach();
    }
    kbc.advect();
    if (bc == 1)
    {
      kbc.adve_f.index({Slice(), -1, 4}) = kbc.coll_f.index({Slice(), -1, 2}).clone().detach();
      kbc.adve_f.index({Slice(), -1, 7}) = kbc.coll_f.index({Slice(), -1, 5}).clone().detach();
      kbc.adve_f.index({Slice(), -1, 8}) = kbc.coll_f.index({Slice(), -1, 6}).clone().detach();
      kbc.adve_f.index({Slice(), 0, 2}) = kbc.coll_f.index({Slice(), 0, 4}).clone().detach();
      kbc.adve_f.index({Slice(), 0, 5}) = kbc.coll_f.index({Slice(), 0, 7}).clone().detach();
      kbc.adve_f.index({Slice(), 0, 6}) = kbc.coll_f.index({Slice(), 0, 8}).clone().detach();
    }
    kbc.m0 = kbc.adve_f.sum(-1).detach().clone();
    kbc.m1 = (torch::matmul(kbc.adve_f, c.transpose(0, 1)) / kbc.m0.unsqueeze(-1)).detach().clone();
  }
  copy_out(kbc.m0, m0);
  copy_out(kbc.m1, m1);
  copy_out(kbc.adve_f, adve);
  REF_CATCH
}

// kbc::eval_equilibrium(adve_f) from m0 {R,C}, m1 {R,C,2} (src/ulbm.cpp:246-262; the drivers' initial state)
int ref_kbc_equilibrium(int R, int C, const double* m0, const double* m1, double* feq)
{
  REF_TRY
  ulbm::d2q9::kbc kbc{R, C, 1.0};
  kbc.m0 = wrap(m0, {R, C});
  kbc.m1 = wrap(m1, {R, C, 2});
  torch::Tensor out = torch::zeros({R, C, 9});
  kbc.eval_equilibrium(out);
  copy_out(out, feq);
  REF_CATCH
}

}  // extern "C"
