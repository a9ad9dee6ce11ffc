// lbm_params.cpp — the parameters.toml surface behind the C ABI (host-only scalar code).
//
// Mirrors, key for key and in the same arithmetic order:
//   params::flow / lattice / simulation        src/params.cpp:7-120
//   colour constants, phi, eta                 src/colour.cpp:11-64
//   driver-local `domain` + [general]          test/mrtcg_rayleigh_taylor.cpp:103-117,360-362
//   boundary file `[name] x=[..] y=[..]`       src/ibm.cpp:78-102
// A missing key yields LBM_ERR_CONFIG with the reference's message
// "<key> not defined in parameters file" (it throws std::runtime_error there).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/lbm_b200.h"
#include "toml_lite.hpp"

namespace lbm
{
void set_error(const char* fmt, ...);
}

namespace
{
using toml_lite::value;

struct config_error
{
  std::string msg;
};

const value* sub(const value* t, const char* key) { return t ? t->find(key) : nullptr; }

double need_double(const value* tbl, const char* key, const char* sep = " ")
{
  const value* v = sub(tbl, key);
  if (v)
  {
    auto d = v->as_double();
    if (d) return *d;
  }
  throw config_error{std::string(key) + sep + "not defined in parameters file"};
}

int need_int(const value* tbl, const char* key)
{
  const value* v = sub(tbl, key);
  if (v)
  {
    auto i = v->as_int();
    if (i) return (int)*i;
  }
  // test/mrtcg_rayleigh_taylor.cpp:31 concatenates without a space
  throw config_error{std::string(key) + "not defined in parameters file"};
}

std::string need_string(const value* tbl, const char* key, const char* sep = " ")
{
  const value* v = sub(tbl, key);
  if (v)
  {
    auto s = v->as_string();
    if (s) return *s;
  }
  throw config_error{std::string(key) + sep + "not defined in parameters file"};
}

void copy_str(char* dst, size_t cap, const std::string& s)
{
  std::snprintf(dst, cap, "%s", s.c_str());
}

template <typename F>
int guarded(F&& body)
{
  try
  {
    body();
    return LBM_OK;
  }
  catch (const config_error& e)
  {
    lbm::set_error("%s", e.msg.c_str());
    return LBM_ERR_CONFIG;
  }
  catch (const toml_lite::parse_error& e)
  {
    lbm::set_error("Parsing failed: %s", e.what());
    return LBM_ERR_CONFIG;
  }
  catch (const std::exception& e)
  {
    lbm::set_error("%s", e.what());
    return LBM_ERR_INVALID;
  }
}
}  // namespace

extern "C"
{

int lbm_params_from_toml(const char* path, int require_simulation, lbm_params* out)
{
  if (!path || !out) { lbm::set_error("lbm_params_from_toml: null argument"); return LBM_ERR_INVALID; }
  return guarded([&] {
    std::memset(out, 0, sizeof(*out));
    auto root = toml_lite::parse_file(path);
    const value* flow = root->find("flow");
    const value* lat = root->find("lattice");
    // params::flow::flow (src/params.cpp:7-29), same key order
    out->flow_rho_0 = need_double(flow, "initial_density");
    out->flow_nu = need_double(flow, "kinematic_viscosity");
    out->flow_u = need_double(flow, "characteristic_velocity");
    out->flow_l = need_double(flow, "characteristic_length");
    out->flow_Re = out->flow_u * out->flow_l / out->flow_nu;
    // params::lattice::lattice (src/params.cpp:31-66)
    const double cs2 = 1.0 / 3.0;
    out->tau = need_double(lat, "relaxation_time");
    out->dx = need_double(lat, "lattice_spacing");
    const double x_mult = need_double(lat, "x_multiplier");
    const double y_mult = need_double(lat, "y_multiplier");
    // characteristic length -> nearest odd integer
    if ((int)std::ceil(out->flow_l / out->dx) % 2 != 0) out->l = (int)std::ceil(out->flow_l / out->dx);
    else out->l = (int)std::floor(out->flow_l / out->dx);
    out->omega = 1.0 / out->tau;
    out->Re = out->flow_Re;
    out->nu = cs2 * (out->tau - 0.5);
    out->u = out->flow_Re * out->nu / out->l;
    out->dt = cs2 * (out->tau - 0.5) * (out->dx * out->dx) / out->flow_nu;
    out->T = (int)std::ceil(1.0 / out->dt);
    out->X = (int)std::ceil(out->l * x_mult);
    out->Y = (int)std::ceil(out->l * y_mult);
    // params::simulation::simulation (src/params.cpp:95-112)
    const value* sim = root->find("simulation");
    if (require_simulation || sim)
    {
      try
      {
        out->stop_time = need_double(sim, "stop_time");
        out->snapshot_period = need_double(sim, "snapshot_period");
        copy_str(out->file_prefix, sizeof(out->file_prefix), need_string(sim, "file_prefix"));
        out->total_steps = (int)std::ceil(out->stop_time * out->T);
        out->snapshot_steps = (int)std::ceil(out->snapshot_period * out->T);
        out->total_snapshots = (int)std::ceil((out->total_steps + 0.0) / out->snapshot_steps);
        out->has_simulation = 1;
      }
      catch (const config_error&)
      {
        if (require_simulation) throw;
        out->has_simulation = 0;
      }
    }
  });
}

int lbm_colour_from_toml(const char* path, const char* table, lbm_colour* out)
{
  if (!path || !table || !out) { lbm::set_error("lbm_colour_from_toml: null argument"); return LBM_ERR_INVALID; }
  return guarded([&] {
    std::memset(out, 0, sizeof(*out));
    auto root = toml_lite::parse_file(path);
    const value* t = root->find(table);
    // colour::colour initialiser list order (src/colour.cpp:12-20); try_double glues key and message (:45)
    out->rho_0 = need_double(t, "initial_density", "");
    out->alpha = need_double(t, "alpha", "");
    out->A = need_double(t, "interfacial_tension_control", "");
    out->nu = need_double(t, "kinematic_viscosity", "");
    out->mu = out->nu * out->rho_0;
    out->beta = need_double(t, "interface_thickness_control", "");
    out->cs2 = 3.0 * (1.0 - out->alpha) / 5.0;      // init_cs2 (:37)
    out->ics2 = 1.0 / out->cs2;
    out->rlx = 1.0 / (0.5 + out->nu / out->cs2);    // init_rlx_param (:38-39)
    const double a = 0.2 * (1.0 - out->alpha), b = 0.05 * (1.0 - out->alpha);  // init_phi (:56-64)
    out->phi[0] = out->alpha;
    for (int q = 1; q < 5; q++) out->phi[q] = a;
    for (int q = 5; q < 9; q++) out->phi[q] = b;
    for (int q = 0; q < 9; q++)  // init_eta (:49-54)
    {
      const double ee = q == 0 ? 0.0 : (q < 5 ? 1.0 : 2.0);
      out->eta[q] = 1.0 + 0.5 * (3.0 * out->cs2 - 1.0) * (3.0 * ee - 4.0);
    }
  });
}

int lbm_two_phase_from_toml(const char* path, int require_general, lbm_two_phase_params* out)
{
  if (!path || !out) { lbm::set_error("lbm_two_phase_from_toml: null argument"); return LBM_ERR_INVALID; }
  return guarded([&] {
    std::memset(out, 0, sizeof(*out));
    auto root = toml_lite::parse_file(path);
    const value* gen = root->find("general");
    if (require_general || gen)
    {
      try
      {
        out->sigma = need_double(gen, "sigma", "");
        out->gravity_magnitude = need_double(gen, "gravity_magnitude", "");
        copy_str(out->name, sizeof(out->name), need_string(gen, "name", ""));
        out->has_general = 1;
      }
      catch (const config_error&)
      {
        if (require_general) throw;
      }
    }
    const value* dom = root->find("domain");
    out->rows = need_int(dom, "rows");
    out->columns = need_int(dom, "columns");
    out->time_steps = need_int(dom, "time_steps");
    out->nr_snapshots = need_int(dom, "nr_snapshots");
    if (out->nr_snapshots <= 0) throw config_error{"nr_snapshots must be positive"};
    out->period_snapshots = (int)(out->time_steps / out->nr_snapshots);
  });
}

int lbm_markers_from_toml(const char* path, const char* name, double* xs, double* ys, int* n)
{
  if (!path || !name || !n) { lbm::set_error("lbm_markers_from_toml: null argument"); return LBM_ERR_INVALID; }
  return guarded([&] {
    auto root = toml_lite::parse_file(path);
    const value* t = root->find(name);
    const value* x = sub(t, "x");
    const value* y = sub(t, "y");
    if (!x || !y || !x->is_array() || !y->is_array() || x->arr.size() != y->arr.size())
      throw config_error{std::string("boundary table [") + name + "] needs arrays x and y of equal length"};
    const int cap = *n;
    *n = (int)x->arr.size();
    if (!xs || !ys) return;
    if (cap < *n) throw config_error{"marker buffers too small"};
    for (int i = 0; i < *n; i++)
    {
      auto xv = x->arr[i]->as_double();
      auto yv = y->arr[i]->as_double();
      if (!xv) throw config_error{"Cannot parse x coordinate"};  // src/ibm.cpp:90
      if (!yv) throw config_error{"Cannot parse y coordinate"};  // src/ibm.cpp:93
      xs[i] = *xv;
      ys[i] = *yv;
    }
  });
}

}  // extern "C"
