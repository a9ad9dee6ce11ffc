// common.hpp — shared helpers of the host drivers.  The drivers mirror the reference's test/*.cpp
// (one main() per case: parse TOML -> allocate -> time loop -> save snapshots) but call the C ABI of
// include/lbm_b200.h instead of libtorch.  Snapshots are written as NumPy .npy files with the same
// shapes the reference gives its torch::save'd tensors ({X,Y,T}, {X,Y,9,T}) and in the reference's own
// on-disk format (lbm_save_pt: the TorchScript archive torch::save writes), under the reference's file
// names; LBM_SNAPSHOT_FORMAT=npy writes NumPy .npy files instead.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../include/lbm_b200.h"

namespace drv
{

inline void check(int status, const char* what)
{
  if (status != LBM_OK)
  {
    std::cerr << what << " failed (" << status << "): " << lbm_last_error() << "\n";
    std::exit(status == LBM_ERR_CONFIG ? 1 : 2);
  }
}
#define DRV_CHECK(call) drv::check((call), #call)

// fp64 C-order array -> .npy (format 1.0)
inline void save_npy(const std::string& path, const std::vector<double>& data, const std::vector<long>& shape)
{
  std::string dict = "{'descr': '<f8', 'fortran_order': False, 'shape': (";
  for (size_t i = 0; i < shape.size(); i++) dict += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
  dict += "), }";
  size_t total = 10 + dict.size() + 1;
  size_t pad = (64 - total % 64) % 64;
  dict += std::string(pad, ' ') + "\n";
  std::ofstream out(path, std::ios::binary);
  if (!out) { std::cerr << "cannot write " << path << "\n"; std::exit(2); }
  const char magic[] = {'\x93', 'N', 'U', 'M', 'P', 'Y', 1, 0};
  out.write(magic, 8);
  const uint16_t hl = (uint16_t)dict.size();
  out.write(reinterpret_cast<const char*>(&hl), 2);
  out.write(dict.data(), dict.size());
  out.write(reinterpret_cast<const char*>(data.data()), data.size() * sizeof(double));
}

// torch::save(tensor, path) of the reference drivers; path ends in ".pt"
inline void save_array(const std::string& path, const std::vector<double>& data, const std::vector<long>& shape)
{
  const char* fmt = std::getenv("LBM_SNAPSHOT_FORMAT");
  if (fmt && std::string(fmt) == "npy")
  {
    save_npy(path.substr(0, path.size() - 3) + ".npy", data, shape);
    return;
  }
  std::vector<long long> sh(shape.begin(), shape.end());
  check(lbm_save_pt(path.c_str(), data.data(), sh.data(), (int)sh.size()), "lbm_save_pt");
}

// {X,Y,T} stack the reference fills with `ux.index({Ellipsis,i}) = ...`
struct Series
{
  long X, Y, C, T;  // C components per node (1 for scalars, 9 for populations)
  std::vector<double> a;
  Series(long X_, long Y_, long T_, long C_ = 1) : X(X_), Y(Y_), C(C_), T(T_), a((size_t)X_ * Y_ * C_ * T_, 0.0) {}
  // src: {X,Y,stride} taking component `comp` (or all C when C > 1)
  void put(long t, const std::vector<double>& src, long stride, long comp, double scale = 1.0)
  {
    if (t >= T) return;
    for (long n = 0; n < X * Y; n++)
      for (long c = 0; c < C; c++) a[((size_t)n * C + c) * T + t] = scale * src[(size_t)n * stride + (C > 1 ? c : comp)];
  }
  void save(const std::string& path) const
  {
    if (C > 1) save_array(path, a, {X, Y, C, T});
    else save_array(path, a, {X, Y, T});
  }
};

// The reference drivers ask before a long run (utils::continue_execution).  Same question and answers on the terminal;
// here the answer is read line-wise (so "yes" / "no" work and trailing input does not leak into the next question), end of
// input means no, and LBM_ASSUME_YES=1 answers for unattended runs (batch queues, the tests).
inline bool continue_execution()
{
  if (const char* yes = std::getenv("LBM_ASSUME_YES"); yes && yes[0] == '1') return true;
  for (std::string line;;)
  {
    std::cout << "\nDo you want to continue (y/n)? " << std::flush;
    if (!std::getline(std::cin, line)) return false;
    const auto first = line.find_first_not_of(" \t\r");
    const char c = first == std::string::npos ? '\0' : line[first];
    if (c == 'y' || c == 'Y') return true;
    if (c == 'n' || c == 'N') return false;
    std::cout << "Invalid input. Please enter 'y' or 'n'." << std::endl;
  }
}

inline void print_params(const lbm_params& p)  // operator<< of params::flow / lattice / simulation (src/params.cpp:68-128)
{
  std::cout << "Flow parameters:\nnu=" << p.flow_nu << " m2/s\nu=" << p.flow_u << " m/s\nl=" << p.flow_l << " m\nrho_0=" << p.flow_rho_0
            << " kg/m3\nRe=" << p.flow_Re << "\n\n";
  std::cout << "Lattice parameters:\nRe=" << p.Re << "\ntau=" << p.tau << "\nomega=" << p.omega << "\ndx=" << p.dx << " m\nl=" << p.l
            << "\nnu=" << p.nu << "\nu=" << p.u << "\ndt=" << p.dt << "s\nT=" << p.T << "\nX=" << p.X << "\nY=" << p.Y << "\n\n";
  if (p.has_simulation)
    std::cout << "Simulation parameters:\nstop time: " << p.stop_time << " s (" << p.total_steps << " steps)\nsaving results each "
              << p.snapshot_period << " s (" << p.snapshot_steps << " steps)\nfor a total of " << p.total_snapshots << " snapshots\n\n";
}

}  // namespace drv
