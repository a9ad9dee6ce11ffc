// TEST INFRASTRUCTURE — the product's collision arithmetic (csrc/lbm_device.cuh) compiled for the HOST, so that the CPU
// suite can hold it against the oracle without a GPU: the functions are __host__ __device__, the same source the
// kernels inline.  (Host code has no fused multiply-add: results differ from the device's in the last bits only.)
#include "../../lattice-boltzmann-method_b200/csrc/lbm_device.cuh"

extern "C"
{

// kbc::collide() of N nodes in place; given = 0: m0, u are the populations' own moments (computed here like
// kbc_collide does itself), given = 1: the caller's m0 {N}, u {N,2}
void host_kbc_collide(double* f, const double* m0, const double* u, int given, long N, double s2)
{
  for (long n = 0; n < N; n++)
  {
    double v[9];
    for (int q = 0; q < 9; q++) v[q] = f[9 * n + q];
    double rho = 0.0, ux = 0.0, uy = 0.0;
    if (given)
    {
      rho = m0[n]; ux = u[2 * n]; uy = u[2 * n + 1];
    }
    lbm::kbc_collide(v, s2, 1.0 / s2, rho, ux, uy, given != 0);
    for (int q = 0; q < 9; q++) f[9 * n + q] = v[q];
  }
}

}  // extern "C"
