// lbm_ops.cu — granular operators, one-to-one with namespace solver (src/solver.hpp:11-36) and
// class differential (src/differential.hpp:48-51), on buffers in the reference's AoS layout.
// They exist for unit parity and for callers that keep the reference's op-by-op loop; the fused
// path is lbm_step.  Each call uploads, runs one CUDA kernel and downloads.
#include <vector>

#include "lbm_internal.hpp"

namespace lbm
{
static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

__global__ void k_op_rho(const double* __restrict__ f, long long N, double* __restrict__ rho)
{
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double v[9];
#pragma unroll
  for (int q = 0; q < 9; q++) v[q] = f[n * 9 + q];
  double r, jx, jy;
  moments(v, r, jx, jy);
  rho[n] = r;
}

__global__ void k_op_u(const double* __restrict__ f, const double* __restrict__ rho, long long N, double* __restrict__ u)
{
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double v[9];
#pragma unroll
  for (int q = 0; q < 9; q++) v[q] = f[n * 9 + q];
  double r, jx, jy;
  moments(v, r, jx, jy);
  if (rho)
  {
    jx /= rho[n];
    jy /= rho[n];
  }
  u[2 * n] = jx;
  u[2 * n + 1] = jy;
}

__global__ void k_op_eq(const double* __restrict__ u, const double* __restrict__ rho, long long N, int incompressible,
                        double* __restrict__ feq)
{
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double ux = u[2 * n], uy = u[2 * n + 1], r = rho[n];
  const double uu = ux * ux + uy * uy;
#pragma unroll
  for (int q = 0; q < 9; q++) feq[n * 9 + q] = incompressible ? feq_incomp(q, r, ux, uy) : feq_comp(q, r, ux, uy, uu);
}

__global__ void k_op_collision(const double* __restrict__ f, const double* __restrict__ feq, double omega, long long N9,
                               double* __restrict__ out)
{
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N9) return;
  out[n] = (1.0 - omega) * f[n] + omega * feq[n];
}

// solver::advect: g(x + c_q, q) = f(x, q), periodic in both axes (src/solver.cpp:76-131)
__global__ void k_op_advect(const double* __restrict__ f, int X, int Y, double* __restrict__ g)
{
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long long)X * Y) return;
  const int x = (int)(n / Y), y = (int)(n % Y);
#pragma unroll
  for (int q = 0; q < 9; q++)
  {
    int xs = x - CX(q), ys = y - CY(q);
    if (xs < 0) xs += X;
    if (xs >= X) xs -= X;
    if (ys < 0) ys += Y;
    if (ys >= Y) ys -= Y;
    g[n * 9 + q] = f[((long long)xs * Y + ys) * 9 + q];
  }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// 5x5 isotropic differences with replicate padding (src/differential.hpp:9-40, src/differential.cpp:3-39)
__global__ void k_op_diff5(const double* __restrict__ psi, int R, int C, double* __restrict__ dx, double* __restrict__ dy)
{
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long long)R * C) return;
  const int i = (int)(n / C), j = (int)(n % C);
  const double xi[3][3] = {{0.0, 960.0, 84.0}, {960.0, 448.0, 32.0}, {84.0, 32.0, 1.0}};  // |offset| indexed
  double sx = 0.0, sy = 0.0;
#pragma unroll
  for (int a = -2; a <= 2; a++)
#pragma unroll
    for (int b = -2; b <= 2; b++)
    {
      const double v = psi[(long long)clampi(i + a, 0, R - 1) * C + clampi(j + b, 0, C - 1)];
      const double w = (1.0 / 5040.0) * xi[a < 0 ? -a : a][b < 0 ? -b : b];
      sx += (w * (double)a) * v;
      sy += (w * (double)b) * v;
    }
  dx[n] = sx;
  dy[n] = sy;
}

// the RK driver's 3x3 operator (test/rk_static_droplet_test.cpp:52-62): "x" runs along axis 1
__global__ void k_op_diff3(const double* __restrict__ psi, int R, int C, double* __restrict__ dx, double* __restrict__ dy)
{
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long long)R * C) return;
  const int i = (int)(n / C), j = (int)(n % C);
  double sx = 0.0, sy = 0.0;
#pragma unroll
  for (int a = -1; a <= 1; a++)
#pragma unroll
    for (int b = -1; b <= 1; b++)
    {
      const double v = psi[(long long)clampi(i + a, 0, R - 1) * C + clampi(j + b, 0, C - 1)];
      const double wa = a == 0 ? 1.0 / 9.0 : 1.0 / 36.0, wb = b == 0 ? 1.0 / 9.0 : 1.0 / 36.0;
      sx += (3.0 * (wa * (double)b)) * v;  // weight by the row offset class, sign by the column offset
      sy += (3.0 * (wb * (double)a)) * v;
    }
  dx[n] = sx;
  dy[n] = sy;
}

struct DevBuf
{
  double* p = nullptr;
  ~DevBuf() { cudaFree(p); }
};

static int need_device()
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
  {
    set_error("no CUDA device (this library has no CPU fallback)");
    return LBM_ERR_CUDA;
  }
  return LBM_OK;
}

static int up(DevBuf& b, const double* h, size_t n)
{
  LBM_CUDA(cudaMalloc(&b.p, n * sizeof(double)));
  if (h) LBM_CUDA(cudaMemcpy(b.p, h, n * sizeof(double), cudaMemcpyHostToDevice));
  return LBM_OK;
}

static int down(double* h, const DevBuf& b, size_t n)
{
  LBM_CUDA(cudaGetLastError());
  LBM_CUDA(cudaMemcpy(h, b.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  return LBM_OK;
}
}  // namespace lbm

using namespace lbm;

#define LBM_OP_PROLOGUE(cond)                                              \
  if (!(cond)) { set_error("%s: bad argument", __func__); return LBM_ERR_INVALID; } \
  LBM_TRY(need_device());                                                  \
  const long long N = (long long)X * Y;

extern "C"
{

int lbm_calc_rho(const double* f, int X, int Y, double* rho)
{
  LBM_OP_PROLOGUE(f && rho && X > 0 && Y > 0)
  DevBuf a, o;
  LBM_TRY(up(a, f, N * 9));
  LBM_TRY(up(o, nullptr, N));
  k_op_rho<<<cdiv(N, 256), 256>>>(a.p, N, o.p);
  return down(rho, o, N);
}

int lbm_calc_u(const double* f, const double* rho, int X, int Y, double* u)
{
  LBM_OP_PROLOGUE(f && rho && u && X > 0 && Y > 0)
  DevBuf a, r, o;
  LBM_TRY(up(a, f, N * 9));
  LBM_TRY(up(r, rho, N));
  LBM_TRY(up(o, nullptr, N * 2));
  k_op_u<<<cdiv(N, 256), 256>>>(a.p, r.p, N, o.p);
  return down(u, o, N * 2);
}

int lbm_calc_incomp_u(const double* f, int X, int Y, double* u)
{
  LBM_OP_PROLOGUE(f && u && X > 0 && Y > 0)
  DevBuf a, o;
  LBM_TRY(up(a, f, N * 9));
  LBM_TRY(up(o, nullptr, N * 2));
  k_op_u<<<cdiv(N, 256), 256>>>(a.p, nullptr, N, o.p);
  return down(u, o, N * 2);
}

static int eq_common(const double* u, const double* rho, int X, int Y, double* feq, int inc)
{
  LBM_OP_PROLOGUE(u && rho && feq && X > 0 && Y > 0)
  DevBuf a, r, o;
  LBM_TRY(up(a, u, N * 2));
  LBM_TRY(up(r, rho, N));
  LBM_TRY(up(o, nullptr, N * 9));
  k_op_eq<<<cdiv(N, 256), 256>>>(a.p, r.p, N, inc, o.p);
  return down(feq, o, N * 9);
}

int lbm_equilibrium(const double* u, const double* rho, int X, int Y, double* feq) { return eq_common(u, rho, X, Y, feq, 0); }
int lbm_incomp_equilibrium(const double* u, const double* rho, int X, int Y, double* feq) { return eq_common(u, rho, X, Y, feq, 1); }

int lbm_collision(const double* f, const double* feq, double omega, int X, int Y, double* fcoll)
{
  LBM_OP_PROLOGUE(f && feq && fcoll && X > 0 && Y > 0)
  DevBuf a, b, o;
  LBM_TRY(up(a, f, N * 9));
  LBM_TRY(up(b, feq, N * 9));
  LBM_TRY(up(o, nullptr, N * 9));
  k_op_collision<<<cdiv(N * 9, 256), 256>>>(a.p, b.p, omega, N * 9, o.p);
  return down(fcoll, o, N * 9);
}

int lbm_advect(const double* f, int X, int Y, double* g)
{
  LBM_OP_PROLOGUE(f && g && X > 0 && Y > 0)
  DevBuf a, o;
  LBM_TRY(up(a, f, N * 9));
  LBM_TRY(up(o, nullptr, N * 9));
  k_op_advect<<<cdiv(N, 256), 256>>>(a.p, X, Y, o.p);
  return down(g, o, N * 9);
}

static int diff_common(const double* psi, int X, int Y, double* dx, double* dy, int three)
{
  LBM_OP_PROLOGUE(psi && dx && dy && X > 0 && Y > 0)
  DevBuf a, ox, oy;
  LBM_TRY(up(a, psi, N));
  LBM_TRY(up(ox, nullptr, N));
  LBM_TRY(up(oy, nullptr, N));
  if (three) k_op_diff3<<<cdiv(N, 256), 256>>>(a.p, X, Y, ox.p, oy.p);
  else k_op_diff5<<<cdiv(N, 256), 256>>>(a.p, X, Y, ox.p, oy.p);
  LBM_TRY(down(dx, ox, N));
  return down(dy, oy, N);
}

int lbm_differential(const double* psi, int R, int C, double* dx, double* dy) { return diff_common(psi, R, C, dx, dy, 0); }
int lbm_differential3(const double* psi, int R, int C, double* dx, double* dy) { return diff_common(psi, R, C, dx, dy, 1); }

}  // extern "C"
