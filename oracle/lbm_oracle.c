/* TEST INFRASTRUCTURE — CPU restatement of the reference's D2Q9 time step.
 * See lbm_oracle.h for scope, layout and parity status (PINNED against oracle/_ref).
 * Written for clarity and fidelity to the reference's evaluation order, not for speed
 * (the OpenMP pragmas only make the cpu_baseline "port" leg of bench.py use the host cores).
 * All citations are relative to /root/reference. */
#include "lbm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* src/solver.cpp:12-21 */
static const double W9[9] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0,
                             1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
static const double CXD[9] = {0.0, 1.0, 0.0, -1.0, 0.0, 1.0, -1.0, -1.0, 1.0};
static const double CYD[9] = {0.0, 0.0, 1.0, 0.0, -1.0, 1.0, 1.0, -1.0, -1.0};
static const int CXI[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int CYI[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const int OPP[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};

#define IDX(x, y, q) ((((size_t)(x)) * Y + (y)) * 9 + (q))
#define N2(x, y) (((size_t)(x)) * Y + (y))

static double* dalloc(size_t n)
{
  double* p = (double*)calloc(n ? n : 1, sizeof(double));
  if (!p) abort();
  return p;
}

/* host threads of the port's OpenMP loops (1 without OpenMP); n > 0 sets the count first.  bench.py's CPU arms state the
 * count explicitly: a launcher's OMP_NUM_THREADS=1 (torch.distributed.run exports it) must not shrink the baseline. */
#ifdef _OPENMP
#include <omp.h>
#endif
int orc_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

void orc_constants(double* w9, double* c18)
{
  for (int q = 0; q < 9; q++)
  {
    w9[q] = W9[q];
    c18[q] = CXD[q];
    c18[9 + q] = CYD[q];
  }
}

/* ------------------------------------------------------------------ granular ops */

/* src/solver.cpp:23-26 */
void orc_calc_rho(const double* f, int X, int Y, double* rho)
{
#pragma omp parallel for
  for (long n = 0; n < (long)X * Y; n++)
  {
    double s = 0.0;
    for (int q = 0; q < 9; q++) s += f[n * 9 + q];
    rho[n] = s;
  }
}

/* src/solver.cpp:28-31 */
void orc_calc_incomp_u(const double* f, int X, int Y, double* u)
{
#pragma omp parallel for
  for (long n = 0; n < (long)X * Y; n++)
  {
    double sx = 0.0, sy = 0.0;
    for (int q = 0; q < 9; q++)
    {
      sx += f[n * 9 + q] * CXD[q];
      sy += f[n * 9 + q] * CYD[q];
    }
    u[n * 2 + 0] = sx;
    u[n * 2 + 1] = sy;
  }
}

/* src/solver.cpp:34-37 */
void orc_calc_u(const double* f, const double* rho, int X, int Y, double* u)
{
  orc_calc_incomp_u(f, X, Y, u);
#pragma omp parallel for
  for (long n = 0; n < (long)X * Y; n++)
  {
    u[n * 2 + 0] /= rho[n];
    u[n * 2 + 1] /= rho[n];
  }
}

/* src/solver.cpp:51-62 */
static void eq_node(double ux, double uy, double rho, double* feq)
{
  double uu = ux * ux + uy * uy;
  for (int q = 0; q < 9; q++)
  {
    double cu = ux * CXD[q] + uy * CYD[q];
    double A = 1.0 + 3.0 * cu + 4.5 * (cu * cu) - 1.5 * uu;
    feq[q] = (rho * A) * W9[q];
  }
}

/* src/solver.cpp:39-49 */
static void inc_eq_node(double ux, double uy, double rho, double* feq)
{
  for (int q = 0; q < 9; q++)
  {
    double cu = ux * CXD[q] + uy * CYD[q];
    double A = rho + 3.0 * cu;
    feq[q] = A * W9[q];
  }
}

void orc_equilibrium(const double* u, const double* rho, int X, int Y, double* feq)
{
#pragma omp parallel for
  for (long n = 0; n < (long)X * Y; n++) eq_node(u[n * 2], u[n * 2 + 1], rho[n], feq + n * 9);
}

void orc_incomp_equilibrium(const double* u, const double* rho, int X, int Y, double* feq)
{
#pragma omp parallel for
  for (long n = 0; n < (long)X * Y; n++) inc_eq_node(u[n * 2], u[n * 2 + 1], rho[n], feq + n * 9);
}

/* src/solver.cpp:65-74 */
void orc_collision(const double* f, const double* feq, double omega, int X, int Y, double* fcoll)
{
#pragma omp parallel for
  for (long n = 0; n < (long)X * Y * 9; n++) fcoll[n] = (1.0 - omega) * f[n] + omega * feq[n];
}

/* src/solver.cpp:76-131 — fully periodic streaming: g(x+c_q, q) = f(x, q) */
void orc_advect(const double* f, int X, int Y, double* g)
{
#pragma omp parallel for
  for (int x = 0; x < X; x++)
    for (int y = 0; y < Y; y++)
      for (int q = 0; q < 9; q++)
      {
        int xs = x - CXI[q];
        int ys = y - CYI[q];
        if (xs < 0) xs += X;
        if (xs >= X) xs -= X;
        if (ys < 0) ys += Y;
        if (ys >= Y) ys -= Y;
        g[IDX(x, y, q)] = f[IDX(xs, ys, q)];
      }
}

/* ------------------------------------------------------------------ finite differences */

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* src/differential.hpp:9-40: weight xi[a][b]/5040 times the offset along the axis */
static const double XI5[5][5] = {{1.0, 32.0, 84.0, 32.0, 1.0},
                                 {32.0, 448.0, 960.0, 448.0, 32.0},
                                 {84.0, 960.0, 0.0, 960.0, 84.0},
                                 {32.0, 448.0, 960.0, 448.0, 32.0},
                                 {1.0, 32.0, 84.0, 32.0, 1.0}};

void orc_diff5(const double* psi, int R, int C, double* dx, double* dy)
{
#pragma omp parallel for
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++)
    {
      double sx = 0.0, sy = 0.0;
      for (int a = 0; a < 5; a++)
        for (int b = 0; b < 5; b++)
        {
          /* replicate padding (src/differential.cpp:8-9) */
          double v = psi[(size_t)clampi(i + a - 2, 0, R - 1) * C + clampi(j + b - 2, 0, C - 1)];
          double w = (1.0 / 5040.0) * XI5[a][b];
          /* kernel_partial_x = -{2,1,0,-1,-2} down the rows (src/differential.hpp:31-40) */
          sx += (w * (double)(a - 2)) * v;
          /* kernel_partial_y = {-2,-1,0,1,2} along the columns (src/differential.hpp:20-29) */
          sy += (w * (double)(b - 2)) * v;
        }
      dx[(size_t)i * C + j] = sx;
      dy[(size_t)i * C + j] = sy;
    }
}

/* test/rk_static_droplet_test.cpp:52-62 */
void orc_diff3(const double* psi, int R, int C, double* dx, double* dy)
{
  static const double KX[3][3] = {{-1.0 / 36.0, 0.0, 1.0 / 36.0}, {-1.0 / 9.0, 0.0, 1.0 / 9.0}, {-1.0 / 36.0, 0.0, 1.0 / 36.0}};
  static const double KY[3][3] = {{1.0 / 36.0, 1.0 / 9.0, 1.0 / 36.0}, {0.0, 0.0, 0.0}, {-1.0 / 36.0, -1.0 / 9.0, -1.0 / 36.0}};
#pragma omp parallel for
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++)
    {
      double sx = 0.0, sy = 0.0;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
        {
          double v = psi[(size_t)clampi(i + a - 1, 0, R - 1) * C + clampi(j + b - 1, 0, C - 1)];
          sx += (3.0 * KX[a][b]) * v;
          sy += (-3.0 * KY[a][b]) * v;
        }
      dx[(size_t)i * C + j] = sx;
      dy[(size_t)i * C + j] = sy;
    }
}

/* ------------------------------------------------------------------ host parameters */

/* src/params.cpp:7-66 */
void orc_params_lattice(const double* in, double* out)
{
  const double fp_nu = in[1], fp_u = in[2], fp_l = in[3];
  const double tau = in[4], dx = in[5], x_mult = in[6], y_mult = in[7];
  const double cs2 = 1.0 / 3.0;
  const double Re = fp_u * fp_l / fp_nu;
  int l;
  if ((int)ceil(fp_l / dx) % 2 != 0) l = (int)ceil(fp_l / dx);
  else l = (int)floor(fp_l / dx);
  const double omega = 1.0 / tau;
  const double nu = cs2 * (tau - 0.5);
  const double u = Re * nu / l;
  const double dt = cs2 * (tau - 0.5) * (dx * dx) / fp_nu;
  const int T = (int)ceil(1.0 / dt);
  const int Xn = (int)ceil(l * x_mult);
  const int Yn = (int)ceil(l * y_mult);
  out[0] = Re; out[1] = omega; out[2] = nu; out[3] = l; out[4] = dt; out[5] = T; out[6] = u;
  out[7] = Xn; out[8] = Yn;
}

/* src/params.cpp:95-112 */
void orc_params_simulation(double stop_time, double snapshot_period, int T, double* out3)
{
  int total_steps = (int)ceil(stop_time * T);
  int snapshot_steps = (int)ceil(snapshot_period * T);
  int total_snapshots = (int)ceil((total_steps + 0.0) / snapshot_steps);
  out3[0] = total_steps; out3[1] = snapshot_steps; out3[2] = total_snapshots;
}

/* src/colour.cpp:37-64 */
static void colour_derive(double alpha, double nu, double* cs2, double* rlx, double* phi, double* eta)
{
  *cs2 = 3.0 * (1.0 - alpha) / 5.0;
  *rlx = 1.0 / (0.5 + nu / (*cs2));
  double a = 0.2 * (1.0 - alpha);
  double b = 0.05 * (1.0 - alpha);
  phi[0] = alpha;
  for (int q = 1; q < 5; q++) phi[q] = a;
  for (int q = 5; q < 9; q++) phi[q] = b;
  for (int q = 0; q < 9; q++)
  {
    double ee = CXD[q] * CXD[q] + CYD[q] * CYD[q];
    eta[q] = 1.0 + 0.5 * (3.0 * (*cs2) - 1.0) * (3.0 * ee - 4.0);
  }
}

void orc_colour_params(double rho_0, double alpha, double nu, double* out)
{
  double cs2, rlx;
  colour_derive(alpha, nu, &cs2, &rlx, out + 4, out + 13);
  out[0] = nu * rho_0;
  out[1] = cs2;
  out[2] = 1.0 / cs2;
  out[3] = rlx;
}

/* ------------------------------------------------------------------ shared driver pieces */

/* pressure-periodic rows: test/horizontal_poiseuille_test.cpp:25-45 (incompressible eq.) and
 * test/specular_boundary_test.cpp:24-43 (compressible eq.) */
static void pressure_periodic(double* fcoll, const double* feq, const double* u, int X, int Y,
                              double rho_in, double rho_out, int incompressible)
{
  for (int y = 0; y < Y; y++)
  {
    double t[9];
    /* inlet: row 0 <- outlet row X-2 */
    if (incompressible) inc_eq_node(u[N2(X - 2, y) * 2], u[N2(X - 2, y) * 2 + 1], rho_in * 1.0, t);
    else eq_node(u[N2(X - 2, y) * 2], u[N2(X - 2, y) * 2 + 1], rho_in * 1.0, t);
    for (int q = 0; q < 9; q++) fcoll[IDX(0, y, q)] = (t[q] + fcoll[IDX(X - 2, y, q)]) - feq[IDX(X - 2, y, q)];
    /* outlet: row X-1 <- inlet row 1 */
    if (incompressible) inc_eq_node(u[N2(1, y) * 2], u[N2(1, y) * 2 + 1], rho_out * 1.0, t);
    else eq_node(u[N2(1, y) * 2], u[N2(1, y) * 2 + 1], rho_out * 1.0, t);
    for (int q = 0; q < 9; q++) fcoll[IDX(X - 1, y, q)] = (t[q] + fcoll[IDX(1, y, q)]) - feq[IDX(1, y, q)];
  }
}

/* half-way bounce-back on the first/last column: test/horizontal_poiseuille_test.cpp:146-152 */
static void walls_bounce_back(double* fadv, const double* fcoll, int X, int Y)
{
  for (int x = 0; x < X; x++)
  {
    fadv[IDX(x, Y - 1, 4)] = fcoll[IDX(x, Y - 1, 2)];
    fadv[IDX(x, Y - 1, 7)] = fcoll[IDX(x, Y - 1, 5)];
    fadv[IDX(x, Y - 1, 8)] = fcoll[IDX(x, Y - 1, 6)];
    fadv[IDX(x, 0, 2)] = fcoll[IDX(x, 0, 4)];
    fadv[IDX(x, 0, 5)] = fcoll[IDX(x, 0, 7)];
    fadv[IDX(x, 0, 6)] = fcoll[IDX(x, 0, 8)];
  }
}

/* specular columns: test/specular_boundary_test.cpp:122-128, test/cylinder_test.cpp:157-163 */
static void walls_specular(double* fadv, const double* fcoll, int X, int Y)
{
  for (int x = 0; x < X; x++)
  {
    fadv[IDX(x, Y - 1, 4)] = fcoll[IDX(x, Y - 1, 2)];
    fadv[IDX(x, Y - 1, 7)] = fcoll[IDX(x, Y - 1, 6)];
    fadv[IDX(x, Y - 1, 8)] = fcoll[IDX(x, Y - 1, 5)];
    fadv[IDX(x, 0, 2)] = fcoll[IDX(x, 0, 4)];
    fadv[IDX(x, 0, 5)] = fcoll[IDX(x, 0, 8)];
    fadv[IDX(x, 0, 6)] = fcoll[IDX(x, 0, 7)];
  }
}

/* anti-bounce-back constant: test/cylinder_test.cpp:135, free_stream_test.cpp:108 */
static void abb_const(double uwx, double uwy, double* abb)
{
  double uu = uwx * uwx + uwy * uwy;
  for (int q = 0; q < 9; q++)
  {
    double cu = uwx * CXD[q] + uwy * CYD[q];
    abb[q] = (2.0 + 9.0 * pow(cu, 2.0) - 3.0 * uu) * W9[q];
  }
}

/* ABB rows 0 and X-1, fixed u_w: test/cylinder_test.cpp:135-154 */
static void abb_rows(double* fadv, const double* fcoll, int X, int Y, double uwx, double uwy)
{
  double abb[9];
  abb_const(uwx, uwy, abb);
  const int rows[2] = {0, X - 1};
  for (int k = 0; k < 2; k++)
    for (int y = 0; y < Y; y++)
      for (int q = 1; q < 9; q++) fadv[IDX(rows[k], y, OPP[q])] = -fcoll[IDX(rows[k], y, q)] + abb[q];
}

/* ------------------------------------------------------------------ drivers 10, 13, 14 */

static void channel_step(double* f, double* u, double* rho, int X, int Y, double omega, double rho_in,
                         double rho_out, int incompressible, int specular, const double* Fg)
{
  size_t N = (size_t)X * Y;
  double* feq = dalloc(N * 9);
  double* fcoll = dalloc(N * 9);
  orc_calc_rho(f, X, Y, rho);
  if (incompressible) orc_calc_incomp_u(f, X, Y, u);
  else orc_calc_u(f, rho, X, Y, u);
  if (Fg) /* test/gravity_test.cpp:143 */
    for (size_t n = 0; n < N; n++)
    {
      u[n * 2] += Fg[0];
      u[n * 2 + 1] += Fg[1];
    }
  if (incompressible) orc_incomp_equilibrium(u, rho, X, Y, feq);
  else orc_equilibrium(u, rho, X, Y, feq);
  if (!Fg) orc_collision(f, feq, omega, X, Y, fcoll);
  else
  {
    /* test/gravity_test.cpp:147-160 ; note ics2 = 1/3 and ics4 = 1/9 as named there (:78-79) */
    const double ics2 = 1.0 / 3.0, ics4 = 1.0 / 9.0;
    for (size_t n = 0; n < N; n++)
    {
      double ux = u[n * 2], uy = u[n * 2 + 1];
      double uF = ux * Fg[0] + uy * Fg[1];
      for (int q = 0; q < 9; q++)
      {
        double cu = ux * CXD[q] + uy * CYD[q];
        double cF = Fg[0] * CXD[q] + Fg[1] * CYD[q];
        double S = ((1 - 0.5 * omega) * ((ics2 + ics4 * cu) * cF - ics2 * uF)) * W9[q];
        double ep = -omega * (f[n * 9 + q] - feq[n * 9 + q]);
        fcoll[n * 9 + q] = (f[n * 9 + q] + ep) + S;
      }
    }
  }
  pressure_periodic(fcoll, feq, u, X, Y, rho_in, rho_out, incompressible);
  orc_advect(fcoll, X, Y, f);
  if (specular) walls_specular(f, fcoll, X, Y);
  else walls_bounce_back(f, fcoll, X, Y);
  free(feq);
  free(fcoll);
}

void orc_poiseuille_step(double* f, double* u, double* rho, int X, int Y, double omega, double rho_in,
                         double rho_out)
{
  channel_step(f, u, rho, X, Y, omega, rho_in, rho_out, 1, 0, NULL);
}

void orc_specular_step(double* f, double* u, double* rho, int X, int Y, double omega, double rho_in,
                       double rho_out)
{
  channel_step(f, u, rho, X, Y, omega, rho_in, rho_out, 0, 1, NULL);
}

void orc_gravity_step(double* f, double* u, double* rho, int X, int Y, double omega, double rho_in,
                      double rho_out, const double* Fg)
{
  channel_step(f, u, rho, X, Y, omega, rho_in, rho_out, 1, 0, Fg);
}

/* ------------------------------------------------------------------ driver 12 */

void orc_free_stream_step(double* f, double* u, double* rho, int X, int Y, double omega, double uwx)
{
  size_t N = (size_t)X * Y;
  double* feq = dalloc(N * 9);
  double* fcoll = dalloc(N * 9);
  orc_calc_rho(f, X, Y, rho);
  orc_calc_incomp_u(f, X, Y, u);
  orc_incomp_equilibrium(u, rho, X, Y, feq);
  orc_collision(f, feq, omega, X, Y, fcoll);
  orc_advect(fcoll, X, Y, f);
  abb_rows(f, fcoll, X, Y, uwx, 0.0);
  walls_specular(f, fcoll, X, Y);
  free(feq);
  free(fcoll);
}

/* ------------------------------------------------------------------ driver 19 */

void orc_decompose_step(double* fA, double* uA, double* rhoA, double* fB, double* uB, double* rhoB,
                        int X, int Y, double omega, double rho_in, double rho_out)
{
  size_t N = (size_t)X * Y;
  double* eqA = dalloc(N * 9);
  double* eqB = dalloc(N * 9);
  double* cA = dalloc(N * 9);
  double* cB = dalloc(N * 9);
  /* :141-152 */
  orc_calc_rho(fA, X, Y, rhoA);
  orc_calc_rho(fB, X, Y, rhoB);
  orc_calc_u(fA, rhoA, X, Y, uA);
  orc_calc_u(fB, rhoB, X, Y, uB);
  orc_equilibrium(uA, rhoA, X, Y, eqA);
  orc_equilibrium(uB, rhoB, X, Y, eqB);
  orc_collision(fA, eqA, omega, X, Y, cA);
  orc_collision(fB, eqB, omega, X, Y, cB);
  /* cross-domain pressure BC :50-73 */
  for (int y = 0; y < Y; y++)
  {
    double t[9];
    eq_node(uB[N2(X - 2, y) * 2], uB[N2(X - 2, y) * 2 + 1], rho_in * 1.0, t);
    for (int q = 0; q < 9; q++) cA[IDX(0, y, q)] = (t[q] + cB[IDX(X - 2, y, q)]) - eqB[IDX(X - 2, y, q)];
    eq_node(uA[N2(1, y) * 2], uA[N2(1, y) * 2 + 1], rho_out * 1.0, t);
    for (int q = 0; q < 9; q++) cB[IDX(X - 1, y, q)] = (t[q] + cA[IDX(1, y, q)]) - eqA[IDX(1, y, q)];
  }
  orc_advect(cA, X, Y, fA);
  orc_advect(cB, X, Y, fB);
  walls_bounce_back(fA, cA, X, Y);
  walls_bounce_back(fB, cB, X, Y);
  /* bind :181-187 */
  for (int y = 0; y < Y; y++) fA[IDX(X - 1, y, 3)] = cB[IDX(0, y, 3)];
  for (int y = 1; y < Y; y++) fA[IDX(X - 1, y, 6)] = cB[IDX(0, y - 1, 6)];
  for (int y = 0; y < Y - 1; y++) fA[IDX(X - 1, y, 7)] = cB[IDX(0, y + 1, 7)];
  for (int y = 0; y < Y; y++) fB[IDX(0, y, 1)] = cA[IDX(X - 1, y, 1)];
  for (int y = 1; y < Y; y++) fB[IDX(0, y, 5)] = cA[IDX(X - 1, y - 1, 5)];
  for (int y = 0; y < Y - 1; y++) fB[IDX(0, y, 8)] = cA[IDX(X - 1, y + 1, 8)];
  free(eqA); free(eqB); free(cA); free(cB);
}

/* ------------------------------------------------------------------ immersed boundary */

struct orc_ibm
{
  int n, m_max;
  long r0, r1, c0, c1; /* ROI slices [r0,r1) x [c0,c1) : src/ibm.cpp:122-156 */
  long* mrow;          /* marker box start row/col, ROI coordinates: src/ibm.cpp:30-36 */
  long* mcol;
  double* phi;         /* {n,16} : src/ibm.cpp:26-28,47-57 */
};

/* src/ibm.cpp:39-45 */
static double peskin(double r_)
{
  double r = fabs(r_);
  if (r <= 1) return 0.125 * (3.0 - 2.0 * r + sqrt(1.0 + 4.0 * r - 4.0 * r * r));
  else if (r <= 2) return 0.125 * (5.0 - 2.0 * r - sqrt(-7.0 + 12.0 * r - 4.0 * r * r));
  return 0.0;
}

orc_ibm* orc_ibm_create(const double* xs, const double* ys, int n, int m_max)
{
  orc_ibm* ib = (orc_ibm*)calloc(1, sizeof(orc_ibm));
  ib->n = n;
  ib->m_max = m_max;
  long r_min = 1000000, r_max = 0, c_min = 1000000, c_max = 0;
  for (int i = 0; i < n; i++)
  {
    if (r_min > (int)(floor(xs[i]) - 2)) r_min = (int)(floor(xs[i]) - 2);
    if (r_max < (int)(floor(xs[i]) + 2)) r_max = (int)(floor(xs[i]) + 2);
    if (c_min > (int)(floor(ys[i]) - 2)) c_min = (int)(floor(ys[i]) - 2);
    if (c_max < (int)(floor(ys[i]) + 2)) c_max = (int)(floor(ys[i]) + 2);
  }
  ib->r0 = r_min; ib->r1 = r_max + 1; ib->c0 = c_min; ib->c1 = c_max + 1;
  ib->mrow = (long*)calloc(n ? n : 1, sizeof(long));
  ib->mcol = (long*)calloc(n ? n : 1, sizeof(long));
  ib->phi = dalloc((size_t)n * 16);
  for (int i = 0; i < n; i++)
  {
    /* marker(x_m - r_off, y_m - c_off): src/ibm.cpp:101 */
    double x = xs[i] - (double)ib->r0, y = ys[i] - (double)ib->c0;
    for (int k = 0; k < 16; k++)
    {
      /* stencil rows: {0,1,2,3,0,1,2,3,...} pairs with x, {0,0,0,0,1,...} with y (src/ibm.cpp:11-13,26) */
      double sx = x - ((double)(k % 4) + floor(x) - 1.0);
      double sy = y - ((double)(k / 4) + floor(y) - 1.0);
      ib->phi[(size_t)i * 16 + k] = peskin(sx) * peskin(sy);
    }
    ib->mrow[i] = (long)floor(x) - 1;
    ib->mcol[i] = (long)floor(y) - 1;
  }
  return ib;
}

void orc_ibm_destroy(orc_ibm* ib)
{
  if (!ib) return;
  free(ib->mrow); free(ib->mcol); free(ib->phi); free(ib);
}

void orc_ibm_roi(const orc_ibm* ib, long* roi)
{
  roi[0] = ib->r0; roi[1] = ib->r1; roi[2] = ib->c0; roi[3] = ib->c1;
}

/* src/ibm.cpp:158-190 */
void orc_ibm_force(orc_ibm* ib, const double* u0, const double* rho0, int X, int Y, double* F_out)
{
  (void)X;
  const long RR = ib->r1 - ib->r0, RC = ib->c1 - ib->c0;
  double* u = dalloc((size_t)RR * RC * 2);
  double* rho = dalloc((size_t)RR * RC);
  double* Fn = dalloc((size_t)RR * RC * 2);
  for (long i = 0; i < RR; i++)
    for (long j = 0; j < RC; j++)
    {
      size_t g = N2(ib->r0 + i, ib->c0 + j);
      u[(i * RC + j) * 2] = u0[g * 2];
      u[(i * RC + j) * 2 + 1] = u0[g * 2 + 1];
      rho[i * RC + j] = rho0[g];
    }
  memset(F_out, 0, sizeof(double) * RR * RC * 2);
  for (int n = 1; n < ib->m_max; n++)
  {
    memset(Fn, 0, sizeof(double) * RR * RC * 2);
    for (int m = 0; m < ib->n; m++)
    {
      const double* phi = ib->phi + (size_t)m * 16;
      double ujx = 0.0, ujy = 0.0, rhoj = 0.0;
      /* box.reshape({16,2}): k = 4*row_local + col_local */
      for (int k = 0; k < 16; k++)
      {
        long i = ib->mrow[m] + k / 4, j = ib->mcol[m] + k % 4;
        ujx += phi[k] * u[(i * RC + j) * 2];
        ujy += phi[k] * u[(i * RC + j) * 2 + 1];
        rhoj += phi[k] * rho[i * RC + j];
      }
      double fjx = -2.0 * rhoj * ujx, fjy = -2.0 * rhoj * ujy;
      for (int k = 0; k < 16; k++)
      {
        long i = ib->mrow[m] + k / 4, j = ib->mcol[m] + k % 4;
        Fn[(i * RC + j) * 2] += phi[k] * fjx;
        Fn[(i * RC + j) * 2 + 1] += phi[k] * fjy;
      }
    }
    for (long k = 0; k < RR * RC; k++)
    {
      u[k * 2] += 0.5 * Fn[k * 2] / rho[k];
      u[k * 2 + 1] += 0.5 * Fn[k * 2 + 1] / rho[k];
      F_out[k * 2] += Fn[k * 2];
      F_out[k * 2 + 1] += Fn[k * 2 + 1];
    }
  }
  free(u); free(rho); free(Fn);
}

/* ------------------------------------------------------------------ driver 11 */

void orc_cylinder_step(double* f, double* u, double* rho, int X, int Y, double omega, double u_lb,
                       orc_ibm* ib, double* F_out)
{
  size_t N = (size_t)X * Y;
  double* feq = dalloc(N * 9);
  double* fcoll = dalloc(N * 9);
  const long RR = ib->r1 - ib->r0, RC = ib->c1 - ib->c0;
  double* F = F_out ? F_out : dalloc((size_t)RR * RC * 2);
  /* :100-108 */
  orc_calc_rho(f, X, Y, rho);
  orc_calc_u(f, rho, X, Y, u);
  orc_equilibrium(u, rho, X, Y, feq);
  /* :110 */
  orc_ibm_force(ib, u, rho, X, Y, F);
  /* :121-125  f_coll = f_adve + (-omega (f_adve - f_equi)) */
#pragma omp parallel for
  for (long n = 0; n < (long)(N * 9); n++) fcoll[n] = f[n] + (-omega * (f[n] - feq[n]));
  /* :116-127 ; ics2 = 1/3, ics4 = 1/9 as named in the driver (:65-66) */
  const double ics2 = 1.0 / 3.0, ics4 = 1.0 / 9.0;
  for (long i = 0; i < RR; i++)
    for (long j = 0; j < RC; j++)
    {
      size_t g = N2(ib->r0 + i, ib->c0 + j);
      double ux = u[g * 2], uy = u[g * 2 + 1];
      double Fx = F[(i * RC + j) * 2], Fy = F[(i * RC + j) * 2 + 1];
      double uF = ux * Fx + uy * Fy;
      for (int q = 0; q < 9; q++)
      {
        double cu = ux * CXD[q] + uy * CYD[q];
        double cF = Fx * CXD[q] + Fy * CYD[q];
        double S = ((1 - 0.5 * omega) * ((ics2 + ics4 * cu) * cF - ics2 * uF)) * W9[q];
        fcoll[g * 9 + q] += S;
      }
    }
  /* :130-163 */
  orc_advect(fcoll, X, Y, f);
  abb_rows(f, fcoll, X, Y, u_lb, 0.0);
  walls_specular(f, fcoll, X, Y);
  free(feq);
  free(fcoll);
  if (!F_out) free(F);
}

/* ------------------------------------------------------------------ driver 15 */

/* test/rectangle_sedimentation_test.cpp:80-107 */
void orc_sedimentation_init(double* f, double* g, double* u, double* rho, double* C, int X, int Y,
                            double u_lb, const double* C_w)
{
  size_t N = (size_t)X * Y;
  for (size_t n = 0; n < N; n++)
  {
    u[n * 2] = 0.0;
    u[n * 2 + 1] = u_lb;
    rho[n] = 1.0;
    C[n] = 0.0;
  }
  for (int x = 0; x < X; x++) C[N2(x, 0)] = C_w[x];
  orc_equilibrium(u, C, X, Y, g);
  orc_incomp_equilibrium(u, rho, X, Y, f);
  orc_calc_rho(f, X, Y, rho);
  orc_calc_u(f, rho, X, Y, u);
}

static void sedimentation_step_impl(double* f, double* g, double* u, double* rho, double* C, int X, int Y,
                                    double omega, double u_lb, double w_s, const double* C_w, int R23, int C28,
                                    int C38, orc_ibm* ib)
{
  size_t N = (size_t)X * Y;
  double* feq = dalloc(N * 9);
  double* geq = dalloc(N * 9);
  double* fc = dalloc(N * 9);
  double* gc = dalloc(N * 9);
  double* us = dalloc(N * 2);
  const int rw = X + R23; /* R23 is negative: row index from the end (:73) */
  /* :123-131 */
  orc_equilibrium(u, rho, X, Y, feq);
  for (size_t n = 0; n < N * 2; n++) us[n] = u[n] + w_s;
  orc_equilibrium(us, C, X, Y, geq);
  if (!ib) orc_collision(f, feq, omega, X, Y, fc);
  else
  {
    /* BASELINE configs[4] "with immersed-boundary coupling": the collision of test/cylinder_test.cpp:110-127 (force
     * density from u, rho; f_coll = f + (-omega (f - f_eq)) + S inside the ROI) in place of solver::collision */
    const long RR = ib->r1 - ib->r0, RC = ib->c1 - ib->c0;
    double* F = dalloc((size_t)RR * RC * 2);
    orc_ibm_force(ib, u, rho, X, Y, F);
    for (size_t n = 0; n < N * 9; n++) fc[n] = f[n] + (-omega * (f[n] - feq[n]));
    const double ics2 = 1.0 / 3.0, ics4 = 1.0 / 9.0;
    for (long i = 0; i < RR; i++)
      for (long j = 0; j < RC; j++)
      {
        size_t gg = N2(ib->r0 + i, ib->c0 + j);
        double ux = u[gg * 2], uy = u[gg * 2 + 1];
        double Fx = F[(i * RC + j) * 2], Fy = F[(i * RC + j) * 2 + 1];
        double uF = ux * Fx + uy * Fy;
        for (int q = 0; q < 9; q++)
        {
          double cu = ux * CXD[q] + uy * CYD[q];
          double cF = Fx * CXD[q] + Fy * CYD[q];
          fc[gg * 9 + q] += ((1 - 0.5 * omega) * ((ics2 + ics4 * cu) * cF - ics2 * uF)) * W9[q];
        }
      }
    free(F);
  }
  orc_collision(g, geq, omega / 1.0, X, Y, gc);
  /* zero gradient :137-141 */
  for (int y = 0; y < Y; y++)
    for (int q = 0; q < 9; q++) gc[IDX(0, y, q)] = gc[IDX(1, y, q)];
  for (int x = 1; x < X - 1; x++)
    for (int q = 0; q < 9; q++) gc[IDX(x, Y - 1, q)] = gc[IDX(x, Y - 2, q)];
  /* :144-145 */
  orc_advect(fc, X, Y, f);
  orc_advect(gc, X, Y, g);
  /* ABB inlet, fixed u_w = (0, u_lb), rows 1..X-2 of column 0 :150-161 */
  double abb[9];
  abb_const(0.0, u_lb, abb);
  for (int x = 1; x < X - 1; x++)
    for (int q = 1; q < 9; q++) f[IDX(x, 0, OPP[q])] = -fc[IDX(x, 0, q)] + abb[q];
  /* ABB outlet with extrapolated wall velocity, all rows of the last column :163-172 */
  for (int x = 0; x < X; x++)
  {
    double uwx = 1.5 * u[N2(x, Y - 1) * 2] - 0.5 * u[N2(x, Y - 2) * 2];
    double uwy = 1.5 * u[N2(x, Y - 1) * 2 + 1] - 0.5 * u[N2(x, Y - 2) * 2 + 1];
    abb_const(uwx, uwy, abb);
    for (int q = 1; q < 9; q++) f[IDX(x, Y - 1, OPP[q])] = -fc[IDX(x, Y - 1, q)] + abb[q];
  }
  /* specular top :175-177, no-slip bottom :180-182 */
  for (int y = 0; y < Y; y++)
  {
    f[IDX(0, y, 8)] = fc[IDX(0, y, 7)];
    f[IDX(0, y, 1)] = fc[IDX(0, y, 3)];
    f[IDX(0, y, 5)] = fc[IDX(0, y, 6)];
    f[IDX(X - 1, y, 7)] = fc[IDX(X - 1, y, 5)];
    f[IDX(X - 1, y, 3)] = fc[IDX(X - 1, y, 1)];
    f[IDX(X - 1, y, 6)] = fc[IDX(X - 1, y, 8)];
  }
  /* rectangle :186-196 */
  for (int x = rw + 1; x < X - 1; x++)
  {
    f[IDX(x, C28, 8)] = fc[IDX(x, C28, 6)];
    f[IDX(x, C28, 4)] = fc[IDX(x, C28, 2)];
    f[IDX(x, C28, 7)] = fc[IDX(x, C28, 5)];
  }
  for (int y = C28; y < C38 + 1; y++)
  {
    f[IDX(rw, y, 6)] = fc[IDX(rw, y, 8)];
    f[IDX(rw, y, 3)] = fc[IDX(rw, y, 1)];
    f[IDX(rw, y, 7)] = fc[IDX(rw, y, 5)];
  }
  for (int x = rw + 1; x < X - 1; x++)
  {
    f[IDX(x, C38, 5)] = fc[IDX(x, C38, 7)];
    f[IDX(x, C38, 2)] = fc[IDX(x, C38, 4)];
    f[IDX(x, C38, 6)] = fc[IDX(x, C38, 8)];
  }
  /* :199-201 */
  orc_calc_rho(f, X, Y, rho);
  orc_calc_u(f, rho, X, Y, u);
  /* ADE inlet :204-218 (u + w_s adds the scalar to both components) */
  for (int x = 1; x < X - 1; x++)
  {
    double ax = u[N2(x, 0) * 2] + w_s, ay = u[N2(x, 0) * 2 + 1] + w_s;
    double aa = ax * ax + ay * ay;
    for (int q = 1; q < 9; q++)
    {
      double cu = ax * CXD[q] + ay * CYD[q];
      double gb = ((1.0 + 3.0 * cu + 4.5 * (cu * cu) - 1.5 * aa) * W9[q]) * C_w[x];
      g[IDX(x, 0, OPP[q])] = -gc[IDX(x, 0, q)] + 2.0 * gb;
    }
  }
  /* rectangle on g :222-236 */
  for (int x = rw + 1; x < X; x++)
  {
    g[IDX(x, C28, 8)] = -gc[IDX(x, C28, 6)];
    g[IDX(x, C28, 4)] = -gc[IDX(x, C28, 2)];
    g[IDX(x, C28, 7)] = -gc[IDX(x, C28, 5)];
  }
  for (int y = C28; y < C38 + 1; y++)
  {
    g[IDX(rw, y, 6)] = -gc[IDX(rw, y, 8)];
    g[IDX(rw, y, 3)] = -gc[IDX(rw, y, 1)];
    g[IDX(rw, y, 7)] = -gc[IDX(rw, y, 5)];
  }
  for (int x = rw + 1; x < X - 1; x++)
  {
    g[IDX(x, C38, 5)] = -gc[IDX(x, C38, 7)];
    g[IDX(x, C38, 2)] = -gc[IDX(x, C38, 4)];
    g[IDX(x, C38, 6)] = -gc[IDX(x, C38, 8)];
  }
  for (int y = 0; y < Y; y++)
  {
    g[IDX(X - 1, y, 6)] = gc[IDX(X - 1, y, 8)];
    g[IDX(X - 1, y, 3)] = gc[IDX(X - 1, y, 1)];
    g[IDX(X - 1, y, 7)] = gc[IDX(X - 1, y, 5)];
  }
  /* :237 */
  orc_calc_rho(g, X, Y, C);
  free(feq); free(geq); free(fc); free(gc); free(us);
}

void orc_sedimentation_step(double* f, double* g, double* u, double* rho, double* C, int X, int Y,
                            double omega, double u_lb, double w_s, const double* C_w, int R23, int C28,
                            int C38)
{
  sedimentation_step_impl(f, g, u, rho, C, X, Y, omega, u_lb, w_s, C_w, R23, C28, C38, NULL);
}

/* NOT a reference driver: the sedimentation loop with an immersed body added the way test/cylinder_test.cpp couples one
 * (BASELINE configs[4] words the case "with immersed-boundary coupling (ibm)"; driver 15 itself has none).  Composed of
 * the two pinned steps, itself unpinned. */
void orc_sedimentation_ibm_step(double* f, double* g, double* u, double* rho, double* C, int X, int Y,
                                double omega, double u_lb, double w_s, const double* C_w, int R23, int C28,
                                int C38, orc_ibm* ib)
{
  sedimentation_step_impl(f, g, u, rho, C, X, Y, omega, u_lb, w_s, C_w, R23, C28, C38, ib);
}

/* ------------------------------------------------------------------ drivers 16 / 18 (MRT colour gradient) */

/* test/mrtcg_rayleigh_taylor.cpp:130-156 */
static const double MM[9][9] = {{1, 1, 1, 1, 1, 1, 1, 1, 1},
                                {-4, -1, -1, -1, -1, 2, 2, 2, 2},
                                {4, -2, -2, -2, -2, 1, 1, 1, 1},
                                {0, 1, 0, -1, 0, 1, -1, -1, 1},
                                {0, -2, 0, 2, 0, 1, -1, -1, 1},
                                {0, 0, 1, 0, -1, 1, 1, -1, -1},
                                {0, 0, -2, 0, 2, 1, 1, -1, -1},
                                {0, 1, -1, 1, -1, 0, 0, 0, 0},
                                {0, 0, 0, 0, 0, 1, -1, 1, -1}};
static const double MI36[9][9] = {{4, -4, 4, 0, 0, 0, 0, 0, 0},
                                  {4, -1, -2, 6, -6, 0, 0, 9, 0},
                                  {4, -1, -2, 0, 0, 6, -6, -9, 0},
                                  {4, -1, -2, -6, 6, 0, 0, 9, 0},
                                  {4, -1, -2, 0, 0, -6, 6, -9, 0},
                                  {4, 2, 1, 6, 3, 6, 3, 0, 9},
                                  {4, 2, 1, -6, -3, 6, 3, 0, -9},
                                  {4, 2, 1, -6, -3, -6, -3, 0, 9},
                                  {4, 2, 1, 6, 3, -6, -3, 0, -9}};
/* :158-163 */
static const double BB9[9] = {-4.0 / 27.0, 2.0 / 27.0, 2.0 / 27.0, 2.0 / 27.0, 2.0 / 27.0,
                              5.0 / 108.0, 5.0 / 108.0, 5.0 / 108.0, 5.0 / 108.0};

static double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

/* test/mrtcg_rayleigh_taylor.cpp:182-210 */
void orc_mrtcg_init_rt(const orc_mrtcg_params* p, double* r_rho, double* b_rho)
{
  const int R = p->R, C = p->C;
  const double middle = R / 2.0;
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++)
    {
      double s = middle - 0.1 * C * cos(2.0 * 3.141592 * c / C);
      r_rho[(size_t)r * C + c] = p->r_rho0 * ((r < s) ? 1.0 : 0.0);
      b_rho[(size_t)r * C + c] = p->b_rho0 * ((r >= s) ? 1.0 : 0.0);
    }
}

/* test/mrtcg_static_droplet.cpp:182-204 */
void orc_mrtcg_init_droplet(const orc_mrtcg_params* p, double* r_rho, double* b_rho)
{
  const int R = p->R, C = p->C;
  const double center = R / 2.0, radius = 25.0;
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++)
    {
      double s = sqrt((r - center) * (r - center) + (c - center) * (c - center));
      r_rho[(size_t)r * C + c] = p->r_rho0 * (1.0 - sigmoid(1.0 * (s - radius)));
      b_rho[(size_t)r * C + c] = p->b_rho0 * sigmoid(1.0 * (s - radius));
    }
}

/* test/mrtcg_rayleigh_taylor.cpp:233-247 */
static void mrtcg_eq_node(double rho_k, const double* phi, const double* eta, double ux, double uy, double* out)
{
  double uu = ux * ux + uy * uy;
  for (int q = 0; q < 9; q++)
  {
    double ue = ux * CXD[q] + uy * CYD[q];
    out[q] = rho_k * (phi[q] + W9[q] * ((3.0 * ue) * eta[q] + 9.0 * (ue * ue) - 3.0 * uu));
  }
}

void orc_mrtcg_init_state(const orc_mrtcg_params* p, const double* r_rho, const double* b_rho, double* rho,
                          double* u, double* r_adv, double* b_adv, int shift_u)
{
  const size_t N = (size_t)p->R * p->C;
  double cs2, rlx, rphi[9], reta[9], bphi[9], beta_[9];
  colour_derive(p->r_alpha, p->r_nu, &cs2, &rlx, rphi, reta);
  colour_derive(p->b_alpha, p->b_nu, &cs2, &rlx, bphi, beta_);
  for (size_t n = 0; n < N; n++)
  {
    rho[n] = r_rho[n] + b_rho[n];
    u[n * 2] = 0.0;
    u[n * 2 + 1] = 0.0;
    if (shift_u) /* test/mrtcg_static_droplet.cpp:457 */
    {
      u[n * 2] = u[n * 2] + 0.5 * p->Fg[0] / rho[n];
      u[n * 2 + 1] = u[n * 2 + 1] + 0.5 * p->Fg[1] / rho[n];
    }
    mrtcg_eq_node(r_rho[n], rphi, reta, u[n * 2], u[n * 2 + 1], r_adv + n * 9);
    mrtcg_eq_node(b_rho[n], bphi, beta_, u[n * 2], u[n * 2 + 1], b_adv + n * 9);
  }
}

typedef struct
{
  double delta, r_omega, b_omega, s1, s2, s3, t2, t3;
} relax_fn;

/* test/mrtcg_rayleigh_taylor.cpp:34-82 (and rk_static_droplet_test.cpp:287-339 in tau space) */
static relax_fn relax_init(double r_val, double b_val, double delta)
{
  relax_fn f;
  f.delta = delta;
  f.r_omega = r_val;
  f.b_omega = b_val;
  f.s1 = 2.0 * r_val * b_val / (r_val + b_val);
  f.s2 = 2.0 * (r_val - f.s1) / delta;
  f.s3 = -f.s2 / (2.0 * delta);
  f.t2 = 2.0 * (f.s1 - b_val) / delta;
  f.t3 = f.t2 / (2.0 * delta);
  return f;
}

/* :84-100 — a NaN psi matches no branch and leaves s_nu untouched */
static double relax_eval(const relax_fn* f, double psi, double old)
{
  double s = old;
  if (psi > f->delta) s = f->r_omega;
  if (f->delta >= psi && psi > 0.0) s = f->s1 + f->s2 * psi + f->s3 * psi * psi;
  if (0.0 >= psi && psi >= -f->delta) s = f->s1 + f->t2 * psi + f->t3 * psi * psi;
  if (psi < -f->delta) s = f->b_omega;
  return s;
}

/* test/mrtcg_rayleigh_taylor.cpp:495-533 */
static void mrtcg_bc(double* adv, const double* col, int X, int Y)
{
  for (int x = 1; x < X - 1; x++)
  {
    adv[IDX(x, 0, 2)] = col[IDX(x, Y - 1, 2)];
    adv[IDX(x, 0, 5)] = col[IDX(x, Y - 1, 5)];
    adv[IDX(x, 0, 6)] = col[IDX(x, Y - 1, 6)];
    adv[IDX(x, Y - 1, 4)] = col[IDX(x, 0, 4)];
    adv[IDX(x, Y - 1, 8)] = col[IDX(x, 0, 8)];
    adv[IDX(x, Y - 1, 7)] = col[IDX(x, 0, 7)];
  }
  for (int y = 0; y < Y; y++)
  {
    adv[IDX(X - 1, y, 3)] = col[IDX(X - 1, y, 1)];
    adv[IDX(X - 1, y, 7)] = col[IDX(X - 1, y, 5)];
    adv[IDX(X - 1, y, 6)] = col[IDX(X - 1, y, 8)];
    adv[IDX(0, y, 1)] = col[IDX(0, y, 3)];
    adv[IDX(0, y, 5)] = col[IDX(0, y, 7)];
    adv[IDX(0, y, 8)] = col[IDX(0, y, 6)];
  }
}

void orc_mrtcg_step(const orc_mrtcg_params* p, double* r_adv, double* b_adv, double* r_rho, double* b_rho,
                    double* rho, double* u, double* phase, double* s_nu, double* grad)
{
  const int X = p->R, Y = p->C;
  const size_t N = (size_t)X * Y;
  double r_cs2, r_rlx, rphi[9], reta[9], b_cs2, b_rlx, bphi[9], beta_[9];
  colour_derive(p->r_alpha, p->r_nu, &r_cs2, &r_rlx, rphi, reta);
  colour_derive(p->b_alpha, p->b_nu, &b_cs2, &b_rlx, bphi, beta_);
  /* relaxation_function{r, b, 0.1}: omegas re-derived from nu, cs2 (:57-66) */
  relax_fn rf = relax_init(1.0 / (0.5 + p->r_nu / r_cs2), 1.0 / (0.5 + p->b_nu / b_cs2), p->delta);

  double* Qx = dalloc(N);
  double* Qy = dalloc(N);
  double* tmp = dalloc(N);
  double* rDxQx = dalloc(N);
  double* rDyQy = dalloc(N);
  double* bDxQx = dalloc(N);
  double* bDyQy = dalloc(N);
  double* gx = dalloc(N);
  double* gy = dalloc(N);
  double* r_col = dalloc(N * 9);
  double* b_col = dalloc(N * 9);

  /* :434-435 */
  for (size_t n = 0; n < N; n++)
  {
    phase[n] = (r_rho[n] / p->r_rho0 - b_rho[n] / p->b_rho0) / (r_rho[n] / p->r_rho0 + b_rho[n] / p->b_rho0);
    s_nu[n] = relax_eval(&rf, phase[n], s_nu[n]);
  }
  /* update_C :320-336 */
  for (size_t n = 0; n < N; n++)
  {
    Qx[n] = ((1.8 * p->r_alpha - 0.8) * r_rho[n]) * u[n * 2];
    Qy[n] = ((1.8 * p->r_alpha - 0.8) * r_rho[n]) * u[n * 2 + 1];
  }
  orc_diff5(Qx, X, Y, rDxQx, tmp);
  orc_diff5(Qy, X, Y, tmp, rDyQy);
  for (size_t n = 0; n < N; n++)
  {
    Qx[n] = ((1.8 * p->b_alpha - 0.8) * b_rho[n]) * u[n * 2];
    Qy[n] = ((1.8 * p->b_alpha - 0.8) * b_rho[n]) * u[n * 2 + 1];
  }
  orc_diff5(Qx, X, Y, bDxQx, tmp);
  orc_diff5(Qy, X, Y, tmp, bDyQy);
  /* D.grad(grad, phase) :443 */
  orc_diff5(phase, X, Y, gx, gy);

  const double Sdiag[9] = {0.0, 1.25, 1.14, 0.0, 1.6, 0.0, 1.6, 0.0, 0.0}; /* :384-387 */
  const double SQ2 = sqrt(2);
#pragma omp parallel for
  for (long n = 0; n < (long)N; n++)
  {
    double ux = u[n * 2], uy = u[n * 2 + 1];
    double feq[9], m[9], om1[2][9], om2[9], xi[9], kap[9], total[9];
    double S[9];
    for (int q = 0; q < 9; q++) S[q] = Sdiag[q];
    S[7] = s_nu[n];
    S[8] = s_nu[n];
    /* MRT operator per colour :249-261 */
    for (int k = 0; k < 2; k++)
    {
      const double* fk = (k == 0 ? r_adv : b_adv) + n * 9;
      double rk = (k == 0 ? r_rho[n] : b_rho[n]);
      mrtcg_eq_node(rk, k == 0 ? rphi : bphi, k == 0 ? reta : beta_, ux, uy, feq);
      double DxQx = (k == 0 ? rDxQx[n] : bDxQx[n]);
      double DyQy = (k == 0 ? rDyQy[n] : bDyQy[n]);
      double Ck[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      Ck[1] = 3.0 * (1.0 - 0.5 * 1.25) * (DxQx + DyQy);
      Ck[7] = (1.0 - 0.5 * s_nu[n]) * (DxQx - DyQy);
      for (int a = 0; a < 9; a++)
      {
        double s = 0.0;
        for (int q = 0; q < 9; q++) s += MM[a][q] * (feq[q] - fk[q]);
        m[a] = S[a] * s + Ck[a];
      }
      for (int q = 0; q < 9; q++)
      {
        double s = 0.0;
        for (int a = 0; a < 9; a++) s += ((1.0 / 36.0) * MI36[q][a]) * m[a];
        om1[k][q] = s;
      }
    }
    grad[n * 2] = gx[n];
    grad[n * 2 + 1] = gy[n];
    double gn = sqrt(gx[n] * gx[n] + gy[n] * gy[n]); /* :444-447 */
    double A = 4.5 * p->sigma * s_nu[n];             /* :450 */
    for (int q = 0; q < 9; q++)
    {
      double ge = gx[n] * CXD[q] + gy[n] * CYD[q];
      double t = ge / (1e-20 + gn);
      xi[q] = (0.5 * gn) * (W9[q] * (t * t) - BB9[q]); /* :290-300 */
      om2[q] = A * xi[q];                                /* :451-452 */
      /* kappa :302-318 ; unit_E = E / {1,1,1,1,1,sqrt2,...} */
      double nrm = (q < 5) ? 1.0 : SQ2;
      double gue = gx[n] * (CXD[q] / nrm) + gy[n] * (CYD[q] / nrm);
      kap[q] = (((r_rho[n] * b_rho[n]) * gue) * (r_rho[n] * rphi[q] + b_rho[n] * bphi[q])) /
               ((rho[n] * rho[n]) * (1e-20 + gn));
      /* :455 */
      total[q] = ((((r_adv[n * 9 + q] + om1[0][q]) + om2[q]) + b_adv[n * 9 + q]) + om1[1][q]) + om2[q];
    }
    double uFg = ux * p->Fg[0] + uy * p->Fg[1];
    for (int q = 0; q < 9; q++)
    {
      double o3r = r_rho[n] * total[q] / rho[n] + p->r_beta * kap[q]; /* :275-288 */
      double o3b = b_rho[n] * total[q] / rho[n] + p->b_beta * kap[q];
      if (p->add_force) /* :460-464 ; ics2 = 3, ics4 = 9 */
      {
        double ue = ux * CXD[q] + uy * CYD[q];
        double Fe = p->Fg[0] * CXD[q] + p->Fg[1] * CYD[q];
        double src = ((1 - 0.5 * s_nu[n]) * ((3.0 + 9.0 * ue) * Fe - 3.0 * uFg)) * W9[q];
        o3r += src;
        o3b += src;
      }
      r_col[n * 9 + q] = o3r;
      b_col[n * 9 + q] = o3b;
    }
  }
  /* :466-470 */
  orc_advect(r_col, X, Y, r_adv);
  orc_advect(b_col, X, Y, b_adv);
  mrtcg_bc(r_adv, r_col, X, Y);
  mrtcg_bc(b_adv, b_col, X, Y);
  /* :472-477 */
  orc_calc_rho(r_adv, X, Y, r_rho);
  orc_calc_rho(b_adv, X, Y, b_rho);
  for (size_t n = 0; n < N; n++)
  {
    rho[n] = r_rho[n] + b_rho[n];
    double sx = 0.0, sy = 0.0;
    for (int q = 0; q < 9; q++)
    {
      double t = r_adv[n * 9 + q] + b_adv[n * 9 + q];
      sx += t * CXD[q];
      sy += t * CYD[q];
    }
    u[n * 2] = sx / rho[n] + 0.5 * p->Fg[0] / rho[n];
    u[n * 2 + 1] = sy / rho[n] + 0.5 * p->Fg[1] / rho[n];
  }
  free(Qx); free(Qy); free(tmp); free(rDxQx); free(rDyQy); free(bDxQx); free(bDyQy);
  free(gx); free(gy); free(r_col); free(b_col);
}

/* ------------------------------------------------------------------ driver 17 (Rothman-Keller droplet) */

/* test/rk_static_droplet_test.cpp:183-199 */
static void rk_eq_node(double rho_k, const double* phi, double ux, double uy, double* out)
{
  const double ics2 = 3.0;
  double uu = ux * ux + uy * uy;
  for (int q = 0; q < 9; q++)
  {
    double ue = ux * CXD[q] + uy * CYD[q];
    out[q] = rho_k * (phi[q] + ((ics2 * ue + (0.5 * ics2 * ics2) * (ue * ue)) - (0.5 * ics2) * uu) * W9[q]);
  }
}

static void rk_phi(double alpha, double* phi)
{
  double a = 0.2 * (1 - alpha), b = 0.05 * (1 - alpha);
  phi[0] = alpha;
  for (int q = 1; q < 5; q++) phi[q] = a;
  for (int q = 5; q < 9; q++) phi[q] = b;
}

/* :363-396, 509-515 */
void orc_rk_init(const orc_rk_params* p, const double* u0, double* r_adv, double* b_adv, double* r_rho,
                 double* b_rho, double* rho_mix)
{
  const int L = p->L;
  const double Cc = L / 2.0, factor = 2.0;
  double rphi[9], bphi[9];
  rk_phi(p->r_alpha, rphi);
  rk_phi(p->b_alpha, bphi);
  for (int r = 0; r < L; r++)
    for (int c = 0; c < L; c++)
    {
      size_t n = (size_t)r * L + c;
      double s = sqrt((r - Cc) * (r - Cc) + (c - Cc) * (c - Cc));
      double rr = p->r_rho0 * (1.0 - sigmoid(factor * (s - p->radius)));
      double bb = p->b_rho0 * sigmoid(factor * (s - p->radius));
      double ux = u0 ? u0[n * 2] : 0.0, uy = u0 ? u0[n * 2 + 1] : 0.0;
      rk_eq_node(rr, rphi, ux, uy, r_adv + n * 9);
      rk_eq_node(bb, bphi, ux, uy, b_adv + n * 9);
    }
  orc_calc_rho(r_adv, L, L, r_rho);
  orc_calc_rho(b_adv, L, L, b_rho);
  for (size_t n = 0; n < (size_t)L * L; n++) rho_mix[n] = r_rho[n] + b_rho[n];
}

/* :204-211 */
static void rk_bc(double* adv, const double* col, int X, int Y)
{
  for (int x = 1; x < X - 1; x++)
    for (int q = 0; q < 9; q++) adv[IDX(x, 0, q)] = col[IDX(x, Y - 1, q)];
  for (int x = 1; x < X - 1; x++)
    for (int q = 0; q < 9; q++) adv[IDX(x, Y - 1, q)] = col[IDX(x, 0, q)];
  for (int y = 0; y < Y; y++)
    for (int q = 0; q < 9; q++) adv[IDX(0, y, q)] = col[IDX(X - 1, y, q)];
  for (int y = 0; y < Y; y++)
    for (int q = 0; q < 9; q++) adv[IDX(X - 1, y, q)] = col[IDX(0, y, q)];
}

void orc_rk_step(const orc_rk_params* p, double* r_adv, double* b_adv, double* r_rho, double* b_rho,
                 double* rho_mix, double* u, double* phase, double* relax, double* grad)
{
  const int X = p->L, Y = p->L;
  const size_t N = (size_t)X * Y;
  const double cs2 = 1.0 / 3.0;
  double rphi[9], bphi[9];
  rk_phi(p->r_alpha, rphi);
  rk_phi(p->b_alpha, bphi);
  /* colour::init_omega :264-265 ; relaxation_function in tau space :320-339 */
  double r_om = 1.0 / (0.5 + p->r_nu / cs2), b_om = 1.0 / (0.5 + p->b_nu / cs2);
  relax_fn rf = relax_init(1.0 / r_om, 1.0 / b_om, p->delta);
  double* gx = dalloc(N);
  double* gy = dalloc(N);
  double* r_col = dalloc(N * 9);
  double* b_col = dalloc(N * 9);
  /* :547-551 */
  for (size_t n = 0; n < N; n++)
    phase[n] = (r_rho[n] / p->r_rho0 - b_rho[n] / p->b_rho0) / (r_rho[n] / p->r_rho0 + b_rho[n] / p->b_rho0);
  orc_diff3(phase, X, Y, gx, gy);
#pragma omp parallel for
  for (long n = 0; n < (long)N; n++)
  {
    grad[n * 2] = gx[n];
    grad[n * 2 + 1] = gy[n];
    double gn = sqrt(gx[n] * gx[n] + gy[n] * gy[n]);
    /* :587-588 — eval() writes tau over the tensor that still holds last step's 1/tau (a NaN phase
     * matches no branch and keeps that value), then pow_(-1) inverts in place */
    double tau = relax_eval(&rf, phase[n], relax[n]);
    relax[n] = pow(tau, -1.0);
    double ux = u[n * 2], uy = u[n * 2 + 1];
    for (int k = 0; k < 2; k++)
    {
      const double* phi = k == 0 ? rphi : bphi;
      double* adv = (k == 0 ? r_adv : b_adv) + n * 9;
      double* col = (k == 0 ? r_col : b_col) + n * 9;
      double rk = k == 0 ? r_rho[n] : b_rho[n];
      double Ak = k == 0 ? p->r_A : p->b_A;
      double feq[9];
      rk_eq_node(rk, phi, ux, uy, feq);
      for (int q = 0; q < 9; q++)
      {
        double om1 = relax[n] * (feq[q] - adv[q]);                                          /* :255-262 */
        double fe = gx[n] * CXD[q] + gy[n] * CYD[q];
        double om2 = ((0.5 * Ak) * gn) * ((pow(fe, 2.0) / (1e-20 + pow(gn, 2.0))) * W9[q] - BB9[q]); /* :239-245 */
        col[q] = adv[q] + (om1 + om2);                                                      /* :176-178,232-236 */
      }
    }
  }
  orc_advect(r_col, X, Y, r_adv);
  rk_bc(r_adv, r_col, X, Y);
  orc_advect(b_col, X, Y, b_adv);
  rk_bc(b_adv, b_col, X, Y);
  /* :602-609 */
  orc_calc_rho(r_adv, X, Y, r_rho);
  orc_calc_rho(b_adv, X, Y, b_rho);
  for (size_t n = 0; n < N; n++)
  {
    rho_mix[n] = r_rho[n] + b_rho[n];
    double sx = 0.0, sy = 0.0;
    for (int q = 0; q < 9; q++)
    {
      double t = r_adv[n * 9 + q] + b_adv[n * 9 + q];
      sx += t * CXD[q];
      sy += t * CYD[q];
    }
    u[n * 2] = sx / rho_mix[n];
    u[n * 2 + 1] = sy / rho_mix[n];
  }
  free(gx); free(gy); free(r_col); free(b_col);
}

/* The per-iteration diagnostic fields driver 17 snapshots next to the state (test/rk_static_droplet_test.cpp:546-600):
 * everything is a function of the state at the TOP of loop iteration t (r_adv, r_rho, b_rho, rho_mix, u), so calling
 * this and then orc_rk_step reproduces iteration t.  None of it feeds the state update (step() only uses
 * omega3 = omega1 + omega2 of grad and |grad|, :213-237).
 *   phase :547 (rhons) ; grad, norm :550-558 ; n = -normalize(grad where |grad| > 0.1 max|grad| else 0) :559-567 ;
 *   K eval_local_curvature :440-446 ; Fs = sigma/2 K grad :572 ; eta eval_eta :398-413 ; kappa eval_kappa :415-438 ;
 *   rparams = 1 / tau(phase) :587-589 (in-out like orc_rk_step's relax) ; omega1, omega2 of the RED colour :255-262, :239-245.
 * Any output pointer may be NULL. */
void orc_rk_diagnostics(const orc_rk_params* p, double sigma, const double* r_adv, const double* r_rho, const double* b_rho,
                        const double* rho_mix, const double* u, double* phase_o, double* grad_o, double* norm_o, double* n_o,
                        double* K_o, double* Fs_o, double* eta_o, double* kappa_o, double* relax_io, double* omega1_o,
                        double* omega2_o)
{
  const int X = p->L, Y = p->L;
  const size_t N = (size_t)X * Y;
  const double cs2 = 1.0 / 3.0, ics2 = 3.0;
  double rphi[9];
  rk_phi(p->r_alpha, rphi);
  double r_om = 1.0 / (0.5 + p->r_nu / cs2), b_om = 1.0 / (0.5 + p->b_nu / cs2);
  relax_fn rf = relax_init(1.0 / r_om, 1.0 / b_om, p->delta);
  double* phase = dalloc(N);
  double* gx = dalloc(N); double* gy = dalloc(N); double* gn = dalloc(N);
  double* nx = dalloc(N); double* ny = dalloc(N);
  double* x_nx = dalloc(N); double* y_nx = dalloc(N); double* x_ny = dalloc(N); double* y_ny = dalloc(N);
  for (size_t n = 0; n < N; n++)
    phase[n] = (r_rho[n] / p->r_rho0 - b_rho[n] / p->b_rho0) / (r_rho[n] / p->r_rho0 + b_rho[n] / p->b_rho0);
  orc_diff3(phase, X, Y, gx, gy); /* grad[...,0] = partial.x = along axis 1, grad[...,1] = partial.y = along axis 0 */
  double gmax = 0.0;
  for (size_t n = 0; n < N; n++)
  {
    gn[n] = sqrt(pow(gx[n], 2.0) + pow(gy[n], 2.0));
    if (gn[n] > gmax) gmax = gn[n];
  }
  for (size_t n = 0; n < N; n++)
  {
    /* torch::where(norm <= 0.1 max, 0, grad), then F::normalize: v / max(||v||_2, 1e-12), negated */
    double cx = gn[n] <= 0.1 * gmax ? 0.0 : gx[n], cy = gn[n] <= 0.1 * gmax ? 0.0 : gy[n];
    double den = sqrt(cx * cx + cy * cy);
    if (den < 1e-12) den = 1e-12;
    nx[n] = -(cx / den);
    ny[n] = -(cy / den);
  }
  orc_diff3(nx, X, Y, x_nx, y_nx);
  orc_diff3(ny, X, Y, x_ny, y_ny);
  for (size_t n = 0; n < N; n++)
  {
    const double K = nx[n] * ny[n] * (y_nx[n] + x_ny[n]) - pow(nx[n], 2.0) * y_ny[n] - pow(ny[n], 2.0) * x_nx[n];
    const double Fsx = 0.5 * sigma * K * gx[n], Fsy = 0.5 * sigma * K * gy[n];
    const double ux = u[n * 2], uy = u[n * 2 + 1];
    if (phase_o) phase_o[n] = phase[n];
    if (grad_o) { grad_o[n * 2] = gx[n]; grad_o[n * 2 + 1] = gy[n]; }
    if (norm_o) norm_o[n] = gn[n];
    if (n_o) { n_o[n * 2] = nx[n]; n_o[n * 2 + 1] = ny[n]; }
    if (K_o) K_o[n] = K;
    if (Fs_o) { Fs_o[n * 2] = Fsx; Fs_o[n * 2 + 1] = Fsy; }
    double relax = 0.0;
    if (relax_io)
    {
      double tau = relax_eval(&rf, phase[n], relax_io[n]);
      relax = pow(tau, -1.0);
      relax_io[n] = relax;
    }
    double feq[9];
    rk_eq_node(r_rho[n], rphi, ux, uy, feq);
    for (int q = 0; q < 9; q++)
    {
      const double ue = ux * CXD[q] + uy * CYD[q];
      if (eta_o) /* sum over the two components of (ics2 (E - u) + ics2 (u.E) E) Fs, times W */
        eta_o[n * 9 + q] = ((ics2 * (CXD[q] - ux) + ics2 * (ue * CXD[q])) * Fsx + (ics2 * (CYD[q] - uy) + ics2 * (ue * CYD[q])) * Fsy) * W9[q];
      if (kappa_o) kappa_o[n * 9 + q] = (r_rho[n] * b_rho[n] / rho_mix[n]) * (((-nx[n]) * CXD[q] + (-ny[n]) * CYD[q]) * W9[q]);
      if (omega1_o && relax_io) omega1_o[n * 9 + q] = relax * (feq[q] - r_adv[n * 9 + q]);
      if (omega2_o)
      {
        const double fe = gx[n] * CXD[q] + gy[n] * CYD[q];
        omega2_o[n * 9 + q] = ((0.5 * p->r_A) * gn[n]) * ((pow(fe, 2.0) / (1e-20 + pow(gn[n], 2.0))) * W9[q] - BB9[q]);
      }
    }
  }
  free(phase); free(gx); free(gy); free(gn); free(nx); free(ny); free(x_nx); free(y_nx); free(x_ny); free(y_ny);
}

/* ------------------------------------------------------------------ test/mrt_rayleigh_taylor.cpp */

void orc_csf_init(const orc_csf_params* p, double* r_rho, double* b_rho, double* rho, double* u, double* r_adv, double* b_adv)
{
  const int R = p->R, C = p->C;
  const size_t N = (size_t)R * C;
  const double middle = R / 2.0;
  double cs2, rlx, rphi[9], reta[9], bphi[9], beta_[9];
  colour_derive(p->r_alpha, p->r_nu, &cs2, &rlx, rphi, reta);
  colour_derive(p->b_alpha, p->b_nu, &cs2, &rlx, bphi, beta_);
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++)
    {
      double s = middle + 0.1 * C * cos(2.0 * 3.141592 * c / C); /* :196 */
      r_rho[(size_t)r * C + c] = p->r_rho0 * ((r < s) ? 1.0 : 0.0);
      b_rho[(size_t)r * C + c] = p->b_rho0 * ((r >= s) ? 1.0 : 0.0);
    }
  for (size_t n = 0; n < N; n++)
  {
    rho[n] = r_rho[n] + b_rho[n];
    u[n * 2] = 0.0 + 0.5 * p->Fg[0] / p->r_rho0; /* :464: shifted by the RED reference density, everywhere */
    u[n * 2 + 1] = 0.0 + 0.5 * p->Fg[1] / p->r_rho0;
    mrtcg_eq_node(r_rho[n], rphi, reta, u[n * 2], u[n * 2 + 1], r_adv + n * 9);
    mrtcg_eq_node(b_rho[n], bphi, beta_, u[n * 2], u[n * 2 + 1], b_adv + n * 9);
  }
}

void orc_csf_step(const orc_csf_params* p, double* r_adv, double* b_adv, double* r_rho, double* b_rho, double* rho,
                  double* u, double* phase, double* s_nu, double* Fs)
{
  const int X = p->R, Y = p->C;
  const size_t N = (size_t)X * Y;
  double r_cs2, r_rlx, rphi[9], reta[9], b_cs2, b_rlx, bphi[9], beta_[9];
  colour_derive(p->r_alpha, p->r_nu, &r_cs2, &r_rlx, rphi, reta);
  colour_derive(p->b_alpha, p->b_nu, &b_cs2, &b_rlx, bphi, beta_);
  relax_fn rf = relax_init(1.0 / (0.5 + p->r_nu / r_cs2), 1.0 / (0.5 + p->b_nu / b_cs2), p->delta);

  double* Qx = dalloc(N);
  double* Qy = dalloc(N);
  double* tmp = dalloc(N);
  double* rDxQx = dalloc(N);
  double* rDyQy = dalloc(N);
  double* bDxQx = dalloc(N);
  double* bDyQy = dalloc(N);
  double* gx = dalloc(N);
  double* gy = dalloc(N);
  double* nx = dalloc(N);
  double* ny = dalloc(N);
  double* dx_nx = dalloc(N);
  double* dy_nx = dalloc(N);
  double* dx_ny = dalloc(N);
  double* dy_ny = dalloc(N);
  double* r_col = dalloc(N * 9);
  double* b_col = dalloc(N * 9);

  for (size_t n = 0; n < N; n++)
  {
    phase[n] = (r_rho[n] / p->r_rho0 - b_rho[n] / p->b_rho0) / (r_rho[n] / p->r_rho0 + b_rho[n] / p->b_rho0);
    s_nu[n] = relax_eval(&rf, phase[n], s_nu[n]);
  }
  for (size_t n = 0; n < N; n++)
  {
    Qx[n] = ((1.8 * p->r_alpha - 0.8) * r_rho[n]) * u[n * 2];
    Qy[n] = ((1.8 * p->r_alpha - 0.8) * r_rho[n]) * u[n * 2 + 1];
  }
  orc_diff5(Qx, X, Y, rDxQx, tmp);
  orc_diff5(Qy, X, Y, tmp, rDyQy);
  for (size_t n = 0; n < N; n++)
  {
    Qx[n] = ((1.8 * p->b_alpha - 0.8) * b_rho[n]) * u[n * 2];
    Qy[n] = ((1.8 * p->b_alpha - 0.8) * b_rho[n]) * u[n * 2 + 1];
  }
  orc_diff5(Qx, X, Y, bDxQx, tmp);
  orc_diff5(Qy, X, Y, tmp, bDyQy);
  orc_diff5(phase, X, Y, gx, gy);
  /* n = -grad / (1e-20 + |grad|) (:508), curvature from D applied to n (:355-364) */
  for (size_t n = 0; n < N; n++)
  {
    double gn = sqrt(gx[n] * gx[n] + gy[n] * gy[n]);
    nx[n] = -gx[n] / (1e-20 + gn);
    ny[n] = -gy[n] / (1e-20 + gn);
  }
  orc_diff5(nx, X, Y, dx_nx, dy_nx);
  orc_diff5(ny, X, Y, dx_ny, dy_ny);

  const double Sdiag[9] = {0.0, 1.25, 1.14, 0.0, 1.6, 0.0, 1.6, 0.0, 0.0};
  const double w2r = p->r_A * (1.0 - 0.5 * r_rlx), w2b = p->b_A * (1.0 - 0.5 * b_rlx); /* :512-513 */
#pragma omp parallel for
  for (long n = 0; n < (long)N; n++)
  {
    double ux = u[n * 2], uy = u[n * 2 + 1];
    double feq[9], m[9], om1[2][9], eta[9], kap[9], total[9];
    double S[9];
    for (int q = 0; q < 9; q++) S[q] = Sdiag[q];
    S[7] = s_nu[n];
    S[8] = s_nu[n];
    for (int k = 0; k < 2; k++)
    {
      const double* fk = (k == 0 ? r_adv : b_adv) + n * 9;
      double rk = (k == 0 ? r_rho[n] : b_rho[n]);
      mrtcg_eq_node(rk, k == 0 ? rphi : bphi, k == 0 ? reta : beta_, ux, uy, feq);
      double DxQx = (k == 0 ? rDxQx[n] : bDxQx[n]);
      double DyQy = (k == 0 ? rDyQy[n] : bDyQy[n]);
      double Ck[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      Ck[1] = 3.0 * (1.0 - 0.5 * 1.25) * (DxQx + DyQy);
      Ck[7] = (1.0 - 0.5 * s_nu[n]) * (DxQx - DyQy);
      for (int a = 0; a < 9; a++)
      {
        double s = 0.0;
        for (int q = 0; q < 9; q++) s += MM[a][q] * (feq[q] - fk[q]);
        m[a] = S[a] * s + Ck[a];
      }
      for (int q = 0; q < 9; q++)
      {
        double s = 0.0;
        for (int a = 0; a < 9; a++) s += ((1.0 / 36.0) * MI36[q][a]) * m[a];
        om1[k][q] = s;
      }
    }
    double gn = sqrt(gx[n] * gx[n] + gy[n] * gy[n]);
    /* eval_local_curvature; interf_tension = -0.5 sigma K grad (:509-510) */
    double K = nx[n] * ny[n] * (dy_nx[n] + dx_ny[n]) - (nx[n] * nx[n]) * dy_ny[n] - (ny[n] * ny[n]) * dx_nx[n];
    double fsx = (-0.5 * p->sigma) * K * gx[n], fsy = (-0.5 * p->sigma) * K * gy[n];
    for (int q = 0; q < 9; q++)
    {
      double ue = ux * CXD[q] + uy * CYD[q];
      /* eval_eta (:366-385): sum over the two components of (3 (E - u) + 9 (u.E) E) Fs, times W */
      eta[q] = ((3.0 * (CXD[q] - ux) + 9.0 * (ue * CXD[q])) * fsx + (3.0 * (CYD[q] - uy) + 9.0 * (ue * CYD[q])) * fsy) * W9[q];
      double ge = gx[n] * CXD[q] + gy[n] * CYD[q];
      /* eval_kappa with E, not unit_E (:304-320) */
      kap[q] = (((r_rho[n] * b_rho[n]) * ge) * (r_rho[n] * rphi[q] + b_rho[n] * bphi[q])) / ((rho[n] * rho[n]) * (1e-20 + gn));
      total[q] = ((((r_adv[n * 9 + q] + om1[0][q]) + w2r * eta[q]) + b_adv[n * 9 + q]) + om1[1][q]) + w2b * eta[q]; /* :522 */
    }
    double uFg = ux * p->Fg[0] + uy * p->Fg[1];
    for (int q = 0; q < 9; q++)
    {
      double o3r = r_rho[n] * total[q] / rho[n] + p->r_beta * kap[q];
      double o3b = b_rho[n] * total[q] / rho[n] + p->b_beta * kap[q];
      double ue = ux * CXD[q] + uy * CYD[q];
      double Fe = p->Fg[0] * CXD[q] + p->Fg[1] * CYD[q];
      double src = ((1 - 0.5 * s_nu[n]) * ((3.0 + 9.0 * ue) * Fe - 3.0 * uFg)) * W9[q]; /* :527-529 */
      r_col[n * 9 + q] = o3r + src;
      b_col[n * 9 + q] = o3b + src;
    }
    Fs[n * 2] = fsx;
    Fs[n * 2 + 1] = fsy;
  }
  orc_advect(r_col, X, Y, r_adv);
  orc_advect(b_col, X, Y, b_adv);
  mrtcg_bc(r_adv, r_col, X, Y);
  mrtcg_bc(b_adv, b_col, X, Y);
  orc_calc_rho(r_adv, X, Y, r_rho);
  orc_calc_rho(b_adv, X, Y, b_rho);
  for (size_t n = 0; n < N; n++)
  {
    rho[n] = r_rho[n] + b_rho[n];
    double sx = 0.0, sy = 0.0;
    for (int q = 0; q < 9; q++)
    {
      double t = r_adv[n * 9 + q] + b_adv[n * 9 + q];
      sx += t * CXD[q];
      sy += t * CYD[q];
    }
    /* :543-544: u = calc_u + 0.5 (Fg + interf_tension) / rho */
    u[n * 2] = sx / rho[n] + 0.5 * (p->Fg[0] + Fs[n * 2]) / rho[n];
    u[n * 2 + 1] = sy / rho[n] + 0.5 * (p->Fg[1] + Fs[n * 2 + 1]) / rho[n];
  }
  free(Qx); free(Qy); free(tmp); free(rDxQx); free(rDyQy); free(bDxQx); free(bDyQy);
  free(gx); free(gy); free(nx); free(ny); free(dx_nx); free(dy_nx); free(dx_ny); free(dy_ny); free(r_col); free(b_col);
}

/* ------------------------------------------------------------------ ulbm::d2q9::kbc */

/* the nine polynomial factors of kbc::eval_equilibrium / eval_iequilibrium (src/ulbm.cpp:230-240, 250-258) */
static void kbc_eq_coef(double ux, double uy, double ux2, double uy2, double* e)
{
  const double cs2 = 1.0 / 3.0, cs4 = 1.0 / 9.0;
  e[0] = 2.0 * cs2 * (0.5 * ux2 + 0.5 * uy2 - 1.0) + cs4 + ux2 * uy2 - ux2 - uy2 + 1.0;
  e[1] = 0.5 * (-cs2 * (ux2 + uy2 + ux - 1.0) - cs4 - ux2 * uy2 + ux2 - uy2 * ux + ux);
  e[2] = 0.5 * (-cs2 * (ux2 + uy2 + uy - 1.0) - cs4 - ux2 * uy2 - ux2 * uy + uy2 + uy);
  e[3] = 0.5 * (-cs2 * (ux2 + uy2 - ux - 1.0) - cs4 - ux2 * uy2 + ux2 + uy2 * ux - ux);
  e[4] = 0.5 * (-cs2 * (ux2 + uy2 - uy - 1.0) - cs4 - ux2 * uy2 + ux2 * uy + uy2 - uy);
  e[5] = 0.25 * (cs2 * (ux2 + uy2 + ux + uy) + cs4 + ux2 * uy2 + ux2 * uy + uy2 * ux + ux * uy);
  e[6] = 0.25 * (cs2 * (ux2 + uy2 - ux + uy) + cs4 + ux2 * uy2 + ux2 * uy - uy2 * ux - ux * uy);
  e[7] = 0.25 * (cs2 * (ux2 + uy2 - ux - uy) + cs4 + ux2 * uy2 - ux2 * uy - uy2 * ux + ux * uy);
  e[8] = 0.25 * (cs2 * (ux2 + uy2 + ux - uy) + cs4 + ux2 * uy2 - ux2 * uy + uy2 * ux - ux * uy);
}

/* kbc::eval_equilibrium reads the members ux2, uy2, which only collide() refreshes (eval_m1_components,
 * src/ulbm.cpp:150-155): called on a fresh object, as test/ulbm_double_shear_flow.cpp:97 does for its initial
 * state, it sees ux2 = uy2 = 0.  fresh_object != 0 reproduces that; 0 uses ux2 = ux^2, uy2 = uy^2. */
void orc_kbc_equilibrium(const double* m0, const double* m1, int X, int Y, int fresh_object, double* feq)
{
#pragma omp parallel for
  for (long n = 0; n < (long)X * Y; n++)
  {
    double e[9];
    const double ux = m1[2 * n], uy = m1[2 * n + 1];
    kbc_eq_coef(ux, uy, fresh_object ? 0.0 : ux * ux, fresh_object ? 0.0 : uy * uy, e);
    for (int q = 0; q < 9; q++) feq[n * 9 + q] = e[q] * m0[n];
  }
}

/* kbc::collide() of one node (src/ulbm.cpp:91-126 with eval_central_momenta :264-320, eval_s_matrix :128-136,
 * eval_gamma :138-148, eval_delta_s :157-189, eval_delta_h :191-224 incl. its `ux2+uy` terms, eval_iequilibrium
 * :226-244).  f = adve_f, out: coll = coll_f, iequi = iequi_f. */
static void kbc_collide_node(const double* f, double m0, double ux, double uy, double s2, double* coll, double* iequi)
{
  const double cs2 = 1.0 / 3.0, cs4 = 1.0 / 9.0, is2 = 1.0 / s2;
  double cmx[9], cmy[9], cmx2[9], cmy2[9], cT[9];
  for (int q = 0; q < 9; q++)
  {
    cmx[q] = CXI[q] == 0 ? -ux : (CXI[q] > 0 ? 1.0 - ux : -1.0 - ux);
    cmy[q] = CYI[q] == 0 ? -uy : (CYI[q] > 0 ? 1.0 - uy : -1.0 - uy);
    cmx2[q] = cmx[q] * cmx[q];
    cmy2[q] = cmy[q] * cmy[q];
  }
  for (int k = 0; k < 9; k++) cT[k] = 0.0;
  for (int q = 0; q < 9; q++)
  {
    cT[0] += f[q];
    cT[1] += f[q] * cmx[q];
    cT[2] += f[q] * cmy[q];
    cT[3] += f[q] * (cmx2[q] + cmy2[q]);
    cT[4] += f[q] * (cmx2[q] - cmy2[q]);
    cT[5] += f[q] * cmx[q] * cmy[q];
    cT[6] += f[q] * cmx2[q] * cmy[q];
    cT[7] += f[q] * cmx[q] * cmy2[q];
    cT[8] += f[q] * cmx2[q] * cmy2[q];
  }
  /* eval_gamma */
  const double ux2 = ux * ux, uy2 = uy * uy;
  const double C3 = cT[3], C4 = cT[4], C5 = cT[5], C6 = cT[6], C7 = cT[7], C8 = cT[8];
  double ds[9], dh[9], e[9];
  ds[0] = -0.5 * C4 * (ux2 - uy2) + 4.0 * C5 * ux * uy - cs4 * m0 - m0 * (ux2 * uy2 - ux2 - uy2 + 1) + (C3 - 2.0 * cs2 * m0) * (0.5 * ux2 + 0.5 * uy2 - 1.0);
  ds[1] = 0.25 * C4 * (ux2 - uy2 + ux + 1) - C5 * uy * (2.0 * ux + 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 + uy2 * ux - ux) - 0.25 * (C3 - 2.0 * cs2 * m0) * (ux2 + uy2 + ux - 1.0);
  ds[2] = -0.25 * C4 * (-ux2 + uy2 + uy + 1) - C5 * ux * (2.0 * uy + 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - uy2 + ux2 * uy - uy) - 0.25 * (C3 - 2.0 * cs2 * m0) * (ux2 + uy2 + uy - 1.0);
  ds[3] = 0.25 * C4 * (ux2 - uy2 - ux + 1) - C5 * uy * (2.0 * ux - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 - uy2 * ux + ux) - 0.25 * (C3 - 2.0 * cs2 * m0) * (ux2 + uy2 - ux - 1.0);
  ds[4] = 0.25 * C4 * (ux2 - uy2 + uy - 1) - C5 * ux * (2.0 * uy - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - uy2 - ux2 * uy + uy) - 0.25 * (C3 - 2.0 * cs2 * m0) * (ux2 + uy2 - uy - 1.0);
  ds[5] = -0.125 * C4 * (ux2 - uy2 + ux - uy) + C5 * (ux * uy + 0.5 * ux + 0.5 * uy + 0.25) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 + ux2 * uy + uy2 * ux + ux * uy) + 0.125 * (C3 - 2.0 * cs2 * m0) * (ux2 + uy2 + ux + uy);
  ds[6] = 0.125 * C4 * (-ux2 + uy2 + ux + uy) + C5 * (ux * uy + 0.5 * ux - 0.5 * uy - 0.25) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 + ux2 * uy - uy2 * ux - ux * uy) + 0.125 * (C3 - 2.0 * cs2 * m0) * (ux2 + uy2 - ux + uy);
  ds[7] = -0.125 * C4 * (ux2 - uy2 - ux + uy) + C5 * (ux * uy - 0.5 * ux - 0.5 * uy + 0.25) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 - ux2 * uy - uy2 * ux + ux * uy) + 0.125 * (C3 - 2.0 * cs2 * m0) * (ux2 + uy2 - ux - uy);
  ds[8] = -0.125 * C4 * (ux2 - uy2 + ux + uy) + C5 * (ux * uy - 0.5 * ux + 0.5 * uy - 0.25) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 - ux2 * uy + uy2 * ux - ux * uy) + 0.125 * (C3 - 2.0 * cs2 * m0) * (ux2 + uy2 + ux - uy);
  dh[0] = 2.0 * C6 * uy + 2.0 * C7 * ux + C8 - 2.0 * cs2 * m0 * (0.5 * ux2 + 0.5 * uy2 - 1.0) - cs4 * m0 - m0 * (ux2 * uy2 - ux2 - uy2 + 1.0);
  dh[1] = -C6 * uy - C7 * (ux + 0.5) - 0.5 * C8 + 0.5 * cs2 * m0 * (ux2 + uy2 + ux - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 + uy2 * ux - ux);
  dh[2] = -C6 * (uy + 0.5) - C7 * ux - 0.5 * C8 + 0.5 * cs2 * m0 * (ux2 + uy2 + uy - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 + ux2 * uy - uy2 - uy);
  dh[3] = -C6 * uy - C7 * (ux - 0.5) - 0.5 * C8 + 0.5 * cs2 * m0 * (ux2 + uy2 - ux - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 - uy2 * ux + ux);
  dh[4] = -C6 * (uy - 0.5) - C7 * ux - 0.5 * C8 + 0.5 * cs2 * m0 * (ux2 + uy2 - uy - 1.0) + 0.5 * cs4 * m0 + 0.5 * m0 * (ux2 * uy2 - ux2 * uy - uy2 + uy);
  /* :211-223 as written: `ux2+uy`, not `ux2*uy` */
  dh[5] = C6 * (0.5 * uy + 0.25) + C7 * (0.5 * ux + 0.25) + 0.25 * C8 - 0.25 * cs2 * m0 * (ux2 + uy2 + ux + uy) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 + ux2 + uy + uy2 * ux + ux * uy);
  dh[6] = C6 * (0.5 * uy + 0.25) + C7 * (0.5 * ux - 0.25) + 0.25 * C8 - 0.25 * cs2 * m0 * (ux2 + uy2 - ux + uy) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 + ux2 + uy - uy2 * ux - ux * uy);
  dh[7] = C6 * (0.5 * uy - 0.25) + C7 * (0.5 * ux - 0.25) + 0.25 * C8 - 0.25 * cs2 * m0 * (ux2 + uy2 - ux - uy) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 - ux2 + uy - uy2 * ux + ux * uy);
  dh[8] = C6 * (0.5 * uy - 0.25) + C7 * (0.5 * ux + 0.25) + 0.25 * C8 - 0.25 * cs2 * m0 * (ux2 + uy2 + ux - uy) - 0.25 * cs4 * m0 - 0.25 * m0 * (ux2 * uy2 - ux2 + uy + uy2 * ux - ux * uy);
  kbc_eq_coef(ux, uy, ux2, uy2, e);
  double num = 0.0, den = 0.0;
  for (int q = 0; q < 9; q++)
  {
    iequi[q] = 1.0 / (e[q] * m0);
    num += ds[q] * dh[q] * iequi[q];
    den += dh[q] * dh[q] * iequi[q];
  }
  const double gamma = is2 - (1.0 - is2) * num / den;
  const double S[9] = {1.0, 1.0, 1.0, s2, s2, s2, gamma * s2, gamma * s2, gamma * s2};
  /* collide() steps 1-5 */
  cT[0] += -m0;
  cT[3] += -2.0 * cs2 * m0;
  cT[8] += -cs4 * m0;
  for (int k = 0; k < 9; k++) cT[k] *= S[k];
  double g[9];
  g[0] = cT[0];
  g[1] = cT[0] * ux + cT[1];
  g[2] = cT[0] * uy + cT[2];
  g[3] = cT[0] * (ux2 + uy2) + 2.0 * cT[1] * ux + 2.0 * cT[2] * uy + cT[3];
  g[4] = cT[0] * (ux2 - uy2) + 2.0 * cT[1] * ux - 2.0 * cT[2] * uy + cT[4];
  g[5] = cT[0] * ux * uy + cT[1] * uy + cT[2] * ux + cT[5];
  g[6] = cT[0] * ux2 * uy + 2.0 * cT[1] * ux * uy + cT[2] * ux2 + 0.5 * cT[3] * uy + 0.5 * cT[4] * uy + 2.0 * cT[5] * ux + cT[6];
  g[7] = cT[0] * ux * uy2 + cT[1] * uy2 + 2.0 * cT[2] * ux * uy + 0.5 * cT[3] * ux - 0.5 * cT[4] * ux + 2.0 * cT[5] * uy + cT[7];
  g[8] = cT[0] * ux2 * uy2 + 2.0 * cT[1] * ux * uy2 + 2.0 * cT[2] * ux2 * uy + 0.5 * cT[3] * (ux2 + uy2) - 0.5 * cT[4] * (ux2 - uy2) + 4.0 * cT[5] * ux * uy + 2.0 * cT[6] * uy + 2.0 * cT[7] * ux + cT[8];
  double c[9];
  c[0] = g[0] - g[3] + g[8];
  c[1] = 0.5 * g[1] + 0.25 * g[3] + 0.25 * g[4] - 0.5 * g[7] - 0.5 * g[8];
  c[2] = 0.5 * g[2] + 0.25 * g[3] - 0.25 * g[4] - 0.5 * g[6] - 0.5 * g[8];
  c[3] = -0.5 * g[1] + 0.25 * g[3] + 0.25 * g[4] + 0.5 * g[7] - 0.5 * g[8];
  c[4] = -0.5 * g[2] + 0.25 * g[3] - 0.25 * g[4] + 0.5 * g[6] - 0.5 * g[8];
  c[5] = 0.25 * (g[5] + g[6] + g[7] + g[8]);
  c[6] = 0.25 * (-g[5] + g[6] - g[7] + g[8]);
  c[7] = 0.25 * (g[5] - g[6] - g[7] + g[8]);
  c[8] = 0.25 * (-g[5] - g[6] + g[7] + g[8]);
  for (int q = 0; q < 9; q++) coll[q] = c[q] * -1.0 + f[q];
}

void orc_kbc_step(double* f, double* m0, double* m1, int X, int Y, double s2, int bc, double rho_in, double rho_out)
{
  size_t N = (size_t)X * Y;
  double* fcoll = dalloc(N * 9);
  double* iequi = dalloc(N * 9);
#pragma omp parallel for
  for (long n = 0; n < (long)N; n++) kbc_collide_node(f + n * 9, m0[n], m1[2 * n], m1[2 * n + 1], s2, fcoll + n * 9, iequi + n * 9);
  if (bc == 1)
  {
    /* periodic_boundary_condition(coll_f, iequi_f.pow(-1), m1, m0, rho_inlet, rho_outlet) (ulbm_poiseuille.cpp:39-60,117):
     * virtual inlet row 0 from row -2, virtual outlet row -1 from row 1, with solver::incomp_equilibrium */
    for (int side = 0; side < 2; side++)
    {
      const int dst = side == 0 ? 0 : X - 1, src = side == 0 ? X - 2 : 1;
      const double rbc = side == 0 ? rho_in : rho_out;
      for (int y = 0; y < Y; y++)
      {
        const double ux = m1[2 * N2(src, y)], uy = m1[2 * N2(src, y) + 1];
        for (int q = 0; q < 9; q++)
        {
          const double te = (rbc * 1.0 + 3.0 * (ux * CXD[q] + uy * CYD[q])) * W9[q];
          fcoll[IDX(dst, y, q)] = (te + fcoll[IDX(src, y, q)]) - 1.0 / iequi[IDX(src, y, q)];
        }
      }
    }
  }
  orc_advect(fcoll, X, Y, f);
  if (bc == 1)
  {
    for (int x = 0; x < X; x++)
    {
      f[IDX(x, Y - 1, 4)] = fcoll[IDX(x, Y - 1, 2)];
      f[IDX(x, Y - 1, 7)] = fcoll[IDX(x, Y - 1, 5)];
      f[IDX(x, Y - 1, 8)] = fcoll[IDX(x, Y - 1, 6)];
      f[IDX(x, 0, 2)] = fcoll[IDX(x, 0, 4)];
      f[IDX(x, 0, 5)] = fcoll[IDX(x, 0, 7)];
      f[IDX(x, 0, 6)] = fcoll[IDX(x, 0, 8)];
    }
  }
#pragma omp parallel for
  for (long n = 0; n < (long)N; n++)
  {
    double s = 0.0, jx = 0.0, jy = 0.0;
    for (int q = 0; q < 9; q++)
    {
      s += f[n * 9 + q];
      jx += f[n * 9 + q] * CXD[q];
      jy += f[n * 9 + q] * CYD[q];
    }
    m0[n] = s;
    m1[2 * n] = jx / s;
    m1[2 * n + 1] = jy / s;
  }
  free(fcoll);
  free(iequi);
}
