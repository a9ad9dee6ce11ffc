/* TEST INFRASTRUCTURE — CPU restatement of the reference's D2Q9 time step.
 *
 * This is the parity ORACLE: plain C, fp64, reference layout (AoS {X,Y,9}
 * row-major: axis 0 = x "rows", axis 1 = y "columns", 9 populations contiguous
 * per node).  Each function cites the reference file:line it restates
 * (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline/--impl reference legs may load it; the product
 * (lattice-boltzmann-method_b200/) never does.
 *
 * Parity status: PINNED — tests/test_oracle_vs_reference.py checks every
 * function here against the unmodified reference sources compiled into
 * oracle/_ref (libref_harness.so + the reference's own driver binaries), and
 * tests/golden/ holds outputs of those reference binaries.
 */
#ifndef LBM_ORACLE_H
#define LBM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants: src/solver.cpp:12-21 */
int orc_num_threads(int n);  /* OpenMP threads of the port (n > 0: set first) */
void orc_constants(double* w9, double* c18);

/* ---- granular ops: src/solver.cpp:23-131 */
void orc_calc_rho(const double* f, int X, int Y, double* rho);
void orc_calc_u(const double* f, const double* rho, int X, int Y, double* u);
void orc_calc_incomp_u(const double* f, int X, int Y, double* u);
void orc_equilibrium(const double* u, const double* rho, int X, int Y, double* feq);
void orc_incomp_equilibrium(const double* u, const double* rho, int X, int Y, double* feq);
void orc_collision(const double* f, const double* feq, double omega, int X, int Y, double* fcoll);
void orc_advect(const double* f, int X, int Y, double* g);

/* ---- finite differences
 * orc_diff5: src/differential.hpp:9-40, src/differential.cpp:3-39  (5x5, replicate pad;
 *            dx = derivative along axis 0, dy = along axis 1)
 * orc_diff3: test/rk_static_droplet_test.cpp:48-107 (3x3, replicate pad; NOTE the driver's
 *            "x" kernel differentiates along axis 1 and its "y" kernel along axis 0) */
void orc_diff5(const double* psi, int R, int C, double* dx, double* dy);
void orc_diff3(const double* psi, int R, int C, double* dx, double* dy);

/* ---- host parameter derivations
 * orc_params_lattice: src/params.cpp:7-66; in = {rho0, nu, u, l, tau, dx, x_mult, y_mult},
 *   out = {Re, omega, nu_lb, l, dt, T, u_lb, X, Y}
 * orc_params_simulation: src/params.cpp:95-112; out = {total_steps, snapshot_steps, total_snapshots}
 * orc_colour_params: src/colour.cpp:37-64; out = {mu, cs2, ics2, rlx, phi[9], eta[9]} */
void orc_params_lattice(const double* in8, double* out9);
void orc_params_simulation(double stop_time, double snapshot_period, int T, double* out3);
void orc_colour_params(double rho_0, double alpha, double nu, double* out22);

/* ---- driver 10: test/horizontal_poiseuille_test.cpp:128-152 (one loop iteration)
 * incompressible BGK, pressure-periodic rows, half-way bounce-back columns.
 * f/u/rho are in-out exactly like the driver's f_adve/u/rho. */
void orc_poiseuille_step(double* f, double* u, double* rho, int X, int Y, double omega,
                         double rho_in, double rho_out);
/* ---- driver 13: test/specular_boundary_test.cpp:103-128 */
void orc_specular_step(double* f, double* u, double* rho, int X, int Y, double omega,
                       double rho_in, double rho_out);
/* ---- driver 14: test/gravity_test.cpp:139-183 (Fg = {Fx, Fy}) */
void orc_gravity_step(double* f, double* u, double* rho, int X, int Y, double omega,
                      double rho_in, double rho_out, const double* Fg);
/* ---- driver 12: test/free_stream_test.cpp:88-134 (u_w = {uwx, 0}) */
void orc_free_stream_step(double* f, double* u, double* rho, int X, int Y, double omega, double uwx);
/* ---- driver 19: test/decompose_domain.cpp:127-187 (domains A over B, both {X,Y}) */
void orc_decompose_step(double* fA, double* uA, double* rhoA, double* fB, double* uB, double* rhoB,
                        int X, int Y, double omega, double rho_in, double rho_out);

/* ---- immersed boundary: src/ibm.cpp:15-190 */
typedef struct orc_ibm orc_ibm;
orc_ibm* orc_ibm_create(const double* xs, const double* ys, int n_markers, int m_max);
void orc_ibm_destroy(orc_ibm* ib);
void orc_ibm_roi(const orc_ibm* ib, long* roi4);
/* F_out: {roi_rows, roi_cols, 2} */
void orc_ibm_force(orc_ibm* ib, const double* u, const double* rho, int X, int Y, double* F_out);

/* ---- driver 11: test/cylinder_test.cpp:100-163; F_out {roi_rows,roi_cols,2} may be NULL */
void orc_cylinder_step(double* f, double* u, double* rho, int X, int Y, double omega, double u_lb,
                       orc_ibm* ib, double* F_out);

/* ---- driver 15: test/rectangle_sedimentation_test.cpp:110-238
 * state: f, g (populations), u, rho, C; C_w {X}; walls R23 (negative, from the end), C28, C38 */
void orc_sedimentation_init(double* f, double* g, double* u, double* rho, double* C, int X, int Y,
                            double u_lb, const double* C_w);
void orc_sedimentation_step(double* f, double* g, double* u, double* rho, double* C, int X, int Y,
                            double omega, double u_lb, double w_s, const double* C_w,
                            int R23, int C28, int C38);
/* the same loop with an immersed body coupled as in test/cylinder_test.cpp:110-127 (BASELINE configs[4] as worded;
 * not a reference driver: composed of the two pinned steps, itself unpinned) */
void orc_sedimentation_ibm_step(double* f, double* g, double* u, double* rho, double* C, int X, int Y,
                                double omega, double u_lb, double w_s, const double* C_w, int R23, int C28,
                                int C38, orc_ibm* ib);

/* ---- drivers 16/18: test/mrtcg_rayleigh_taylor.cpp:413-478, test/mrtcg_static_droplet.cpp:463-528 */
typedef struct
{
  int R, C;
  double r_rho0, r_alpha, r_nu, r_beta;
  double b_rho0, b_alpha, b_nu, b_beta;
  double sigma;     /* RT: [general].sigma ; droplet: 0.1 */
  double delta;     /* 0.1, hard-coded in both drivers */
  double Fg[2];     /* RT: {g, 0} ; droplet: {0, -6.25e-6} */
  int add_force;    /* RT: 1 ; droplet: 0 (source computed but not added) */
} orc_mrtcg_params;
void orc_mrtcg_init_rt(const orc_mrtcg_params* p, double* r_rho, double* b_rho);       /* :182-210 */
void orc_mrtcg_init_droplet(const orc_mrtcg_params* p, double* r_rho, double* b_rho);  /* droplet :182-204 */
/* builds rho, u, adv_f from the initial densities (RT :407-410 ; droplet :455-459 incl. the u shift) */
void orc_mrtcg_init_state(const orc_mrtcg_params* p, const double* r_rho, const double* b_rho,
                          double* rho, double* u, double* r_adv, double* b_adv, int shift_u);
/* one loop iteration; phase, s_nu, grad are outputs (s_nu is in-out: NaN phase keeps the old value) */
void orc_mrtcg_step(const orc_mrtcg_params* p, double* r_adv, double* b_adv, double* r_rho,
                    double* b_rho, double* rho, double* u, double* phase, double* s_nu, double* grad);

/* ---- driver 17: test/rk_static_droplet_test.cpp:544-615 */
typedef struct
{
  int L;                 /* grid is L x L */
  double radius;         /* 25.0 */
  double r_rho0, r_alpha, r_A, r_nu;
  double b_rho0, b_alpha, b_A, b_nu;
  double delta;          /* 0.98 */
} orc_rk_params;
void orc_rk_init(const orc_rk_params* p, const double* u0, double* r_adv, double* b_adv,
                 double* r_rho, double* b_rho, double* rho_mix);
void orc_rk_step(const orc_rk_params* p, double* r_adv, double* b_adv, double* r_rho, double* b_rho,
                 double* rho_mix, double* u, double* phase, double* relax, double* grad);
/* the diagnostic fields the driver snapshots per iteration (:546-600), from the state at the top of the iteration;
 * relax_io is in-out like orc_rk_step's relax; omega1 / omega2 are the red colour's; NULL outputs are skipped */
void orc_rk_diagnostics(const orc_rk_params* p, double sigma, const double* r_adv, const double* r_rho, const double* b_rho,
                        const double* rho_mix, const double* u, double* phase_o, double* grad_o, double* norm_o, double* n_o,
                        double* K_o, double* Fs_o, double* eta_o, double* kappa_o, double* relax_io, double* omega1_o,
                        double* omega2_o);

/* ---- test/mrt_rayleigh_taylor.cpp: the MRT colour-gradient model with a continuum-surface-force perturbation
 * (curvature from a second pass of `differential`) -- SURVEY 8(f) rank 2.  The driver only runs at
 * 1024 x 256 (E_rep, :180); pinned against its own snapshots (tests/golden/mrt_csf_1024x256.npz). */
typedef struct
{
  int R, C;
  double r_rho0, r_alpha, r_nu, r_beta, r_A;
  double b_rho0, b_alpha, b_nu, b_beta, b_A;
  double sigma, delta, Fg[2];
} orc_csf_params;
/* init_rho_cosine (:184-210, interface at R/2 + 0.1 C cos), u = 0.5 Fg / rho_r0 (:464), adv_f = equilibrium */
void orc_csf_init(const orc_csf_params* p, double* r_rho, double* b_rho, double* rho, double* u, double* r_adv, double* b_adv);
/* one loop iteration (:490-545); Fs {R,C,2} = interf_tension carried into the next u, s_nu {R,C} */
void orc_csf_step(const orc_csf_params* p, double* r_adv, double* b_adv, double* r_rho, double* b_rho, double* rho,
                  double* u, double* phase, double* s_nu, double* Fs);

/* ---- ulbm::d2q9::kbc (src/ulbm.hpp:11-90, src/ulbm.cpp) and its two drivers -- SURVEY 8(f) rank 3.
 * Pinned against the compiled reference: ref_kbc_run / ref_kbc_equilibrium of oracle/ref_harness.cpp
 * (tests/test_oracle_vs_reference.py) and the snapshots of test/ulbm_double_shear_flow.cpp
 * (tests/golden/kbc_double_shear_128.npz). */
/* kbc::eval_equilibrium (src/ulbm.cpp:246-262): m0 {X,Y}, m1 {X,Y,2} -> feq {X,Y,9}.  fresh_object != 0: the members
 * ux2, uy2 are still zero, as in the driver's initialisation (test/ulbm_double_shear_flow.cpp:97) */
void orc_kbc_equilibrium(const double* m0, const double* m1, int X, int Y, int fresh_object, double* feq);
/* one loop iteration: collide(); [bc = 1: pressure rows on coll_f]; advect(); [bc = 1: bounce-back columns];
 * m0 = sum f, m1 = f c / m0.  bc = 0: test/ulbm_double_shear_flow.cpp:115-142 (fully periodic),
 * bc = 1: test/ulbm_poiseuille.cpp:116-143.  f = adve_f {X,Y,9}, m0 {X,Y}, m1 {X,Y,2} in/out. */
void orc_kbc_step(double* f, double* m0, double* m1, int X, int Y, double s2, int bc, double rho_in, double rho_out);

#ifdef __cplusplus
}
#endif
#endif
