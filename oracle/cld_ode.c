/*
 * ORACLE (test infrastructure, NOT product code).  Parity status: "parity unpinned" by the
 * reference (its tests hold no numeric fixtures); pinned here by the analytic identities in
 * tests/test_oracle_cld.py (R R^T = Sigma_t from an independent Lyapunov integration, expm, ...).
 *
 * Plain-C restatement of the R(t) integrator of the reference:
 *   cld_jax/sde_lib.py:93-118  CLD._get_s_R_fn  (lax.scan over ts, carry emitted BEFORE the update)
 *   cld_jax/deis.py:5-17       runge_kutta       (classic RK4, used when is_R_rk)
 *   cld_jax/sde_lib.py:214-234 s_F, s_G
 *   cld_jax/sde_lib.py:17-26   inv_2x2
 * Everything in fp64 (the authors ran it in fp32; see SURVEY.md 8c "irreducible mismatch").
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libcld_ode.so oracle/cld_ode.c   (oracle/Makefile)
 */
#include <math.h>
#include <stddef.h>

typedef struct { double m_inv, beta_0, beta_1, Gamma; } cld_par;

static double beta_at(const cld_par* p, double t) { return p->beta_0 + p->beta_1 * t; }

/* grad = F R + 0.5 G G^T R^{-T}   (sde_lib.py:94-97) ; R stored row-major a b c d */
static void ode_rhs(const cld_par* p, const double R[4], double t, double out[4]) {
  double b = beta_at(p, t);
  double F00 = 0.0, F01 = b * p->m_inv, F10 = -b, F11 = -p->Gamma * b * p->m_inv;
  double g2 = 2.0 * p->Gamma * b;                 /* (G G^T)[1][1], G = diag(0, sqrt(2 Gamma beta)) */
  double det = R[0] * R[3] - R[1] * R[2];
  /* inv(R)^T = 1/det [[d, -c], [-b, a]] */
  double iT10 = -R[1] / det, iT11 = R[0] / det;
  out[0] = F00 * R[0] + F01 * R[2];
  out[1] = F00 * R[1] + F01 * R[3];
  out[2] = F10 * R[0] + F11 * R[2] + 0.5 * g2 * iT10;
  out[3] = F10 * R[1] + F11 * R[3] + 0.5 * g2 * iT11;
}

static void rk4_step(const cld_par* p, double R[4], double t, double dt) {
  double k1[4], k2[4], k3[4], k4[4], x[4];
  int i;
  ode_rhs(p, R, t, k1);
  for (i = 0; i < 4; ++i) x[i] = R[i] + k1[i] * dt / 2;
  ode_rhs(p, x, t + dt / 2, k2);
  for (i = 0; i < 4; ++i) x[i] = R[i] + k2[i] * dt / 2;
  ode_rhs(p, x, t + dt / 2, k3);
  for (i = 0; i < 4; ++i) x[i] = R[i] + k3[i] * dt;
  ode_rhs(p, x, t + dt, k4);
  for (i = 0; i < 4; ++i) R[i] = R[i] + dt / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

/* "mid point integral" branch, sde_lib.py:103-107: averages F and G (not G G^T) at t and t+dt */
static void euler_mid_step(const cld_par* p, double R[4], double t, double dt) {
  double b0 = beta_at(p, t), b1 = beta_at(p, t + dt);
  double F01 = 0.5 * (b0 + b1) * p->m_inv, F10 = -0.5 * (b0 + b1), F11 = -p->Gamma * 0.5 * (b0 + b1) * p->m_inv;
  double g = 0.5 * (sqrt(2.0 * p->Gamma * b0) + sqrt(2.0 * p->Gamma * b1));
  double g2 = g * g;
  double det = R[0] * R[3] - R[1] * R[2];
  double iT10 = -R[1] / det, iT11 = R[0] / det;
  double n0 = R[0] + dt * (F01 * R[2]);
  double n1 = R[1] + dt * (F01 * R[3]);
  double n2 = R[2] + dt * (F10 * R[0] + F11 * R[2] + 0.5 * g2 * iT10);
  double n3 = R[3] + dt * (F10 * R[1] + F11 * R[3] + 0.5 * g2 * iT11);
  R[0] = n0; R[1] = n1; R[2] = n2; R[3] = n3;
}

/*
 * Emits Rs[k] = R(ts[k]) for k = 0..n_ts-1 where the carry is emitted before each update
 * (lax.scan returns `carry` as the per-step output, sde_lib.py:100,107).  The integrator step is
 * R_dt (sde_lib.py:100,106), independent of the actual spacing of ts.
 */
void oracle_cld_scan_R(double m_inv, double beta_0, double beta_1, double Gamma,
                       const double R0[4], const double* ts, size_t n_ts, double R_dt,
                       int is_rk, double* Rs_out) {
  cld_par p = { m_inv, beta_0, beta_1, Gamma };
  double R[4] = { R0[0], R0[1], R0[2], R0[3] };
  size_t k; int i;
  for (k = 0; k < n_ts; ++k) {
    for (i = 0; i < 4; ++i) Rs_out[4 * k + i] = R[i];
    if (is_rk) rk4_step(&p, R, ts[k], R_dt); else euler_mid_step(&p, R, ts[k], R_dt);
  }
}
