// fp64 host tables.  Follows (relative to /root/reference):
//   cld_jax/sampling.py:241-249      get_rev_ts
//   cld_jax/sde_lib.py:32-43,93-118  R(t): ODE scan (RK4 or midpoint Euler) + 100 000-node interpolation table
//   cld_jax/sde_lib.py:182-234       Psi, eps integrand, F, G
//   cld_jax/sde_lib.py:289-319       prepare_order0_coef, get_deis_coef
//   cld_jax/deis.py:19-95            DEIS Adams-Bashforth coefficient tables (10 000-node left Riemann sums)
//   cld_jax/sampling.py:30-39        denoising step
//   blur_jax/sde_lib.py:18-97        blur schedule;  blur_jax/sampling.py:60-75 order-0 update
// The authors integrate in fp32 and cache pickles; here everything is fp64 and recomputed (sub-second).
#include "tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace gddim {

static Mat2 rk4(const Mat2& x, double t, double dt, const std::function<Mat2(const Mat2&, double)>& fn);

void rev_timesteps(double T, double eps, int ts_order, int num_step, double* out) {
  const double a = std::pow(T, 1.0 / ts_order), b = std::pow(eps, 1.0 / ts_order);
  // numpy.linspace(a, b, n+1): a + k*step, last element set to b
  const double step = (b - a) / num_step;
  for (int k = 0; k <= num_step; ++k) {
    const double v = (k == num_step) ? b : a + k * step;
    out[k] = std::pow(v, (double)ts_order);
  }
}

CldTables::CldTables(double m_inv_, double beta_0_, double beta_1_, double vv_gamma, double numerical_eps,
                     double R_dt_, bool is_rk_)
    : m_inv(m_inv_), beta_0(beta_0_), beta_1(beta_1_), Gamma(2.0 / std::sqrt(m_inv_)), R_dt(R_dt_), is_rk(is_rk_) {
  R0 = {std::sqrt(numerical_eps), 0.0, 0.0, std::sqrt(vv_gamma / m_inv + numerical_eps)};
  const long n = (long)(1.0 / R_dt);          // 99999 for R_dt = 1e-5 (floating point), 1000000 for 1e-6
  const long len = n + 1;
  const double tstep = (1.0 + R_dt) / (double)len;   // linspace(0, 1 + R_dt, len, endpoint=False)
  // subsample indices: linspace(0, len-1, 100000).astype(int)
  const long nsub = 100000;
  std::vector<long> idx(nsub);
  const double istep = (double)(len - 1) / (double)(nsub - 1);
  for (long k = 0; k < nsub; ++k) idx[k] = (k == nsub - 1) ? len - 1 : (long)(k * istep);
  xp_.resize(nsub);
  fp_.resize(nsub);
  Mat2 R = R0;
  long next = 0;
  for (long k = 0; k < len; ++k) {
    const double t = k * tstep;
    while (next < nsub && idx[next] == k) {   // the scan emits the carry *before* the update
      xp_[next] = t;
      fp_[next] = R;
      ++next;
    }
    if (next >= nsub) break;
    if (is_rk) {
      const double dt = R_dt;
      const Mat2 g1 = ode_rhs(R, t);
      const Mat2 x2 = {R.a + g1.a * dt / 2, R.b + g1.b * dt / 2, R.c + g1.c * dt / 2, R.d + g1.d * dt / 2};
      const Mat2 g2 = ode_rhs(x2, t + dt / 2);
      const Mat2 x3 = {R.a + g2.a * dt / 2, R.b + g2.b * dt / 2, R.c + g2.c * dt / 2, R.d + g2.d * dt / 2};
      const Mat2 g3 = ode_rhs(x3, t + dt / 2);
      const Mat2 x4 = {R.a + g3.a * dt, R.b + g3.b * dt, R.c + g3.c * dt, R.d + g3.d * dt};
      const Mat2 g4 = ode_rhs(x4, t + dt);
      R = {R.a + dt / 6 * (g1.a + 2 * g2.a + 2 * g3.a + g4.a), R.b + dt / 6 * (g1.b + 2 * g2.b + 2 * g3.b + g4.b),
           R.c + dt / 6 * (g1.c + 2 * g2.c + 2 * g3.c + g4.c), R.d + dt / 6 * (g1.d + 2 * g2.d + 2 * g3.d + g4.d)};
    } else {
      // "mid point integral": F and G (not G G^T) averaged between t and t + dt
      const Mat2 f0 = F(t), f1 = F(t + R_dt), g0 = G(t), g1 = G(t + R_dt);
      const Mat2 Fm = {(f0.a + f1.a) / 2, (f0.b + f1.b) / 2, (f0.c + f1.c) / 2, (f0.d + f1.d) / 2};
      const Mat2 Gm = {0, 0, 0, (g0.d + g1.d) / 2};
      const Mat2 t1 = mul(Fm, R);
      const Mat2 t2 = mul(mul(Gm, Gm), tr(inv(R)));
      R = {R.a + R_dt * (t1.a + 0.5 * t2.a), R.b + R_dt * (t1.b + 0.5 * t2.b), R.c + R_dt * (t1.c + 0.5 * t2.c),
           R.d + R_dt * (t1.d + 0.5 * t2.d)};
    }
  }
}

Mat2 CldTables::F(double t) const {
  const double b = beta(t);
  return {0.0, b * m_inv, -b, -Gamma * b * m_inv};
}
Mat2 CldTables::G(double t) const { return {0.0, 0.0, 0.0, std::sqrt(2.0 * Gamma * beta(t))}; }

Mat2 CldTables::ode_rhs(const Mat2& R, double t) const {
  const Mat2 g = G(t);
  const Mat2 t1 = mul(F(t), R);
  const Mat2 t2 = mul(mul(g, tr(g)), tr(inv(R)));
  return {t1.a + 0.5 * t2.a, t1.b + 0.5 * t2.b, t1.c + 0.5 * t2.c, t1.d + 0.5 * t2.d};
}

Mat2 CldTables::R(double t) const {
  // i = clip(searchsorted(xp, t, side='right'), 1, len-1)
  long i = std::upper_bound(xp_.begin(), xp_.end(), t) - xp_.begin();
  const long n = (long)xp_.size();
  i = std::min(std::max(i, 1L), n - 1);
  const double dx = xp_[i] - xp_[i - 1];
  if (dx == 0) return fp_[i];
  const double w = (t - xp_[i - 1]) / dx;
  const Mat2 &p = fp_[i - 1], &q = fp_[i];
  return {p.a + w * (q.a - p.a), p.b + w * (q.b - p.b), p.c + w * (q.c - p.c), p.d + w * (q.d - p.d)};
}

Mat2 CldTables::psi(double s, double t) const {
  const double B = beta_int(t) - beta_int(s);
  const double a = 2.0 * std::sqrt(m_inv);
  const double k = std::exp(-a * B / 2);
  return {(1 + a * B / 2) * k, 0.25 * a * a * B * k, -B * k, (1 - a * B / 2) * k};
}

Mat2 CldTables::eps_integrand(double t) const {
  const Mat2 g = G(t);
  const Mat2 m = mul(mul(g, g), tr(inv(R(t))));
  return {0.5 * m.a, 0.5 * m.b, 0.5 * m.c, 0.5 * m.d};
}

// ---- generic DEIS table (deis.py:19-95) -----------------------------------------------------------------------
static Mat2 quad_generic(const PsiFn& psi, const IntegrandFn& integrand, double t_start, double t_end,
                         const double* ts_poly, int n_poly, int coef_idx, int num_item) {
  const double dt = (t_end - t_start) / num_item;
  Mat2 acc = {0, 0, 0, 0};
  for (int k = 0; k < num_item; ++k) {
    const double tau = t_start + k * dt;
    double num = 1.0, den = 1.0;
    for (int q = 0; q < n_poly; ++q)
      if (q != coef_idx) {
        num *= tau - ts_poly[q];
        den *= ts_poly[coef_idx] - ts_poly[q];
      }
    const double w = num / den;
    const Mat2 m = mul(psi(tau, t_end), integrand(tau));
    acc.a += m.a * w; acc.b += m.b * w; acc.c += m.c * w; acc.d += m.d * w;
  }
  return {acc.a * dt, acc.b * dt, acc.c * dt, acc.d * dt};
}

static void coef_row_generic(const PsiFn& psi, const IntegrandFn& integrand, int highest_order, int order,
                             double t_start, double t_end, const double* ts_poly, double* out) {
  std::memset(out, 0, sizeof(double) * (highest_order + 1) * 4);
  for (int j = 0; j <= order; ++j) {
    const Mat2 m = quad_generic(psi, integrand, t_start, t_end, ts_poly, order + 1, order - j, 10000);
    out[j * 4 + 0] = m.a; out[j * 4 + 1] = m.b; out[j * 4 + 2] = m.c; out[j * 4 + 3] = m.d;
  }
}

void deis_ab_eps_coef(const PsiFn& psi, const IntegrandFn& integrand, int highest_order, const double* ts, int n_ts,
                      int order, std::vector<double>& out) {
  const int stride = (highest_order + 1) * 4;
  if (order == 0) {
    for (int i = 0; i + 1 < n_ts; ++i) {
      out.resize(out.size() + stride);
      coef_row_generic(psi, integrand, highest_order, 0, ts[i], ts[i + 1], ts + i, out.data() + out.size() - stride);
    }
    return;
  }
  deis_ab_eps_coef(psi, integrand, highest_order, ts, order + 1, order - 1, out);
  for (int k = 0; k < n_ts - order - 1; ++k) {
    out.resize(out.size() + stride);
    coef_row_generic(psi, integrand, highest_order, order, ts[order + k], ts[order + k + 1], ts + k,
                     out.data() + out.size() - stride);
  }
}

Mat2 mvn_factor_svd(const Mat2& m) {
  // M M^T = U S^2 U^T (symmetric 2x2 eigen-decomposition), singular values descending like LAPACK
  const double p = m.a * m.a + m.b * m.b, q = m.a * m.c + m.b * m.d, r = m.c * m.c + m.d * m.d;
  const double tr2 = 0.5 * (p + r), df = 0.5 * (p - r);
  const double rad = std::sqrt(df * df + q * q);
  const double l1 = tr2 + rad, l2 = std::max(tr2 - rad, 0.0);
  double u1x, u1y;
  if (rad < 1e-300) { u1x = 1.0; u1y = 0.0; }
  else if (df >= 0) { u1x = df + rad; u1y = q; }
  else { u1x = q; u1y = rad - df; }
  const double n1 = std::sqrt(u1x * u1x + u1y * u1y);
  if (n1 > 0) { u1x /= n1; u1y /= n1; } else { u1x = 1.0; u1y = 0.0; }
  double u2x = -u1y, u2y = u1x;
  if ((std::fabs(u1x) >= std::fabs(u1y) ? u1x : u1y) < 0) { u1x = -u1x; u1y = -u1y; }
  if ((std::fabs(u2x) >= std::fabs(u2y) ? u2x : u2y) < 0) { u2x = -u2x; u2y = -u2y; }
  const double s1 = std::sqrt(std::sqrt(l1)), s2 = std::sqrt(std::sqrt(l2));   // sqrt of the singular values
  return {u1x * s1, u2x * s2, u1y * s1, u2y * s2};
}

// ---- LambdaSDE (sde_lib.py:334-466) -------------------------------------------------------------------------------
static Mat2 rk4(const Mat2& x, double t, double dt, const std::function<Mat2(const Mat2&, double)>& fn) {
  auto axpy = [](const Mat2& a, const Mat2& g, double s) { return Mat2{a.a + g.a * s, a.b + g.b * s, a.c + g.c * s, a.d + g.d * s}; };
  const Mat2 g1 = fn(x, t);
  const Mat2 g2 = fn(axpy(x, g1, dt / 2), t + dt / 2);
  const Mat2 g3 = fn(axpy(x, g2, dt / 2), t + dt / 2);
  const Mat2 g4 = fn(axpy(x, g3, dt), t + dt);
  return {x.a + dt / 6 * (g1.a + 2 * g2.a + 2 * g3.a + g4.a), x.b + dt / 6 * (g1.b + 2 * g2.b + 2 * g3.b + g4.b),
          x.c + dt / 6 * (g1.c + 2 * g2.c + 2 * g3.c + g4.c), x.d + dt / 6 * (g1.d + 2 * g2.d + 2 * g3.d + g4.d)};
}

LambdaTables::LambdaTables(const CldTables& sde_, double lambda_coef_, bool use_order0_)
    : sde(sde_), lambda_coef(lambda_coef_), use_order0(use_order0_) {
  const double dt = 1e-5;                              // sde_lib.py:358
  const long n = (long)(1.0 / dt);                     // 99999 (floating point), as in the reference
  const long len = n + 1;
  const double tstep = (1.0 + dt) / (double)len;
  xp_.resize(len);
  fp_.resize(len);
  Mat2 x = {1, 0, 0, 1};
  auto fn = [this](const Mat2& v, double t) { return mul(hat_F(t), v); };
  for (long k = 0; k < len; ++k) {
    const double t = k * tstep;
    xp_[k] = t;
    fp_[k] = x;
    x = rk4(x, t, dt, fn);
  }
}

Mat2 LambdaTables::hat_F(double t) const {
  const Mat2 g = sde.G(t), f = sde.F(t), r = sde.R(t);
  const Mat2 m = mul(mul(g, tr(g)), inv(mul(r, tr(r))));
  const double k = 0.5 * (1 + lambda_coef * lambda_coef);
  return {f.a + k * m.a, f.b + k * m.b, f.c + k * m.c, f.d + k * m.d};
}

Mat2 LambdaTables::hat_psi_02t(double t) const {
  long i = std::upper_bound(xp_.begin(), xp_.end(), t) - xp_.begin();
  const long n = (long)xp_.size();
  i = std::min(std::max(i, 1L), n - 1);
  const double dx = xp_[i] - xp_[i - 1];
  if (dx == 0) return fp_[i];
  const double w = (t - xp_[i - 1]) / dx;
  const Mat2 &p = fp_[i - 1], &q = fp_[i];
  return {p.a + w * (q.a - p.a), p.b + w * (q.b - p.b), p.c + w * (q.c - p.c), p.d + w * (q.d - p.d)};
}

Mat2 LambdaTables::hat_psi(double s, double t) const { return mul(hat_psi_02t(t), inv(hat_psi_02t(s))); }

Mat2 LambdaTables::cond_rev_cov(double s, double t) const {
  // literal restatement of sde_lib.py:381-399 (x @ hat_F, not its transpose; grid with endpoint=False)
  const double sign = t > s ? 1.0 : -1.0;
  const int n = 10000;
  const double dt = (t - s) / n;
  const double gstep = (t - s) / (n + 1);
  const double l2 = lambda_coef * lambda_coef;
  auto fn = [this, sign, l2](const Mat2& x, double tt) {
    const Mat2 hf = hat_F(tt), g = sde.G(tt);
    const Mat2 a = mul(hf, x), b = mul(x, hf), gg = mul(g, tr(g));
    return Mat2{a.a + b.a + sign * l2 * gg.a, a.b + b.b + sign * l2 * gg.b, a.c + b.c + sign * l2 * gg.c,
                a.d + b.d + sign * l2 * gg.d};
  };
  Mat2 cov = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i) cov = rk4(cov, s + i * gstep, dt, fn);
  return cov;
}

void LambdaTables::deis_coef(int order, const double* rev_ts, int n_ts, double* out) const {
  const int N = n_ts - 1;
  const int per = order + 4;
  std::memset(out, 0, sizeof(double) * (size_t)N * per * 4);
  auto put = [](double* o, const Mat2& m) { o[0] = m.a; o[1] = m.b; o[2] = m.c; o[3] = m.d; };
  if (use_order0 && order == 0) {
    for (int i = 0; i < N; ++i) {
      const double s = rev_ts[i], t = rev_ts[i + 1];
      const Mat2 xc = sde.psi(s, t);
      const Mat2 hp = hat_psi(s, t);
      const Mat2 ec = mul(Mat2{hp.a - xc.a, hp.b - xc.b, hp.c - xc.c, hp.d - xc.d}, sde.R(s));
      double* o = out + (size_t)i * per * 4;
      put(o, xc); put(o + 4, ec); put(o + 12, cond_rev_cov(s, t));
    }
    return;
  }
  const double k = 0.5 * (1 + lambda_coef * lambda_coef);
  PsiFn psi = [this](double tau, double t_end) { return hat_psi(tau, t_end); };
  IntegrandFn integrand = [this, k](double tau) {
    const Mat2 g = sde.G(tau), r = sde.R(tau);
    const Mat2 m = mul(mul(mul(g, tr(g)), inv(mul(r, tr(r)))), sde.psi(0.0, tau));
    return Mat2{k * m.a, k * m.b, k * m.c, k * m.d};
  };
  std::vector<double> eps;
  deis_ab_eps_coef(psi, integrand, order + 1, rev_ts, n_ts, order, eps);
  for (int i = 0; i < N; ++i) {
    const double s = rev_ts[i], t = rev_ts[i + 1];
    double* o = out + (size_t)i * per * 4;
    put(o, sde.psi(s, t));
    const Mat2 last = mul(sde.psi(s, 0.0), sde.R(s));
    for (int j = 0; j < order + 2; ++j) {
      const double* e = eps.data() + ((size_t)i * (order + 2) + j) * 4;
      put(o + 4 + j * 4, mul(Mat2{e[0], e[1], e[2], e[3]}, last));
    }
    put(o + (size_t)(per - 1) * 4, cond_rev_cov(s, t));
  }
}

// sum_k Psi(tau_k, t_end) integrand(tau_k) l_j(tau_k) dt over tau_k = linspace(t_start, t_end, n, endpoint=False)
Mat2 CldTables::quad(double t_start, double t_end, const double* ts_poly, int n_poly, int coef_idx,
                     int num_item) const {
  const double dt = (t_end - t_start) / num_item;
  Mat2 acc = {0, 0, 0, 0};
  for (int k = 0; k < num_item; ++k) {
    const double tau = t_start + k * dt;
    double w = 1.0;
    if (ts_poly != nullptr) {
      double num = 1.0, den = 1.0;
      for (int q = 0; q < n_poly; ++q)
        if (q != coef_idx) {
          num *= tau - ts_poly[q];
          den *= ts_poly[coef_idx] - ts_poly[q];
        }
      w = num / den;
    }
    const Mat2 m = mul(psi(tau, t_end), eps_integrand(tau));
    acc.a += m.a * w; acc.b += m.b * w; acc.c += m.c * w; acc.d += m.d * w;
  }
  return {acc.a * dt, acc.b * dt, acc.c * dt, acc.d * dt};
}

void CldTables::coef_row(int highest_order, int order, double t_start, double t_end, const double* ts_poly,
                         double* out) const {
  std::memset(out, 0, sizeof(double) * (highest_order + 1) * 4);
  for (int j = 0; j <= order; ++j) {
    const int coef_idx = order - j;            // jnp.flip(arange(order+1)): newest node first
    const Mat2 m = quad(t_start, t_end, ts_poly, order + 1, coef_idx, 10000);
    out[j * 4 + 0] = m.a; out[j * 4 + 1] = m.b; out[j * 4 + 2] = m.c; out[j * 4 + 3] = m.d;
  }
}

void CldTables::ab_eps_coef(int highest_order, const double* ts, int n_ts, int order, std::vector<double>& out) const {
  const int stride = (highest_order + 1) * 4;
  if (order == 0) {
    for (int i = 0; i + 1 < n_ts; ++i) {
      out.resize(out.size() + stride);
      coef_row(highest_order, 0, ts[i], ts[i + 1], ts + i, out.data() + out.size() - stride);
    }
    return;
  }
  ab_eps_coef(highest_order, ts, order + 1, order - 1, out);       // warm-up rows at lower order
  for (int k = 0; k < n_ts - order - 1; ++k) {
    out.resize(out.size() + stride);
    coef_row(highest_order, order, ts[order + k], ts[order + k + 1], ts + k, out.data() + out.size() - stride);
  }
}

void CldTables::deis_coef(int order, const double* rev_ts, int n_ts, double* out) const {
  const int N = n_ts - 1;
  const int highest = order + 1;
  std::vector<double> eps;
  eps.reserve((size_t)N * (highest + 1) * 4);
  ab_eps_coef(highest, rev_ts, n_ts, order, eps);
  const int per = order + 3;
  for (int i = 0; i < N; ++i) {
    const Mat2 p = psi(rev_ts[i], rev_ts[i + 1]);
    double* o = out + (size_t)i * per * 4;
    o[0] = p.a; o[1] = p.b; o[2] = p.c; o[3] = p.d;
    std::memcpy(o + 4, eps.data() + (size_t)i * (highest + 1) * 4, sizeof(double) * (highest + 1) * 4);
  }
}

void CldTables::order0_coef(const double* rev_ts, int n_ts, double* mean_out, double* eps_out) const {
  for (int i = 0; i + 1 < n_ts; ++i) {
    const Mat2 p = psi(rev_ts[i], rev_ts[i + 1]);
    mean_out[i * 4 + 0] = p.a; mean_out[i * 4 + 1] = p.b; mean_out[i * 4 + 2] = p.c; mean_out[i * 4 + 3] = p.d;
    const Mat2 e = quad(rev_ts[i], rev_ts[i + 1], nullptr, 0, 0, 1000);
    eps_out[i * 4 + 0] = e.a; eps_out[i * 4 + 1] = e.b; eps_out[i * 4 + 2] = e.c; eps_out[i * 4 + 3] = e.d;
  }
}

Mat2 CldTables::chol_cov(double t) const {
  const Mat2 r = R(t);
  const Mat2 s = mul(r, tr(r));                       // Sigma_t
  const double l00 = std::sqrt(s.a), l10 = s.c / l00;
  return {l00, 0.0, l10, std::sqrt(s.d - l10 * l10)};
}

void CldTables::ldeis_coef(int order, const double* rev_ts, int n_ts, double* out) const {
  const int N = n_ts - 1, highest = order + 1, per = order + 3;
  PsiFn p = [this](double tau, double t_end) { return psi(tau, t_end); };
  IntegrandFn f = [this](double tau) {
    const Mat2 g = G(tau);
    const Mat2 m = mul(mul(g, g), tr(inv(chol_cov(tau))));
    return Mat2{0.5 * m.a, 0.5 * m.b, 0.5 * m.c, 0.5 * m.d};
  };
  std::vector<double> eps;
  deis_ab_eps_coef(p, f, highest, rev_ts, n_ts, order, eps);
  for (int i = 0; i < N; ++i) {
    const Mat2 x = psi(rev_ts[i], rev_ts[i + 1]);
    double* o = out + (size_t)i * per * 4;
    o[0] = x.a; o[1] = x.b; o[2] = x.c; o[3] = x.d;
    std::memcpy(o + 4, eps.data() + (size_t)i * (highest + 1) * 4, sizeof(double) * (highest + 1) * 4);
  }
}

Mat2 CldTables::f1_psi(double s, double t) const {
  const double bi = beta_int(t) - beta_int(s);
  const double sm = std::sqrt(1.0 / m_inv), ism = std::sqrt(m_inv);
  return {std::cos(bi * ism), ism * std::sin(bi * ism), -sm * std::sin(bi * ism), std::cos(bi * ism)};
}

void CldTables::build_psi2() const {
  if (!p2x_.empty()) return;
  const long N = 100000;                         // sampling.py:273
  const double dt = 1.0 / N;
  p2x_.resize(N + 1);
  p2f_.resize(N + 1);
  Mat2 x = {1, 0, 0, 1};
  double t = 0.0;
  auto fn = [this](const Mat2& v, double tt) {
    const Mat2 f2 = {0.0, 0.0, 0.0, -Gamma * beta(tt) * m_inv};
    return mul(mul(mul(f1_psi(tt, 0.0), f2), f1_psi(0.0, tt)), v);
  };
  for (long k = 0; k <= N; ++k) {                // scan emits (prev_psi2, cur_t) before the update
    p2x_[k] = t;
    p2f_[k] = x;
    x = rk4(x, t, dt, fn);
    t += dt;
  }
}

Mat2 CldTables::psi2(double t) const {
  build_psi2();
  long i = std::upper_bound(p2x_.begin(), p2x_.end(), t) - p2x_.begin();
  const long n = (long)p2x_.size();
  i = std::min(std::max(i, 1L), n - 1);
  const double dx = p2x_[i] - p2x_[i - 1];
  if (dx == 0) return p2f_[i];
  const double w = (t - p2x_[i - 1]) / dx;
  const Mat2 &p = p2f_[i - 1], &q = p2f_[i];
  return {p.a + w * (q.a - p.a), p.b + w * (q.b - p.b), p.c + w * (q.c - p.c), p.d + w * (q.d - p.d)};
}

void CldTables::mldeis_coef(int order, const double* rev_ts, int n_ts, double* out) const {
  const int N = n_ts - 1, highest = order + 1, per = order + 3;
  PsiFn p = [this](double s, double t) { return mul(psi2(t), inv(psi2(s))); };
  IntegrandFn f = [this](double tau) {
    const Mat2 g = G(tau);
    const Mat2 m = mul(mul(mul(f1_psi(tau, 0.0), g), tr(g)), tr(inv(R(tau))));
    return Mat2{0.5 * m.a, 0.5 * m.b, 0.5 * m.c, 0.5 * m.d};
  };
  std::vector<double> eps;
  deis_ab_eps_coef(p, f, highest, rev_ts, n_ts, order, eps);
  for (int i = 0; i < N; ++i) {
    const Mat2 x = p(rev_ts[i], rev_ts[i + 1]);
    double* o = out + (size_t)i * per * 4;
    o[0] = x.a; o[1] = x.b; o[2] = x.c; o[3] = x.d;
    std::memcpy(o + 4, eps.data() + (size_t)i * (highest + 1) * 4, sizeof(double) * (highest + 1) * 4);
  }
}

void CldTables::denoise_coef(double t, Mat2* A, Mat2* C) const {
  // u' = u + (F u)(-t) - (G G score)(-t),  score = -R^{-T} eps   =>   A = I - t F,  C = -t G G R^{-T}
  const Mat2 f = F(t), g = G(t);
  *A = {1.0 - t * f.a, -t * f.b, -t * f.c, 1.0 - t * f.d};
  const Mat2 m = mul(mul(g, g), tr(inv(R(t))));
  *C = {-t * m.a, -t * m.b, -t * m.c, -t * m.d};
}

// ---- blur ---------------------------------------------------------------------------------------------
BlurTables::BlurTables(double sigma_blur_max_, double sampling_eps_, double min_scale_, int img_dim_)
    : sigma_blur_max(sigma_blur_max_), sampling_eps(sampling_eps_), min_scale(min_scale_), img_dim(img_dim_) {
  alpha_start = t2alpha(0.0);
}
double BlurTables::t2alpha(double t) const {
  const double c = std::cos((t + 0.004) / 1.008 * M_PI / 2);
  return c * c;
}
double BlurTables::alpha2t(double a) const { return std::acos(std::sqrt(a)) * 2 / M_PI * 1.008 - 0.004; }
double BlurTables::rho2t(double rho) const {
  const double s = rho + std::sqrt(1 - alpha_start);
  return alpha2t(alpha_start / (s * s + alpha_start));
}
void BlurTables::freq_scaling(double t, double* out) const {
  const double s = std::sin(t * M_PI / 2);
  const double sigma_blur = sigma_blur_max * s * s;
  const double tau = sigma_blur * sigma_blur / 2;
  for (int h = 0; h < img_dim; ++h)
    for (int w = 0; w < img_dim; ++w) {
      const double fh = M_PI * h / img_dim, fw = M_PI * w / img_dim;
      out[h * img_dim + w] = std::exp(-tau * (fh * fh + fw * fw)) * (1 - min_scale) + min_scale;
    }
}
void BlurTables::y_mean_coef(double t, double* out) const {
  freq_scaling(t, out);
  const double sa = std::sqrt(t2alpha(t));
  for (int i = 0; i < img_dim * img_dim; ++i) out[i] *= sa;
}
double BlurTables::y_std_coef(double t) const { return std::sqrt(1 - t2alpha(t)); }

void BlurTables::order0_coef(const double* rev_ts, int n_ts, double* a_out, double* b_out) const {
  const int F = img_dim * img_dim;
  std::vector<double> m(F), mn(F);
  for (int i = 0; i + 1 < n_ts; ++i) {
    const double t = rev_ts[i], tn = rev_ts[i + 1];
    y_mean_coef(t, m.data());
    y_mean_coef(tn, mn.data());
    const double s = y_std_coef(t), sn = y_std_coef(tn);
    for (int f = 0; f < F; ++f) {
      // y0 = (y - s e)/m ; y' = mn y0 + sn e  =>  a = mn/m, b = sn - a s
      const double a = mn[f] / m[f];
      a_out[(size_t)i * F + f] = a;
      b_out[(size_t)i * F + f] = sn - a * s;
    }
  }
}

}  // namespace gddim
