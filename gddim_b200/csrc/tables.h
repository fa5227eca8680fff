// Host-side (fp64) coefficient tables of the CLD gDDIM / DEIS sampler and the blur-diffusion DDIM sampler.
// Built once per sampler; nothing here runs per step.
#pragma once
#include <functional>
#include <vector>

namespace gddim {

struct Mat2 {
  double a, b, c, d;   // row-major [[a, b], [c, d]]
};
inline Mat2 mul(const Mat2& x, const Mat2& y) {
  return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}
inline Mat2 inv(const Mat2& m) {
  const double k = 1.0 / (m.a * m.d - m.b * m.c);
  return {m.d * k, -m.b * k, -m.c * k, m.a * k};
}
inline Mat2 tr(const Mat2& m) { return {m.a, m.c, m.b, m.d}; }

// Generic DEIS Adams-Bashforth table (cld_jax/deis.py:19-95) for any transition psi(tau, t_end) and integrand(tau):
// appends N rows of (highest_order+1) 2x2 matrices to `out`.
using PsiFn = std::function<Mat2(double, double)>;
using IntegrandFn = std::function<Mat2(double)>;
void deis_ab_eps_coef(const PsiFn& psi, const IntegrandFn& integrand, int highest_order, const double* ts, int n_ts,
                      int order, std::vector<double>& out);
// factor F with F F^T = |cov| in the sense of jax.random.multivariate_normal(method='svd'): U * sqrt(S), columns
// sign-normalised so that their largest-magnitude entry is positive
Mat2 mvn_factor_svd(const Mat2& cov);

// linspace(T^(1/p), eps^(1/p), n+1)^p
void rev_timesteps(double T, double eps, int ts_order, int num_step, double* out);

class CldTables {
 public:
  CldTables(double m_inv, double beta_0, double beta_1, double vv_gamma, double numerical_eps, double R_dt,
            bool is_rk);
  double m_inv, beta_0, beta_1, Gamma, R_dt;
  bool is_rk;
  Mat2 R0;
  double sampling_eps = 1e-3, T = 1.0;

  double beta(double t) const { return beta_0 + beta_1 * t; }
  double beta_int(double t) const { return beta_0 * t + 0.5 * beta_1 * t * t; }
  Mat2 F(double t) const;
  Mat2 G(double t) const;
  Mat2 R(double t) const;                    // piecewise-linear table lookup
  Mat2 psi(double s, double t) const;        // closed-form transition matrix
  Mat2 eps_integrand(double t) const;        // 0.5 G G R^{-T}
  // [N, order+3, 2, 2]: [i,0] = Psi(t_i, t_{i+1}); [i,1+j] = DEIS Adams-Bashforth coefficient j; trailing zero
  void deis_coef(int order, const double* rev_ts, int n_ts, double* out) const;
  // [N,2,2] mean and eps matrices of the order-0 gDDIM sampler (1000-node quadrature)
  void order0_coef(const double* rev_ts, int n_ts, double* mean_out, double* eps_out) const;
  // denoising step as u' = A u + C eps
  void denoise_coef(double t, Mat2* A, Mat2* C) const;
  // LSDE (sde_lib.py:469-519): same table with the Cholesky factor L_t of Sigma_t in the integrand
  Mat2 chol_cov(double t) const;
  // MLCLD (cld_jax/sampling.py:272-325, sde_lib.py:120-181): rotating frame psi1 = expm(int F1), psi2 table
  Mat2 f1_psi(double s, double t) const;
  Mat2 psi2(double t) const;
  void mldeis_coef(int order, const double* rev_ts, int n_ts, double* out) const;
  void ldeis_coef(int order, const double* rev_ts, int n_ts, double* out) const;

 private:
  std::vector<double> xp_;
  std::vector<Mat2> fp_;
  mutable std::vector<double> p2x_;      // psi2 table, built on first use
  mutable std::vector<Mat2> p2f_;
  void build_psi2() const;
  Mat2 ode_rhs(const Mat2& R, double t) const;
  Mat2 quad(double t_start, double t_end, const double* ts_poly, int n_poly, int coef_idx, int num_item) const;
  void coef_row(int highest_order, int order, double t_start, double t_end, const double* ts_poly, double* out) const;
  void ab_eps_coef(int highest_order, const double* ts, int n_ts, int order, std::vector<double>& out) const;
};

// Stochastic gDDIM (cld_jax/sde_lib.py:334-466 LambdaSDE): hat-Psi table, conditional covariance, coefficient tables
class LambdaTables {
 public:
  LambdaTables(const CldTables& sde, double lambda_coef, bool use_order0);
  const CldTables& sde;
  double lambda_coef;
  bool use_order0;
  Mat2 hat_F(double t) const;
  Mat2 hat_psi_02t(double t) const;
  Mat2 hat_psi(double s, double t) const;
  Mat2 cond_rev_cov(double s, double t) const;
  // [N, order+4, 2, 2]: x_coef, order+2 eps slots, covariance
  void deis_coef(int order, const double* rev_ts, int n_ts, double* out) const;

 private:
  std::vector<double> xp_;
  std::vector<Mat2> fp_;
};

class BlurTables {
 public:
  BlurTables(double sigma_blur_max, double sampling_eps, double min_scale = 0.001, int img_dim = 32);
  double sigma_blur_max, sampling_eps, min_scale;
  int img_dim;
  double alpha_start;
  double t2alpha(double t) const;
  double alpha2t(double a) const;
  double rho2t(double rho) const;
  double sampling_T() const { return rho2t(80.0); }
  void freq_scaling(double t, double* out /*[dim*dim]*/) const;
  void y_mean_coef(double t, double* out /*[dim*dim]*/) const;
  double y_std_coef(double t) const;
  // per-step per-frequency DDIM coefficients: y' = a .* y + b .* eps_y
  void order0_coef(const double* rev_ts, int n_ts, double* a_out, double* b_out) const;
};

}  // namespace gddim
