// Host-side No-U-Turn sampler over the fused likelihood.
//
// The reference hands its likelihood to NumPyro: `NUTS(numpyro_model)` + `MCMC(...).run(...)`
// (examples/utils.py:63-84; kernel table gwinferno/pipeline/analysis.py:21).  NumPyro/JAX do not exist
// in this environment, and with a 0.1 ms GPU evaluation an interpreted sampler loop costs more than
// the likelihood it drives, so the loop lives here: O(P) host arithmetic per leapfrog step around one
// gwi_loglike_host call, no Python in between.  Two layers:
//   * gwi_nuts_sample    -- the sampler for ANY potential given as a C callback (tests use it on
//                           analytic targets; a JAX/NumPyro-free caller can plug its own model in);
//   * gwi_posterior_*    -- the potential of the reference's B-spline analyses: -(log L + log prior)
//                           with the Gaussian coefficient priors and the P-spline difference penalty
//                           (gwinferno/models/bsplines/smoothing.py:8-28, pipeline/utils.py:163-216;
//                           `fix_first_zero`: pipeline/utils.py:213-214).
// Algorithm: Hoffman & Gelman (2014) Algorithm 6 -- slice-variant NUTS, dual-averaging step size,
// diagonal mass matrix estimated once in the middle of warm-up -- the same as gwinferno_b200/nuts.py
// (the NumPy implementation, kept as the readable specification and cross-check).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "gwi_internal.h"

namespace gwi {
namespace {

// xoshiro256++ seeded through splitmix64; normals by the polar method
struct Rng {
  uint64_t s[4];
  bool have_spare = false;
  double spare = 0.0;
  explicit Rng(uint64_t seed) {
    uint64_t z = seed;
    for (int i = 0; i < 4; ++i) {
      z += 0x9E3779B97F4A7C15ull;
      uint64_t x = z;
      x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
      x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
      s[i] = x ^ (x >> 31);
    }
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0, 1)
  double normal() {
    if (have_spare) {
      have_spare = false;
      return spare;
    }
    double u, v, q;
    do {
      u = 2.0 * uniform() - 1.0;
      v = 2.0 * uniform() - 1.0;
      q = u * u + v * v;
    } while (q >= 1.0 || q == 0.0);
    const double f = std::sqrt(-2.0 * std::log(q) / q);
    spare = v * f;
    have_spare = true;
    return u * f;
  }
};

typedef std::vector<double> Vec;

struct Sampler {
  gwi_potential_fn fn;
  void* ctx;
  int dim;
  Vec inv_mass;  // diagonal of the inverse mass matrix (= variance estimate of the target)
  // dense variant (GWI_NUTS_DENSE_MASS): inverse mass = covariance estimate `cov` (row-major), its lower
  // Cholesky factor `chol` for the momentum draws; used once the first slow window has been estimated
  bool dense = false, want_dense = false;
  Vec cov, chol;
  mutable Vec tmp_v;
  Rng rng;
  int64_t n_evals = 0, n_leapfrog = 0;

  Sampler(gwi_potential_fn f, void* c, int d, uint64_t seed) : fn(f), ctx(c), dim(d), inv_mass(d, 1.0), rng(seed) {}

  double U(const Vec& theta, Vec& grad) {
    ++n_evals;
    double u = fn(ctx, theta.data(), grad.data());
    if (!(u == u)) u = std::numeric_limits<double>::infinity();
    return u;
  }
  // v = M^-1 r
  void velocity(const Vec& r, Vec& v) const {
    v.resize(dim);
    if (!dense) {
      for (int i = 0; i < dim; ++i) v[i] = inv_mass[i] * r[i];
      return;
    }
    // four independent partial sums per row: one running sum is a serial chain of `dim` dependent additions (the compiler
    // may not re-associate them), ~4 cycles each -- at dim = 164 that was the sampler's largest host cost per leapfrog step
    const double* rp = r.data();
    for (int i = 0; i < dim; ++i) {
      const double* row = cov.data() + (size_t)i * dim;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int j = 0;
      for (; j + 4 <= dim; j += 4) {
        a0 += row[j] * rp[j];
        a1 += row[j + 1] * rp[j + 1];
        a2 += row[j + 2] * rp[j + 2];
        a3 += row[j + 3] * rp[j + 3];
      }
      for (; j < dim; ++j) a0 += row[j] * rp[j];
      v[i] = (a0 + a1) + (a2 + a3);
    }
  }
  double kinetic(const Vec& r) const {
    double k = 0.0;
    if (!dense) {
      for (int i = 0; i < dim; ++i) k += inv_mass[i] * r[i] * r[i];
      return 0.5 * k;
    }
    velocity(r, tmp_v);
    for (int i = 0; i < dim; ++i) k += r[i] * tmp_v[i];
    return 0.5 * k;
  }
  // r ~ N(0, M), M = cov^-1 = (L L^T)^-1  =>  r = L^-T z
  void draw_momentum(Vec& r) {
    if (!dense) {
      for (int i = 0; i < dim; ++i) r[i] = rng.normal() / std::sqrt(inv_mass[i]);
      return;
    }
    for (int i = 0; i < dim; ++i) r[i] = rng.normal();
    for (int i = dim - 1; i >= 0; --i) {
      double a = r[i];
      for (int j = i + 1; j < dim; ++j) a -= chol[(size_t)j * dim + i] * r[j];
      r[i] = a / chol[(size_t)i * dim + i];
    }
  }
  // one leapfrog step from (theta, r, grad) in place; returns the new potential
  double leapfrog(Vec& theta, Vec& r, Vec& grad, double eps) {
    for (int i = 0; i < dim; ++i) r[i] -= 0.5 * eps * grad[i];
    if (!dense) {
      for (int i = 0; i < dim; ++i) theta[i] += eps * inv_mass[i] * r[i];
    } else {
      velocity(r, tmp_v);
      for (int i = 0; i < dim; ++i) theta[i] += eps * tmp_v[i];
    }
    const double u = U(theta, grad);
    for (int i = 0; i < dim; ++i) r[i] -= 0.5 * eps * grad[i];
    return u;
  }
  static double finite_or_inf(double h) { return std::isfinite(h) ? h : std::numeric_limits<double>::infinity(); }

  double find_reasonable_eps(const Vec& theta, double u0, const Vec& g0) {
    double eps = 0.05;
    Vec r(dim);
    draw_momentum(r);
    const double h0 = u0 + kinetic(r);
    Vec th1 = theta, r1 = r, g1 = g0;
    double u1 = leapfrog(th1, r1, g1, eps);
    double h1 = u1 + kinetic(r1);
    const double a = (std::isfinite(h1) && h0 - h1 > std::log(0.5)) ? 1.0 : -1.0;
    for (int it = 0; it < 30; ++it) {
      th1 = theta;
      r1 = r;
      g1 = g0;
      u1 = leapfrog(th1, r1, g1, eps);
      h1 = finite_or_inf(u1 + kinetic(r1));
      if (a * (h0 - h1) <= -a * std::log(2.0)) break;
      eps *= std::pow(2.0, a);
    }
    return eps;
  }

  struct Tree {
    Vec thm, rm, gm, thp, rp, gp, th1, g1;
    double u1 = 0.0;
    int64_t n1 = 0;
    int s1 = 0;
    double alpha = 0.0;
    int64_t n_alpha = 0;
  };

  bool no_uturn(const Vec& thm, const Vec& rm, const Vec& thp, const Vec& rp) const {
    double a = 0.0, b = 0.0;
    Vec vm, vp;
    velocity(rm, vm);
    velocity(rp, vp);
    for (int i = 0; i < dim; ++i) {
      const double d = thp[i] - thm[i];
      a += d * vm[i];
      b += d * vp[i];
    }
    return a >= 0.0 && b >= 0.0;
  }

  // Algorithm 6, BuildTree(theta, r, u, v, j, eps, theta0, r0) with the joint log-density -H
  void build_tree(const Vec& theta, const Vec& r, const Vec& grad, double logu, int v, int j, double eps, double h0, Tree& T) {
    if (j == 0) {
      T.th1 = theta;
      T.thm = r;  // scratch: momentum
      T.g1 = grad;
      T.u1 = leapfrog(T.th1, T.thm, T.g1, v * eps);
      ++n_leapfrog;
      const double h1 = finite_or_inf(T.u1 + kinetic(T.thm));
      T.rm = T.thm;
      T.rp = T.thm;
      T.thm = T.th1;
      T.thp = T.th1;
      T.gm = T.g1;
      T.gp = T.g1;
      T.n1 = logu <= -h1 ? 1 : 0;
      T.s1 = logu < 1000.0 - h1 ? 1 : 0;
      T.alpha = std::isfinite(h1) ? std::fmin(1.0, std::exp(std::fmin(0.0, h0 - h1))) : 0.0;
      T.n_alpha = 1;
      return;
    }
    build_tree(theta, r, grad, logu, v, j - 1, eps, h0, T);
    if (T.s1 == 1) {
      Tree T2;
      if (v == -1) {
        build_tree(T.thm, T.rm, T.gm, logu, v, j - 1, eps, h0, T2);
        T.thm.swap(T2.thm);
        T.rm.swap(T2.rm);
        T.gm.swap(T2.gm);
      } else {
        build_tree(T.thp, T.rp, T.gp, logu, v, j - 1, eps, h0, T2);
        T.thp.swap(T2.thp);
        T.rp.swap(T2.rp);
        T.gp.swap(T2.gp);
      }
      if (T.n1 + T2.n1 > 0 && rng.uniform() < (double)T2.n1 / (double)(T.n1 + T2.n1)) {
        T.th1.swap(T2.th1);
        T.g1.swap(T2.g1);
        T.u1 = T2.u1;
      }
      T.alpha += T2.alpha;
      T.n_alpha += T2.n_alpha;
      T.s1 = T2.s1 * (no_uturn(T.thm, T.rm, T.thp, T.rp) ? 1 : 0);
      T.n1 += T2.n1;
    }
  }

  // ---- multinomial NUTS (Betancourt 2017; the scheme of Stan's base_nuts and of NumPyro's default
  //      kernel): states of a trajectory are drawn with weights exp(-H), biased towards the newer
  //      subtree at the top level; the U-turn criterion uses the summed momentum rho and is also
  //      demanded across the junction of the two half-trees.
  struct Point {
    Vec theta, r, grad;
    double u = 0.0;
  };
  static double log_add_exp(double a, double b) {
    if (a == -std::numeric_limits<double>::infinity()) return b;
    if (b == -std::numeric_limits<double>::infinity()) return a;
    const double m = a > b ? a : b;
    return m + std::log(std::exp(a - m) + std::exp(b - m));
  }
  void sharp(const Vec& r, Vec& out) const { velocity(r, out); }
  bool criterion(const Vec& sharp_minus, const Vec& sharp_plus, const Vec& rho) const {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < dim; ++i) {
      a += sharp_plus[i] * rho[i];
      b += sharp_minus[i] * rho[i];
    }
    return a > 0.0 && b > 0.0;
  }
  struct MultiState {
    Point z;  // the integrator's moving end
    double H0 = 0.0, eps = 0.0;
    int64_t n_leapfrog = 0;
    double sum_metro = 0.0;
    bool divergent = false;
  };
  bool build_multi(MultiState& S, int depth, Point& z_propose, Vec& sharp_beg, Vec& sharp_end, Vec& rho, Vec& p_beg, Vec& p_end, int sign,
                   double& log_sum_weight) {
    if (depth == 0) {
      S.z.u = leapfrog(S.z.theta, S.z.r, S.z.grad, sign * S.eps);
      ++S.n_leapfrog;
      ++n_leapfrog;
      // M^-1 r once: it is both the U-turn criterion's "sharp" vector and, dotted with r, twice the kinetic energy
      sharp(S.z.r, sharp_beg);
      double kin = 0.0;
      if (dense) {
        for (int i = 0; i < dim; ++i) kin += S.z.r[i] * sharp_beg[i];
        kin *= 0.5;
      } else {
        kin = kinetic(S.z.r);
      }
      const double h = finite_or_inf(S.z.u + kin);
      if (h - S.H0 > 1000.0) S.divergent = true;
      log_sum_weight = log_add_exp(log_sum_weight, S.H0 - h);
      S.sum_metro += S.H0 - h > 0.0 ? 1.0 : std::exp(S.H0 - h);
      z_propose = S.z;
      sharp_end = sharp_beg;
      for (int i = 0; i < dim; ++i) rho[i] += S.z.r[i];
      p_beg = S.z.r;
      p_end = S.z.r;
      return !S.divergent;
    }
    const double ninf = -std::numeric_limits<double>::infinity();
    double lsw_init = ninf;
    Vec p_init_end(dim), sharp_init_end(dim), rho_init(dim, 0.0);
    if (!build_multi(S, depth - 1, z_propose, sharp_beg, sharp_init_end, rho_init, p_beg, p_init_end, sign, lsw_init)) return false;
    Point z_propose_final = S.z;
    double lsw_final = ninf;
    Vec p_final_beg(dim), sharp_final_beg(dim), rho_final(dim, 0.0);
    if (!build_multi(S, depth - 1, z_propose_final, sharp_final_beg, sharp_end, rho_final, p_final_beg, p_end, sign, lsw_final)) return false;
    const double lsw_subtree = log_add_exp(lsw_init, lsw_final);
    log_sum_weight = log_add_exp(log_sum_weight, lsw_subtree);
    if (lsw_final > lsw_subtree) {
      z_propose = z_propose_final;
    } else if (rng.uniform() < std::exp(lsw_final - lsw_subtree)) {
      z_propose = z_propose_final;
    }
    Vec rho_subtree(dim), ext(dim);
    for (int i = 0; i < dim; ++i) {
      rho_subtree[i] = rho_init[i] + rho_final[i];
      rho[i] += rho_subtree[i];
    }
    bool persist = criterion(sharp_beg, sharp_end, rho_subtree);
    for (int i = 0; i < dim; ++i) ext[i] = rho_init[i] + p_final_beg[i];
    persist = persist && criterion(sharp_beg, sharp_final_beg, ext);
    for (int i = 0; i < dim; ++i) ext[i] = rho_final[i] + p_init_end[i];
    persist = persist && criterion(sharp_init_end, sharp_end, ext);
    return persist;
  }
  // one transition from (theta, grad, u); returns the mean acceptance statistic of the trajectory
  double transition_multi(Vec& theta, Vec& grad, double& u, double eps, int max_depth) {
    MultiState S;
    S.eps = eps;
    S.z.theta = theta;
    S.z.grad = grad;
    S.z.u = u;
    S.z.r.resize(dim);
    draw_momentum(S.z.r);
    S.H0 = u + kinetic(S.z.r);
    Point z_fwd = S.z, z_bck = S.z, z_sample = S.z, z_propose = S.z;
    Vec p_fwd_fwd = S.z.r, p_fwd_bck = S.z.r, p_bck_fwd = S.z.r, p_bck_bck = S.z.r;
    Vec sharp_fwd_fwd;
    sharp(S.z.r, sharp_fwd_fwd);
    Vec sharp_fwd_bck = sharp_fwd_fwd, sharp_bck_fwd = sharp_fwd_fwd, sharp_bck_bck = sharp_fwd_fwd;
    Vec rho = S.z.r, rho_fwd(dim), rho_bck(dim), ext(dim);
    double log_sum_weight = 0.0;
    int depth = 0;
    while (depth < max_depth) {
      std::fill(rho_fwd.begin(), rho_fwd.end(), 0.0);
      std::fill(rho_bck.begin(), rho_bck.end(), 0.0);
      bool valid;
      double lsw_subtree = -std::numeric_limits<double>::infinity();
      if (rng.uniform() > 0.5) {
        S.z = z_fwd;
        rho_bck = rho;
        p_bck_fwd = p_fwd_fwd;
        sharp_bck_fwd = sharp_fwd_fwd;
        valid = build_multi(S, depth, z_propose, sharp_fwd_bck, sharp_fwd_fwd, rho_fwd, p_fwd_bck, p_fwd_fwd, +1, lsw_subtree);
        z_fwd = S.z;
      } else {
        S.z = z_bck;
        rho_fwd = rho;
        p_fwd_bck = p_bck_bck;
        sharp_fwd_bck = sharp_bck_bck;
        valid = build_multi(S, depth, z_propose, sharp_bck_fwd, sharp_bck_bck, rho_bck, p_bck_fwd, p_bck_bck, -1, lsw_subtree);
        z_bck = S.z;
      }
      if (!valid) break;
      ++depth;
      if (lsw_subtree > log_sum_weight) {
        z_sample = z_propose;
      } else if (rng.uniform() < std::exp(lsw_subtree - log_sum_weight)) {
        z_sample = z_propose;
      }
      log_sum_weight = log_add_exp(log_sum_weight, lsw_subtree);
      for (int i = 0; i < dim; ++i) rho[i] = rho_bck[i] + rho_fwd[i];
      bool persist = criterion(sharp_bck_bck, sharp_fwd_fwd, rho);
      for (int i = 0; i < dim; ++i) ext[i] = rho_bck[i] + p_fwd_bck[i];
      persist = persist && criterion(sharp_bck_bck, sharp_fwd_bck, ext);
      for (int i = 0; i < dim; ++i) ext[i] = rho_fwd[i] + p_bck_fwd[i];
      persist = persist && criterion(sharp_bck_fwd, sharp_fwd_fwd, ext);
      if (!persist) break;
    }
    theta = z_sample.theta;
    grad = z_sample.grad;
    u = z_sample.u;
    return S.n_leapfrog > 0 ? S.sum_metro / (double)S.n_leapfrog : 0.0;
  }

  // Dense inverse mass from warm[first:]: sample covariance shrunk towards its diagonal with the
  // analytic intensity of Schaefer & Strimmer (2005) (off-diagonal entries that the window cannot
  // resolve are damped; with few draws this degrades gracefully to the diagonal estimate), then Stan's
  // regularisation towards 1e-3 I.  Falls back to the diagonal estimate if the factorisation fails.
  bool estimate_dense(const std::vector<Vec>& warm, size_t first) {
    const size_t nn = warm.size() - first;
    if (nn < 10) return false;
    const size_t D = (size_t)dim;
    Vec mean(D, 0.0);
    for (size_t k = first; k < warm.size(); ++k)
      for (size_t i = 0; i < D; ++i) mean[i] += warm[k][i];
    for (size_t i = 0; i < D; ++i) mean[i] /= (double)nn;
    Vec S(D * D, 0.0), W2(D * D, 0.0), x(D);
    for (size_t k = first; k < warm.size(); ++k) {
      for (size_t i = 0; i < D; ++i) x[i] = warm[k][i] - mean[i];
      for (size_t i = 0; i < D; ++i)
        for (size_t j = 0; j <= i; ++j) {
          const double w = x[i] * x[j];
          S[i * D + j] += w;
          W2[i * D + j] += w * w;
        }
    }
    const double n = (double)nn;
    double num = 0.0, den = 0.0;
    for (size_t i = 0; i < D; ++i)
      for (size_t j = 0; j < i; ++j) {
        const double wbar = S[i * D + j] / n;
        const double var_w = (W2[i * D + j] / n - wbar * wbar) * n / (n - 1.0);  // variance of w_k
        const double sij = S[i * D + j] / (n - 1.0);
        num += var_w * n / ((n - 1.0) * (n - 1.0));  // Var(s_ij)
        den += sij * sij;
      }
    double lambda = den > 0.0 ? num / den : 1.0;
    lambda = std::fmin(1.0, std::fmax(0.0, lambda));
    cov.assign(D * D, 0.0);
    const double a = n / (n + 5.0), b = 1e-3 * (5.0 / (n + 5.0));
    for (size_t i = 0; i < D; ++i)
      for (size_t j = 0; j <= i; ++j) {
        double c = S[i * D + j] / (n - 1.0);
        if (i != j) c *= 1.0 - lambda;
        c *= a;
        if (i == j) c += b;
        cov[i * D + j] = c;
        cov[j * D + i] = c;
      }
    // Cholesky cov = L L^T
    chol.assign(D * D, 0.0);
    for (size_t i = 0; i < D; ++i)
      for (size_t j = 0; j <= i; ++j) {
        double acc = cov[i * D + j];
        for (size_t k = 0; k < j; ++k) acc -= chol[i * D + k] * chol[j * D + k];
        if (i == j) {
          if (!(acc > 0.0)) return false;
          chol[i * D + i] = std::sqrt(acc);
        } else {
          chol[i * D + j] = acc / chol[j * D + j];
        }
      }
    for (size_t i = 0; i < D; ++i) inv_mass[i] = cov[i * D + i];
    shrinkage = lambda;
    return true;
  }
  double shrinkage = 1.0;

  // regularised diagonal variance of warm[first:] (Stan's shrinkage towards 1e-3)
  void estimate_mass(const std::vector<Vec>& warm, size_t first) {
    if (want_dense) {
      dense = estimate_dense(warm, first);
      if (dense) return;
    }
    const size_t nn = warm.size() - first;
    for (int i = 0; i < dim; ++i) {
      double mean = 0.0;
      for (size_t k = first; k < warm.size(); ++k) mean += warm[k][i];
      mean /= (double)nn;
      double var = 0.0;
      for (size_t k = first; k < warm.size(); ++k) var += (warm[k][i] - mean) * (warm[k][i] - mean);
      var /= (double)nn;
      inv_mass[i] = ((double)nn / ((double)nn + 5.0)) * var + 1e-3 * (5.0 / ((double)nn + 5.0));
    }
  }

  int run(const double* theta0, const gwi_nuts_opts& o, double* samples, gwi_nuts_info* info) {
    Vec theta(theta0, theta0 + dim), grad(dim);
    double u = U(theta, grad);
    if (!std::isfinite(u)) {
      set_error("NUTS: the potential is not finite at the starting point");
      return GWI_ERR_INVALID;
    }
    double eps = find_reasonable_eps(theta, u, grad);
    double mu = std::log(10.0 * eps);
    const double gamma = 0.05, t0 = 10.0, kappa = 0.75;
    double eps_bar = o.n_warmup > 0 ? 1.0 : eps, Hbar = 0.0;  // no warm-up: sample with the step size the search found
    std::vector<Vec> warm;
    const bool multinomial = (o.flags & GWI_NUTS_MULTINOMIAL) != 0, windowed = (o.flags & GWI_NUTS_WINDOWED_ADAPT) != 0;
    want_dense = (o.flags & GWI_NUTS_DENSE_MASS) != 0;
    // windowed adaptation (Stan's schedule): a fast initial buffer (15 % of warm-up), slow windows that
    // double in length and each end with a mass-matrix update + step-size restart, a final fast buffer (10 %)
    int slow_start = -1, slow_end = -1, win_end = -1, win_len = 0;
    if (windowed && o.n_warmup >= 20) {
      slow_start = std::max(1, (int)(0.15 * o.n_warmup));
      slow_end = o.n_warmup - std::max(1, (int)(0.10 * o.n_warmup));
      win_len = std::max(5, (slow_end - slow_start) / 7);  // window lengths ~ 1 : 2 : 4
      win_end = slow_start + win_len;
      if (win_end + 2 * win_len > slow_end) win_end = slow_end;  // no room for a second window
      warm.reserve(o.n_warmup);
    }
    size_t win_first = 0;  // first warm-up draw of the current slow window
    int adapt_origin = 0;  // windowed adaptation restarts the dual-averaging iteration count with every window
    double accept_sum = 0.0;
    int64_t leapfrog_sampling0 = 0;
    auto t_sampling = std::chrono::steady_clock::now();
    Vec r0(dim), thm, thp, rm, rp, gm, gp;
    const int n_total = o.n_warmup + o.n_samples;
    for (int m = 0; m < n_total; ++m) {
      if (m == o.n_warmup) {
        t_sampling = std::chrono::steady_clock::now();
        leapfrog_sampling0 = n_leapfrog;
      }
      draw_momentum(r0);
      const double h0 = u + kinetic(r0);
      const double logu = std::log(rng.uniform()) - h0;
      thm = thp = theta;
      rm = rp = r0;
      gm = gp = grad;
      int j = 0, s = 1;
      int64_t n = 1;
      const double step = m < o.n_warmup ? eps : eps_bar;
      double a = 0.0;
      int64_t na = 1;
      if (multinomial) {
        a = transition_multi(theta, grad, u, step, o.max_depth);
        s = 0;
      }
      while (s == 1 && j < o.max_depth) {
        const int v = rng.uniform() < 0.5 ? -1 : 1;
        Tree T;
        if (v == -1) {
          build_tree(thm, rm, gm, logu, v, j, step, h0, T);
          thm.swap(T.thm);
          rm.swap(T.rm);
          gm.swap(T.gm);
        } else {
          build_tree(thp, rp, gp, logu, v, j, step, h0, T);
          thp.swap(T.thp);
          rp.swap(T.rp);
          gp.swap(T.gp);
        }
        if (T.s1 == 1 && rng.uniform() < std::fmin(1.0, (double)T.n1 / (double)n)) {
          theta.swap(T.th1);
          grad.swap(T.g1);
          u = T.u1;
        }
        n += T.n1;
        s = T.s1 * (no_uturn(thm, rm, thp, rp) ? 1 : 0);
        a = T.alpha;
        na = T.n_alpha;
        ++j;
      }
      const double acc = a / (double)(na > 0 ? na : 1);
      if (m < o.n_warmup) {
        const double mm = (double)(m + 1 - adapt_origin);
        Hbar = (1.0 - 1.0 / (mm + t0)) * Hbar + (o.target_accept - acc) / (mm + t0);
        eps = std::exp(mu - std::sqrt(mm) / gamma * Hbar);
        const double eta = std::pow(mm, -kappa);
        eps_bar = std::exp(eta * std::log(eps) + (1.0 - eta) * std::log(eps_bar));
        warm.push_back(theta);
        bool update = false;
        size_t first = 0;
        if (windowed && slow_start > 0) {
          if (m + 1 == slow_start) win_first = warm.size();
          if (m + 1 == win_end) {
            update = warm.size() - win_first >= 5;
            first = win_first;
            win_first = warm.size();
            if (win_end >= slow_end) {
              win_end = -1;  // the final fast buffer only tunes the step size
            } else {
              win_len *= 2;
              win_end += win_len;
              if (win_end + 2 * win_len > slow_end) win_end = slow_end;  // the last window absorbs the remainder
            }
          }
        } else if (m + 1 == o.n_warmup / 2 && warm.size() >= 20) {
          // one mass-matrix update in the middle of warm-up (diagonal, regularised sample variance)
          update = true;
          first = warm.size() / 4;
        }
        if (update) {
          estimate_mass(warm, first);
          u = U(theta, grad);
          eps = find_reasonable_eps(theta, u, grad);
          mu = std::log(10.0 * eps);
          eps_bar = 1.0;
          Hbar = 0.0;
          adapt_origin = m + 1;  // the dual averaging restarts with the new metric (otherwise eta = mm^-0.75 is already ~0.01 and
                                 // the reset eps_bar = 1 is never averaged away: step size biased towards 1)
        }
      } else {
        std::memcpy(samples + (size_t)(m - o.n_warmup) * dim, theta.data(), sizeof(double) * dim);
        accept_sum += acc;
      }
    }
    if (info) {
      info->step_size = eps_bar;
      info->mean_accept = o.n_samples > 0 ? accept_sum / o.n_samples : 0.0;
      info->sampling_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_sampling).count();
      info->leapfrogs_sampling = n_leapfrog - leapfrog_sampling0;
      info->leapfrogs_total = n_leapfrog;
      info->n_evals = n_evals;
    }
    return GWI_OK;
  }
};

}  // namespace
}  // namespace gwi

using namespace gwi;

// -(log L + log prior) of one gwi_model: Lambda = theta scattered into the free slots
struct gwi_posterior {
  gwi_model* model = nullptr;
  gwi_like_opts opts{};
  int n_params = 0, dim = 0;
  std::vector<int> free_slot;  // free_slot[i] = Lambda index of theta[i]
  struct Block {
    int first, count;
    std::vector<double> Q;  // count x count: I / sigma^2 + tau D^T D
  };
  std::vector<Block> blocks;
  std::vector<double> lam, out, gp;
  int last_rc = GWI_OK;
  int64_t n_evals = 0;
};

extern "C" {

int gwi_nuts_sample(gwi_potential_fn fn, void* ctx, int32_t dim, const double* theta0, const gwi_nuts_opts* opts, double* samples, gwi_nuts_info* info) {
  if (!fn || dim < 1 || !theta0 || !opts || (!samples && opts->n_samples > 0) || opts->n_warmup < 0 || opts->n_samples < 0 || opts->max_depth < 1 ||
      opts->max_depth > 20 || !(opts->target_accept > 0.0 && opts->target_accept < 1.0)) {
    set_error("gwi_nuts_sample: bad argument");
    return GWI_ERR_INVALID;
  }
  try {
    Sampler S(fn, ctx, dim, (uint64_t)opts->seed);
    return S.run(theta0, *opts, samples, info);
  } catch (const std::bad_alloc&) {
    set_error("out of host memory in the sampler");
    return GWI_ERR_ALLOC;
  }
}

int gwi_posterior_create(gwi_model* m, const gwi_like_opts* opts, const gwi_prior_block* blocks, int32_t n_blocks, gwi_posterior** out) {
  if (!m || !opts || !out || n_blocks < 0 || (n_blocks > 0 && !blocks)) {
    set_error("gwi_posterior_create: null argument");
    return GWI_ERR_INVALID;
  }
  const int64_t psize = gwi_partial_size(m);
  if (psize < 0) return GWI_ERR_INVALID;
  const int P = (int)((psize - 8) / 3);  // PR_HEADER + 3P
  gwi_posterior* p = new (std::nothrow) gwi_posterior();
  if (!p) return GWI_ERR_ALLOC;
  try {
    p->model = m;
    p->opts = *opts;
    p->n_params = P;
    std::vector<char> is_free(P, 1);
    for (int b = 0; b < n_blocks; ++b) {
      const gwi_prior_block& B = blocks[b];
      if (B.first < 0 || B.count < 1 || B.first + B.count > P || !(B.sigma > 0.0) || B.diff_degree < 0) {
        delete p;
        set_error("gwi_posterior_create: prior block outside the parameter vector, or sigma <= 0");
        return GWI_ERR_INVALID;
      }
      if (B.fix_first_zero) is_free[B.first] = 0;
      gwi_posterior::Block X;
      X.first = B.first;
      X.count = B.count;
      const int n = B.count;
      X.Q.assign((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i) X.Q[(size_t)i * n + i] = 1.0 / (B.sigma * B.sigma);
      if (B.tau >= 0.0 && n > B.diff_degree) {
        // D = diff(I, n = degree): rows of binomial coefficients with alternating sign
        const int deg = B.diff_degree, rows = n - deg;
        std::vector<double> coef(deg + 1, 0.0);
        coef[0] = 1.0;
        for (int k = 0; k < deg; ++k) {
          for (int i = k + 1; i >= 1; --i) coef[i] = coef[i - 1] - coef[i];
          coef[0] = -coef[0];
        }
        for (int r = 0; r < rows; ++r)
          for (int a = 0; a <= deg; ++a)
            for (int c = 0; c <= deg; ++c) X.Q[(size_t)(r + a) * n + (r + c)] += B.tau * coef[a] * coef[c];
      }
      p->blocks.push_back(std::move(X));
    }
    for (int i = 0; i < P; ++i)
      if (is_free[i]) p->free_slot.push_back(i);
    p->dim = (int)p->free_slot.size();
    p->lam.assign(P, 0.0);
    p->gp.assign(P, 0.0);
    p->out.assign((size_t)GWI_LIKE_HEADER + P, 0.0);
  } catch (const std::bad_alloc&) {
    delete p;
    return GWI_ERR_ALLOC;
  }
  *out = p;
  return GWI_OK;
}

void gwi_posterior_destroy(gwi_posterior* p) { delete p; }

int gwi_posterior_dim(const gwi_posterior* p) { return p ? p->dim : (int)GWI_ERR_INVALID; }

// theta -> Lambda: the free slots, everything else 0
static void posterior_fill_lambda(const gwi_posterior* p, const double* theta, double* lam) {
  std::fill(lam, lam + p->n_params, 0.0);
  for (int i = 0; i < p->dim; ++i) lam[p->free_slot[i]] = theta[i];
}

// (Lambda, likelihood row) -> potential and its gradient in theta; `gp`: n_params doubles of scratch
static double posterior_finish(const gwi_posterior* p, const double* lam, const double* out, double* gp, double* grad) {
  const double inf = std::numeric_limits<double>::infinity();
  for (int i = 0; i < p->dim; ++i) grad[i] = 0.0;
  const double log_l = out[GWI_LIKE_LOG_L];
  if (!std::isfinite(log_l) || log_l < -1e300) return inf;
  double lp = 0.0;
  std::fill(gp, gp + p->n_params, 0.0);
  for (const auto& B : p->blocks) {
    const double* c = lam + B.first;
    const int n = B.count;
    for (int i = 0; i < n; ++i) {
      double qc = 0.0;
      const double* row = B.Q.data() + (size_t)i * n;
      for (int j = 0; j < n; ++j) qc += row[j] * c[j];
      lp -= 0.5 * c[i] * qc;
      gp[B.first + i] -= qc;
    }
  }
  for (int i = 0; i < p->dim; ++i) {
    const int s = p->free_slot[i];
    grad[i] = -(out[GWI_LIKE_HEADER + s] + gp[s]);
  }
  return -(log_l + lp);
}

// matches gwi_potential_fn (ctx = the gwi_posterior); +inf where the likelihood's cuts fail (the
// reference's nan_to_num(-inf) sentinel, analysis.py:272-277) or the evaluation reports an error
double gwi_posterior_potential(void* ctx, const double* theta, double* grad) {
  gwi_posterior* p = static_cast<gwi_posterior*>(ctx);
  posterior_fill_lambda(p, theta, p->lam.data());
  ++p->n_evals;
  const int rc = gwi_loglike_host(p->model, p->lam.data(), &p->opts, p->out.data());
  if (rc != GWI_OK) {
    for (int i = 0; i < p->dim; ++i) grad[i] = 0.0;
    if (rc != GWI_ERR_RANGE) p->last_rc = rc;
    return std::numeric_limits<double>::infinity();
  }
  return posterior_finish(p, p->lam.data(), p->out.data(), p->gp.data(), grad);
}

int gwi_nuts_sample_posterior(gwi_posterior* p, const double* theta0, const gwi_nuts_opts* opts, double* samples, gwi_nuts_info* info) {
  if (!p) {
    set_error("null argument");
    return GWI_ERR_INVALID;
  }
  p->last_rc = GWI_OK;
  const int64_t before = p->n_evals;
  const int rc = gwi_nuts_sample(gwi_posterior_potential, p, p->dim, theta0, opts, samples, info);
  if (info) info->n_evals = p->n_evals - before;
  if (rc != GWI_OK) return rc;
  return p->last_rc;  // a CUDA / argument error inside an evaluation surfaces here
}

// ---- several chains advanced together (one batched likelihood call per round of leapfrog steps) ----
namespace {
enum : int { CH_RUNNING = 0, CH_PENDING = 1, CH_READY = 2, CH_ENDED = 3 };
struct ChainShared {
  gwi_posterior* p = nullptr;
  int K = 0;
  size_t row_out = 0;
  std::vector<double> lam, out;     // [K][P], [K][row_out]: one row per chain
  std::vector<std::atomic<int>> state;
  std::vector<int> rc_eval;         // result code of the chain's last evaluation (written before CH_READY)
  std::vector<uint64_t> ticket;     // arrival order of the pending requests (written before CH_PENDING)
  std::atomic<uint64_t> next_ticket{0};
  std::atomic<int> fatal{GWI_OK};   // a CUDA / argument error: no more evaluations, every potential is +inf from then on
  explicit ChainShared(int k) : K(k), state(k), rc_eval(k, GWI_OK), ticket(k, 0) {
    for (auto& s : state) s.store(CH_RUNNING);
  }
};
struct ChainCtx {
  ChainShared* sh;
  int c;
  std::vector<double> gp;
  int64_t n_evals = 0;
};
double chain_potential(void* ctx, const double* theta, double* grad) {
  ChainCtx* x = static_cast<ChainCtx*>(ctx);
  ChainShared& S = *x->sh;
  const gwi_posterior* p = S.p;
  const double inf = std::numeric_limits<double>::infinity();
  double* lam = S.lam.data() + (size_t)x->c * p->n_params;
  posterior_fill_lambda(p, theta, lam);
  ++x->n_evals;
  if (S.fatal.load(std::memory_order_acquire) != GWI_OK) {
    for (int i = 0; i < p->dim; ++i) grad[i] = 0.0;
    return inf;
  }
  S.ticket[x->c] = S.next_ticket.fetch_add(1, std::memory_order_relaxed);
  S.state[x->c].store(CH_PENDING, std::memory_order_release);
  for (unsigned spin = 0; S.state[x->c].load(std::memory_order_acquire) != CH_READY; ++spin)
    if (spin > 64) std::this_thread::yield();
  S.state[x->c].store(CH_RUNNING, std::memory_order_relaxed);
  if (S.rc_eval[x->c] != GWI_OK) {
    for (int i = 0; i < p->dim; ++i) grad[i] = 0.0;
    return inf;
  }
  return posterior_finish(p, lam, S.out.data() + (size_t)x->c * S.row_out, x->gp.data(), grad);
}
}  // namespace

int gwi_nuts_sample_posterior_chains(gwi_posterior* p, int32_t n_chains, const double* theta0, const gwi_nuts_opts* opts, double* samples, gwi_nuts_info* info) {
  if (!p || !theta0 || !opts || n_chains < 1 || n_chains > 4096 || (!samples && opts->n_samples > 0) || opts->n_warmup < 0 || opts->n_samples < 0 ||
      opts->max_depth < 1 || opts->max_depth > 20 || !(opts->target_accept > 0.0 && opts->target_accept < 1.0)) {
    set_error("gwi_nuts_sample_posterior_chains: bad argument");
    return GWI_ERR_INVALID;
  }
  try {
    const int K = n_chains, P = p->n_params, dim = p->dim;
    ChainShared S(K);
    S.p = p;
    S.row_out = (size_t)GWI_LIKE_HEADER + P;
    S.lam.assign((size_t)K * P, 0.0);
    S.out.assign((size_t)K * S.row_out, 0.0);
    std::vector<ChainCtx> ctx(K);
    std::vector<int> rc_chain(K, GWI_OK);
    std::vector<gwi_nuts_info> inf(K);
    std::vector<std::thread> threads;
    threads.reserve(K);
    for (int c = 0; c < K; ++c) {
      ctx[c].sh = &S;
      ctx[c].c = c;
      ctx[c].gp.assign(P, 0.0);
      threads.emplace_back([&, c]() {
        try {
          gwi_nuts_opts o = *opts;
          o.seed = opts->seed + c;
          Sampler smp(chain_potential, &ctx[c], dim, (uint64_t)o.seed);
          rc_chain[c] = smp.run(theta0 + (size_t)c * dim, o, samples ? samples + (size_t)c * o.n_samples * dim : nullptr, &inf[c]);
        } catch (...) {
          rc_chain[c] = GWI_ERR_ALLOC;
        }
        S.state[c].store(CH_ENDED, std::memory_order_release);
      });
    }
    // coordinator: a batch goes to the GPU as soon as `width` chains (or all that are still running) stand at a gradient,
    // oldest request first.  width = the chain count the model's plan was laid out for (gwi_model_desc.batch_hint): with
    // batch_hint = n_chains every round evaluates all chains together; with batch_hint = n_chains / 2 the chains fall into
    // two alternating groups -- one computes its leapfrog arithmetic on the host while the other is being evaluated -- and
    // the GPU never waits for the host (measured on B200, cfg2, 16 chains: see DESIGN.md section 5b).
    const int width = std::max(1, std::min<int>(K, gwi_model_batch_hint(p->model)));
    std::vector<double> lam_b((size_t)K * P), out_b((size_t)K * S.row_out);
    std::vector<int> who;
    who.reserve(K);
    for (unsigned spin = 0;; ++spin) {
      int ended = 0, pending = 0;
      for (int c = 0; c < K; ++c) {
        const int st = S.state[c].load(std::memory_order_acquire);
        ended += st == CH_ENDED;
        pending += st == CH_PENDING;
      }
      if (ended == K) break;
      if (pending == 0 || pending < std::min(width, K - ended)) {
        if (spin > 64) std::this_thread::yield();
        continue;
      }
      spin = 0;
      who.clear();
      for (int c = 0; c < K; ++c)
        if (S.state[c].load(std::memory_order_acquire) == CH_PENDING) who.push_back(c);
      std::sort(who.begin(), who.end(), [&](int a, int b) { return S.ticket[a] < S.ticket[b]; });
      if ((int)who.size() > width) who.resize(width);
      const int n = (int)who.size();
      for (int i = 0; i < n; ++i) std::memcpy(lam_b.data() + (size_t)i * P, S.lam.data() + (size_t)who[i] * P, sizeof(double) * P);
      const int rc = gwi_loglike_batch_host(p->model, lam_b.data(), n, &p->opts, out_b.data());
      p->n_evals += n;
      if (rc != GWI_OK && rc != GWI_ERR_RANGE) S.fatal.store(rc, std::memory_order_release);
      for (int i = 0; i < n; ++i) {
        const int c = who[i];
        const double* row = out_b.data() + (size_t)i * S.row_out;
        std::memcpy(S.out.data() + (size_t)c * S.row_out, row, sizeof(double) * S.row_out);
        S.rc_eval[c] = (rc != GWI_OK && rc != GWI_ERR_RANGE) ? rc : (row[GWI_LIKE_STATUS] != 0.0 ? (int)GWI_ERR_RANGE : (int)GWI_OK);
        S.state[c].store(CH_READY, std::memory_order_release);
      }
    }
    for (auto& t : threads) t.join();
    for (int c = 0; c < K; ++c) {
      if (info) {
        info[c] = inf[c];
        info[c].n_evals = ctx[c].n_evals;
      }
    }
    if (S.fatal.load() != GWI_OK) return S.fatal.load();  // (gwi_last_error holds the failing call's message: it ran on this thread)
    for (int c = 0; c < K; ++c)
      if (rc_chain[c] != GWI_OK) {
        set_error("NUTS chain " + std::to_string(c) + ": the potential is not finite at the starting point, or out of memory");
        return rc_chain[c];
      }
    return GWI_OK;
  } catch (const std::bad_alloc&) {
    set_error("out of host memory in the sampler");
    return GWI_ERR_ALLOC;
  } catch (const std::system_error&) {
    set_error("could not start the chain threads");
    return GWI_ERR_ALLOC;
  }
}

}  // extern "C"
