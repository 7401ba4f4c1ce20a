// lkgpu_kriging.cpp -- see lkgpu_kriging.hpp.  Mirrors, step by step, Kriging::fit of the reference
// (src/lib/Kriging.cpp:1591-2215): fit_setup_impl, bounds, the start-point stream, reparametrisation, one
// lbfgsb::Optimizer run per start with the restart rule, argmin with the strict '<' tie rule, and the commit
// formulas.  The objective itself (populate_Model + reductions) is one lkgpu_objective_fun call.
#include "lkgpu_kriging.hpp"
#include "lkgpu_comm.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <exception>
#include <limits>
#include <mutex>
#include <random>
#include <stdexcept>
#include <thread>

#include "../../include/lkgpu.h"
#include "lbfgsb_cpp/lbfgsb.hpp"

// lbfgsb::Optimizer::minimize finds data(T&) by ADL: put the overload into Armadillo's namespace.
namespace arma {
inline double* data(vec& x) { return x.memptr(); }
}  // namespace arma

namespace lkgpu {

namespace {
constexpr double NUGGET_ALPHA_LOWER = 1e-3;  // Kriging.cpp:678

// The f2c'd Lbfgsb.3.0 (dependencies/lbfgsb_cpp/Lbfgsb.3.0/{lbfgsb,linpack,blas}.c) keeps its scratch variables in
// `static` locals, so setulb must never run on two threads at once (all state that persists between its calls lives
// in the optimiser's own wa / iwa / *save arrays).  Concurrent multistart workers therefore hold this mutex whenever
// they are inside lbfgsb::Optimizer::minimize and drop it only for the duration of an objective evaluation -- the
// part that runs on the GPU and is worth overlapping.
std::mutex g_lbfgsb_mutex;
std::atomic<long long> g_fit_seq{0};  // key of a sharded fit's start queue (lkgpu_comm.hpp)
bool env_is_one(const char* name) {
  const char* v = getenv(name);
  return v && v[0] == '1';
}

void check(int rc) {
  if (rc != 0) throw std::runtime_error(lkgpu_last_error());
}
int kernel_id(const std::string& k) {
  if (k == "gauss") return LKGPU_KERNEL_GAUSS;
  if (k == "exp") return LKGPU_KERNEL_EXP;
  if (k == "matern3_2") return LKGPU_KERNEL_MATERN32;
  if (k == "matern5_2") return LKGPU_KERNEL_MATERN52;
  throw std::invalid_argument("Unsupported covariance kernel: " + k);
}
int objective_id(const std::string& o) {
  if (o == "LL") return LKGPU_OBJ_LL;
  if (o == "LOO") return LKGPU_OBJ_LOO;
  if (o == "LMP") return LKGPU_OBJ_LMP;
  throw std::invalid_argument("Unsupported fit objective: " + o + " (supported here: LL, LOO, LMP)");
}
// Optim::parse_method (Optim.cpp:151-177): "BFGS" -> 1, "BFGS20" -> 20
int parse_multistart(const std::string& optim) {
  if (optim.rfind("BFGS", 0) != 0) throw std::runtime_error("Unsupported optim: " + optim + " (supported are: none, BFGS[#])");
  std::string digits;
  for (size_t i = 4; i < optim.size() && std::isdigit((unsigned char)optim[i]); ++i) digits += optim[i];
  return digits.empty() ? 1 : std::max(1, std::stoi(digits));
}
// Random (Random.cpp:18-43): process-global std::mt19937(123) + uniform_real_distribution, column-major fill
struct ReferenceRandom {
  std::mt19937 engine{123};
  std::uniform_real_distribution<double> dist{0.0, 1.0};
  arma::mat randu_mat(arma::uword n, arma::uword m) {
    arma::mat out(n, m);
    out.imbue([&]() { return dist(engine); });
    return out;
  }
  arma::vec randu_vec(arma::uword n) {
    arma::vec out(n);
    out.imbue([&]() { return dist(engine); });
    return out;
  }
};
}  // namespace

arma::mat regression_model_matrix(const std::string& regmodel, const arma::mat& X) {
  const arma::uword n = X.n_rows, d = X.n_cols;
  if (regmodel == "none") return arma::mat(n, 0);
  std::vector<arma::vec> cols;
  cols.push_back(arma::ones<arma::vec>(n));
  if (regmodel == "constant") {
  } else if (regmodel == "linear") {
    for (arma::uword i = 0; i < d; ++i) cols.push_back(X.col(i));
  } else if (regmodel == "interactive" || regmodel == "quadratic") {
    for (arma::uword i = 0; i < d; ++i) {
      cols.push_back(X.col(i));
      const arma::uword upto = regmodel == "interactive" ? i : i + 1;
      for (arma::uword j = 0; j < upto; ++j) cols.push_back(X.col(i) % X.col(j));
    }
  } else {
    throw std::invalid_argument("Unsupported regression model: " + regmodel);
  }
  arma::mat F(n, cols.size());
  for (size_t c = 0; c < cols.size(); ++c) F.col(c) = cols[c];
  return F;
}

Kriging::Kriging(const std::string& kernel, NoiseModel noise_model, int device)
    : m_kernel(kernel), m_noise_model(noise_model), m_device(device) {
  kernel_id(kernel);
}
Kriging::~Kriging() { close(); }
void Kriging::close() {
  if (m_h) lkgpu_destroy(m_h);
  m_h = nullptr;
}

// ---- reparametrisation: Optim::reparam_* (Optim.cpp:53-60) and nugget_reparam_* (Kriging.cpp:678-702) ----
arma::vec Kriging::reparam_to(const arma::vec& v) const {
  if (!config.reparametrize) return v;
  const arma::uword d = m_X.n_cols;
  arma::vec out = v;
  out.head(d) = arma::log(v.head(d));
  if (m_noise_model == NoiseModel::Nugget) out[d] = -std::log(1.0 + NUGGET_ALPHA_LOWER - v[d]);
  else if (m_noise_model == NoiseModel::Heterogeneous) out[d] = std::log(v[d]);
  return out;
}
arma::vec Kriging::reparam_from(const arma::vec& g) const {
  if (!config.reparametrize) return g;
  const arma::uword d = m_X.n_cols;
  arma::vec out = g;
  out.head(d) = arma::exp(g.head(d));
  if (m_noise_model == NoiseModel::Nugget) out[d] = 1.0 + NUGGET_ALPHA_LOWER - std::exp(-g[d]);
  else if (m_noise_model == NoiseModel::Heterogeneous) out[d] = std::exp(g[d]);
  return out;
}
arma::vec Kriging::reparam_deriv(const arma::vec& v, const arma::vec& grad) const {
  if (!config.reparametrize) return grad;
  const arma::uword d = m_X.n_cols;
  arma::vec out = grad % v;
  if (m_noise_model == NoiseModel::Nugget) out[d] = grad[d] * (1.0 + NUGGET_ALPHA_LOWER - v[d]);
  return out;
}

void Kriging::push_params() {
  check(lkgpu_set_params(m_h, m_est_sigma2, m_sigma2, m_est_nugget, m_nugget, m_alpha));
  // fixed trend coefficients: the committed z is ystar - M beta (Kriging.cpp:2168-2172)
  check(lkgpu_set_fixed_beta(m_h, m_est_beta ? nullptr : m_beta.memptr()));
}

double Kriging::objective(int obj, const arma::vec& gamma, arma::vec* grad) {
  m_have_scalars = false;  // the device now holds the model at gamma
  ++m_n_eval;
  return objective_on(m_h, obj, gamma, grad);
}

double Kriging::objective_on(void* h, int obj, const arma::vec& gamma, arma::vec* grad) const {
  double val = 0.0;
  arma::vec g(gamma.n_elem, arma::fill::zeros);
  check(lkgpu_objective_fun(h, obj, gamma.memptr(), (int)gamma.n_elem, grad != nullptr, &val,
                            grad ? g.memptr() : nullptr, nullptr));
  if (grad) *grad = g;
  return val;
}

// Number of engine handles with overlapping evaluations for this process's multistart rows (the batched-occupancy
// path of BASELINE cfg 5): a factorisation of n <= 8192 cannot fill 148 SMs, so such fits keep several starts in
// flight, one handle and one host thread each.  set_concurrent_starts(K) overrides; results do not depend on it.
void Kriging::set_comm(ShardComm* comm) {
  m_comm = comm;
  m_rank = comm ? comm->rank() : 0;
  m_world = comm ? comm->world() : 1;
}

int Kriging::concurrency(int n_starts, arma::uword n) const {
  int want = m_concurrent_starts > 0 ? m_concurrent_starts : (n <= 3072 ? 8 : (n <= 8192 ? 4 : 1));
  want = std::max(1, std::min(want, n_starts));
  if (want > 1) {
    unsigned long long free_b = 0, total_b = 0;
    if (lkgpu_mem_info(m_device, &free_b, &total_b) == 0) {
      const unsigned long long N = (n + 127) / 128 * 128;
      const unsigned long long per = 3ull * 8ull * N * N + (64ull << 20);
      want = std::max(1, std::min<int>(want, 1 + (int)(0.7 * (double)free_b / (double)per)));
    }
  }
  return want;
}

void Kriging::model_scalars(const arma::vec& theta, double extra, double* SSE, arma::vec* betahat) {
  if (!(m_have_scalars && m_scalars_extra == extra && m_scalars_theta.n_elem == theta.n_elem
        && arma::all(m_scalars_theta == theta))) {
    lkgpu_out out;
    std::memset(&out, 0, sizeof(out));
    arma::vec b(m_F.n_cols);
    out.betahat = b.memptr();
    check(lkgpu_eval(m_h, LKGPU_OBJ_LL, theta.memptr(), extra, 0, &out));
    m_have_scalars = true;
    m_scalars_theta = theta;
    m_scalars_extra = extra;
    m_scalars_SSE = out.SSEstar;
    m_scalars_beta = b;
  }
  *SSE = m_scalars_SSE;
  *betahat = m_scalars_beta;
}

double Kriging::sigma2_variogram() const {
  // Heterogeneous sigma2 bounds (Kriging.cpp:1784-1797): half the mean squared increment over the ordered pairs
  // (diagonal included) whose squared distance is at least the median -- on the device.
  double v = 0.0;
  check(lkgpu_sigma2_variogram(m_h, &v));
  return v;
}

void Kriging::fit(const arma::vec& y, const arma::mat& X, const std::string& regmodel, bool normalize,
                  const std::string& optim, const std::string& objective, const Parameters& parameters) {
  if (m_noise_model == NoiseModel::Heterogeneous)
    throw std::runtime_error("fit(y, noise, X, ...) requires a noise vector for NoiseModel::Heterogeneous");
  fit_impl(y, nullptr, X, regmodel, normalize, optim, objective, parameters);
}
void Kriging::fit(const arma::vec& y, const arma::vec& noise, const arma::mat& X, const std::string& regmodel,
                  bool normalize, const std::string& optim, const std::string& objective, const Parameters& parameters) {
  if (m_noise_model != NoiseModel::Heterogeneous)
    throw std::runtime_error("fit(y, noise, X, ...) requires NoiseModel::Heterogeneous");
  fit_impl(y, &noise, X, regmodel, normalize, optim, objective, parameters);
}

void Kriging::fit_impl(const arma::vec& y, const arma::vec* noise, const arma::mat& X, const std::string& regmodel,
                       bool normalize, const std::string& optim, const std::string& objective_name,
                       const Parameters& prm) {
  const arma::uword n = X.n_rows, d = X.n_cols;
  if (y.n_elem != n)
    throw std::runtime_error("Dimension of new data should be the same:\n X: (" + std::to_string(n) + "x" +
                             std::to_string(d) + "), y: (" + std::to_string(y.n_elem) + ")");
  if (noise && noise->n_elem != n) throw std::runtime_error("noise vector must have the same length as y");
  const int obj = objective_id(objective_name);
  if (obj == LKGPU_OBJ_LOO && m_noise_model != NoiseModel::None)
    throw std::invalid_argument("LOO objective not supported for Nugget/Heterogeneous noise modes");
  if (obj == LKGPU_OBJ_LMP && m_noise_model == NoiseModel::Heterogeneous)
    throw std::invalid_argument("LMP objective not supported for Heterogeneous noise mode");
  m_objective = objective_name;
  m_regmodel = regmodel;
  m_optim = optim;

  // ---- fit_setup_impl (KrigingImpl.cpp:764-841) ----
  m_normalize = normalize;
  if (normalize) {
    m_centerX = arma::min(X, 0);
    m_scaleX = arma::max(X, 0) - arma::min(X, 0);
    m_centerY = y.min();
    m_scaleY = y.max() - y.min();
  } else {
    m_centerX = arma::zeros<arma::rowvec>(d);
    m_scaleX = arma::ones<arma::rowvec>(d);
    m_centerY = 0.0;
    m_scaleY = 1.0;
  }
  m_X = X;
  m_X.each_row() -= m_centerX;
  m_X.each_row() /= m_scaleX;
  m_y = (y - m_centerY) / m_scaleY;
  if (noise) m_noise = *noise;  // stored raw even when normalize = true (quirk (iii))
  m_F = regression_model_matrix(regmodel, m_X);
  const arma::uword p = m_F.n_cols;
  if (p == 0 && obj != LKGPU_OBJ_LL)
    throw std::runtime_error("regmodel='none' (no trend column) is supported for objective='LL' only");
  m_est_beta = true;
  if (!prm.is_beta_estim && prm.beta.has_value() && prm.beta->n_elem > 0) {
    m_est_beta = false;
    m_beta = *prm.beta / (normalize ? m_scaleY : 1.0);
  }
  std::optional<arma::mat> theta0;
  if (prm.theta.has_value()) {
    arma::mat t = *prm.theta;
    if (t.n_cols != d && t.n_rows == d) t = t.t();
    if (normalize) t.each_row() /= m_scaleX;
    if (t.n_cols != d)
      throw std::runtime_error("Dimension of theta should be nx" + std::to_string(d) + " instead of " +
                               std::to_string(t.n_rows) + "x" + std::to_string(t.n_cols));
    theta0 = t;
  }
  const double scaleY2 = normalize ? m_scaleY * m_scaleY : 1.0;

  close();
  check(lkgpu_create(&m_h, m_device, (int)n, (int)d, (int)p, m_X.memptr(), m_y.memptr(), m_F.memptr(),
                     noise ? m_noise.memptr() : nullptr, kernel_id(m_kernel), (int)m_noise_model));
  m_sigma2 = 1.0;
  m_nugget = 0.0;
  m_alpha = 1.0;
  m_is_empty = true;
  m_have_scalars = false;
  m_results.clear();
  m_n_eval = 0;
  const NoiseModel nm = m_noise_model;

  if (optim == "none") {
    if (!theta0.has_value())
      throw std::runtime_error("Theta should be given (1x" + std::to_string(d) + ") matrix, when optim=none");
    m_theta = theta0->row(0).t();
    m_est_theta = false;
    double sigma2 = -1.0;
    m_est_sigma2 = prm.is_sigma2_estim;
    if (prm.sigma2.has_value()) sigma2 = *prm.sigma2 / scaleY2;
    else m_est_sigma2 = true;
    double nugget_param = 0.0, extra = 1.0;
    m_est_nugget = true;
    if (nm == NoiseModel::Nugget) {
      m_est_nugget = prm.is_nugget_estim;
      if (prm.nugget.has_value()) nugget_param = *prm.nugget / scaleY2;
      m_alpha = (sigma2 > 0 && sigma2 + nugget_param > 0) ? sigma2 / (sigma2 + nugget_param) : 1.0 - NUGGET_ALPHA_LOWER;
      extra = m_alpha;
    } else if (nm == NoiseModel::Heterogeneous) {
      extra = sigma2 > 0 ? sigma2 : m_sigma2;
    }
    double SSE;
    arma::vec betahat;
    model_scalars(m_theta, extra, &SSE, &betahat);
    check(lkgpu_commit_model(m_h));
    m_commit_extra = extra;
    m_is_empty = false;
    if (m_est_beta) m_beta = betahat;
    if (nm == NoiseModel::Nugget) {
      if (m_est_sigma2) {
        const double tv = SSE / n;
        m_sigma2 = m_alpha * tv;
        m_nugget = m_est_nugget ? (1.0 - m_alpha) * tv : nugget_param;
      } else {
        m_sigma2 = sigma2;
        m_nugget = m_est_nugget ? 0.0 : nugget_param;
      }
    } else if (m_est_sigma2) {
      m_sigma2 = SSE / n;
    } else {
      m_sigma2 = sigma2;
    }
    push_params();
    return;
  }

  // ---- bounds, starts (Kriging.cpp:1703-1829) ----
  const int multistart_req = parse_multistart(optim);
  arma::vec theta_lower(d), theta_upper(d);
  check(lkgpu_theta_bounds(m_h, config.theta_lower_factor, config.theta_upper_factor,
                           config.variogram_bounds_heuristic, theta_lower.memptr(), theta_upper.memptr()));
  ReferenceRandom rng;
  arma::uword multistart = multistart_req;
  arma::mat theta0_rand = rng.randu_mat(multistart, d);
  theta0_rand.each_row() %= (theta_upper - theta_lower).t();
  theta0_rand.each_row() += theta_lower.t();
  arma::mat starts;
  if (theta0.has_value()) {
    multistart = std::max<arma::uword>(multistart, theta0->n_rows);
    starts = arma::join_cols(*theta0, theta0_rand);
    starts = starts.rows(0, multistart - 1);
  } else {
    starts = theta0_rand;
  }
  arma::vec extra0;
  double extra_lo = 0.0, extra_up = 1.0;
  if (nm == NoiseModel::Nugget) {
    extra_lo = NUGGET_ALPHA_LOWER;
    extra_up = 1.0;
    if (prm.sigma2.has_value() && prm.nugget.has_value()) {
      const double s = *prm.sigma2, nu = *prm.nugget;
      extra0 = arma::vec{(s > 0 && s + nu > 0) ? s / (s + nu) : extra_lo + (extra_up - extra_lo) * 0.5};
    } else {
      extra0 = extra_lo + (extra_up - extra_lo) * (1.0 - arma::pow(rng.randu_vec(starts.n_rows), 3.0));
    }
  } else if (nm == NoiseModel::Heterogeneous) {
    const double s2v = sigma2_variogram();
    extra_lo = 0.1 * (s2v - m_noise.max());
    extra_up = 10.0 * (s2v - m_noise.min());
    if (prm.sigma2.has_value()) extra0 = arma::vec{*prm.sigma2 / (normalize ? m_scaleY : 1.0)};
    else extra0 = extra_lo + (extra_up - extra_lo) * rng.randu_vec(starts.n_rows);
  }
  const arma::uword gd = d + (nm == NoiseModel::None ? 0 : 1);
  arma::vec lo_full = theta_lower, up_full = theta_upper;
  if (gd > d) {
    lo_full.resize(gd);
    up_full.resize(gd);
    lo_full[d] = extra_lo;
    up_full[d] = extra_up;
  }
  const arma::vec gamma_lower = reparam_to(lo_full), gamma_upper = reparam_to(up_full);

  // ---- estimation flags (Kriging.cpp:1831-1849) ----
  m_est_sigma2 = prm.is_sigma2_estim;
  if (!m_est_sigma2 && prm.sigma2.has_value()) m_sigma2 = *prm.sigma2 / scaleY2;
  else m_est_sigma2 = true;
  m_est_nugget = true;
  if (nm == NoiseModel::Nugget) {
    m_est_nugget = prm.is_nugget_estim;
    if (!m_est_nugget && prm.nugget.has_value()) m_nugget = *prm.nugget / scaleY2;
    else m_est_nugget = true;
  }
  push_params();

  const double sign = obj == LKGPU_OBJ_LOO ? 1.0 : -1.0;
  // fit_ofn on engine handle h; `locked`: the caller holds g_lbfgsb_mutex, which is dropped while the device works
  auto fit_ofn = [&](void* h, const arma::vec& gamma, arma::vec* grad_out, bool locked) -> double {
    const arma::vec v = reparam_from(gamma);
    arma::vec g;
    if (locked) g_lbfgsb_mutex.unlock();
    double val;
    try {
      val = objective_on(h, obj, v, grad_out ? &g : nullptr);
    } catch (...) {
      if (locked) g_lbfgsb_mutex.lock();
      throw;
    }
    if (locked) g_lbfgsb_mutex.lock();
    if (grad_out) *grad_out = sign * reparam_deriv(v, g);
    return sign * val;
  };
  const double nn = (double)n * (double)n;
  const double pgtol = obj == LKGPU_OBJ_LOO ? config.gradient_tolerance / nn : config.gradient_tolerance;
  const double factr = obj == LKGPU_OBJ_LOO ? config.objective_rel_tolerance / 1e-13 / nn
                                              : config.objective_rel_tolerance / 1e-13;

  // ---- one L-BFGS-B run per start (optimize_worker, Kriging.cpp:1904-2084) ----
  auto optimize_worker = [&](arma::uword s, void* h) -> StartResult {
    StartResult res;
    res.start_index = (int)s;
    res.objective_value = std::numeric_limits<double>::infinity();
    int n_eval = 0;
    std::unique_lock<std::mutex> lk(g_lbfgsb_mutex);
    try {
      const arma::vec theta_start = starts.row(s % multistart).t();
      arma::vec full = theta_start;
      if (gd > d) {
        full.resize(gd);
        full[d] = extra0[s % extra0.n_elem];
      }
      arma::vec gamma_tmp = reparam_to(full);
      arma::vec lo_loc = arma::min(gamma_tmp, gamma_lower), up_loc = arma::max(gamma_tmp, gamma_upper);
      // (the reference's warm-up populate_Model at theta_start, Kriging.cpp:1943-1948, is skipped: its result is
      //  unconditionally overwritten by the first fit_ofn call at the same point)
      lbfgsb::Optimizer optimizer{(unsigned int)gd};
      optimizer.iprint = -1;
      optimizer.max_iter = config.max_iteration;
      optimizer.pgtol = pgtol;
      optimizer.factr = factr;
      std::vector<int> bounds_type(gd, 2);
      int retry = 0;
      double best_f = std::numeric_limits<double>::infinity();
      arma::vec best_gamma = gamma_tmp;
      while (retry <= config.max_restart) {
        auto r = optimizer.minimize(
            [&](const arma::vec& x, arma::vec& grad) -> double {
              ++n_eval;
              return fit_ofn(h, x, &grad, true);
            },
            gamma_tmp, lo_loc.memptr(), up_loc.memptr(), bounds_type.data());
        if (r.f_opt < best_f) {
          best_f = r.f_opt;
          best_gamma = gamma_tmp;
        }
        const arma::vec theta_part = reparam_from(gamma_tmp).head(d);
        const double sol_to_lb = arma::min(arma::abs(theta_part - theta_lower));
        if (retry < config.max_restart
            && (r.task.rfind("ABNORMAL_TERMINATION_IN_LNSRCH", 0) == 0 || r.num_iters <= 2
                || sol_to_lb < arma::datum::eps || r.f_opt > best_f)) {
          arma::vec restart = (theta_start + theta_lower) / std::pow(2.0, retry + 1);
          if (gd > d) {
            restart.resize(gd);
            restart[d] = extra0[s % extra0.n_elem];
          }
          gamma_tmp = reparam_to(restart);
          lo_loc = arma::min(gamma_tmp, lo_loc);
          up_loc = arma::max(gamma_tmp, up_loc);
          ++retry;
        } else {
          break;
        }
      }
      ++n_eval;
      res.objective_value = fit_ofn(h, best_gamma, nullptr, true);  // final evaluation (Kriging.cpp:2044)
      res.gamma = best_gamma;
      res.success = true;
      res.retries = retry;
    } catch (const std::exception& e) {  // one failed start must not kill the fit (Kriging.cpp:2075-2081)
      res.success = false;
      res.error_message = e.what();
    }
    res.n_eval = n_eval;
    return res;
  };

  // this process's starts, several of them in flight when n is mid-size: static {s : s mod world ==
  // rank}, or -- sharded fit with more starts than processes -- drawn from the shared ticket counter
  const bool dynamic = m_comm != nullptr && (int)multistart > m_world && !env_is_one("LKGPU_STATIC_STARTS");
  std::vector<arma::uword> mine;
  if (!dynamic)
    for (arma::uword s = 0; s < multistart; ++s)
      if ((int)(s % (arma::uword)m_world) == m_rank) mine.push_back(s);
  const long long queue_key = ++g_fit_seq;  // every process runs the same sequence of fits
  std::atomic<size_t> next_static{0};
  auto next_start = [&]() -> long long {
    if (dynamic) {
      const long long t = m_comm->next_ticket(queue_key);
      return t < (long long)multistart ? t : -1;
    }
    const size_t k = next_static.fetch_add(1);
    return k < mine.size() ? (long long)mine[k] : -1;
  };
  const int share = dynamic ? (int)((multistart + m_world - 1) / m_world) : (int)mine.size();
  const int ncon = concurrency(share, n);
  m_last_concurrency = ncon;
  std::vector<StartResult> results;
  std::mutex results_mutex;
  auto run_starts = [&](void* h) {
    for (long long s = next_start(); s >= 0; s = next_start()) {
      StartResult r = optimize_worker((arma::uword)s, h);
      std::lock_guard<std::mutex> lk(results_mutex);
      results.push_back(std::move(r));
    }
  };
  if (ncon <= 1) {
    run_starts(m_h);
  } else {
    // one engine handle (own workspaces, own CUDA streams) and one host thread per start in flight
    std::vector<void*> handles{m_h};
    try {
      for (int w = 1; w < ncon; ++w) {
        void* h = nullptr;
        check(lkgpu_create(&h, m_device, (int)n, (int)d, (int)p, m_X.memptr(), m_y.memptr(), m_F.memptr(),
                           noise ? m_noise.memptr() : nullptr, kernel_id(m_kernel), (int)m_noise_model));
        handles.push_back(h);
        check(lkgpu_set_params(h, m_est_sigma2, m_sigma2, m_est_nugget, m_nugget, m_alpha));
      }
      std::vector<std::thread> pool;
      std::exception_ptr failure;  // e.g. the start queue's connection lost: rethrown on this thread after the join
      for (int w = 0; w < ncon; ++w)
        pool.emplace_back([&, w]() {
          try {
            run_starts(handles[w]);
          } catch (...) {
            std::lock_guard<std::mutex> lk(results_mutex);
            if (!failure) failure = std::current_exception();
          }
        });
      for (auto& t : pool) t.join();
      if (failure) std::rethrow_exception(failure);
    } catch (...) {
      for (size_t w = 1; w < handles.size(); ++w) lkgpu_destroy(handles[w]);
      throw;
    }
    for (size_t w = 1; w < handles.size(); ++w) lkgpu_destroy(handles[w]);
  }
  std::sort(results.begin(), results.end(),
            [](const StartResult& a, const StartResult& b) { return a.start_index < b.start_index; });
  m_local_n_eval = 0;
  m_local_starts.clear();
  for (const StartResult& r : results) {
    m_local_n_eval += r.n_eval;
    m_local_starts.push_back(r.start_index);
  }
  if (m_comm != nullptr) {
    // one all-gather of a row per start: [start, success, objective, n_eval, retries, gamma]; afterwards every
    // process holds every start's result (the owner's bits) in start order
    const size_t width = 5 + gd;
    std::vector<double> rows;
    for (const StartResult& r : results) {
      rows.push_back((double)r.start_index);
      rows.push_back(r.success ? 1.0 : 0.0);
      rows.push_back(r.objective_value);
      rows.push_back((double)r.n_eval);
      rows.push_back((double)r.retries);
      for (arma::uword q = 0; q < gd; ++q) rows.push_back(r.success ? r.gamma[q] : 0.0);
    }
    const std::vector<double> all = m_comm->allgather(rows);
    if (all.size() != width * multistart)
      throw std::runtime_error("sharded fit: " + std::to_string(all.size() / width) + " start results gathered, " +
                               std::to_string(multistart) + " expected");
    results.assign(multistart, StartResult());
    for (size_t k = 0; k < multistart; ++k) {
      const double* row = all.data() + k * width;
      StartResult r;
      r.start_index = (int)row[0];
      r.success = row[1] > 0.5;
      r.objective_value = row[2];
      r.n_eval = (int)row[3];
      r.retries = (int)row[4];
      r.gamma = arma::vec(row + 5, gd);
      if (r.start_index < 0 || r.start_index >= (int)multistart || results[r.start_index].start_index >= 0)
        throw std::runtime_error("sharded fit: start results do not cover the starts exactly once");
      results[r.start_index] = r;
    }
  }
  for (const StartResult& r : results) {
    m_n_eval += r.n_eval;
    m_results.push_back(r);
  }
  m_have_scalars = false;

  if (m_world > 1 && m_comm == nullptr) return;  // set_shard: the caller exchanges start_results() and calls commit(gamma*)
  // ---- argmin over successful starts, strict '<' in start order (Kriging.cpp:2097-2114) ----
  int best = -1;
  double min_ofn = std::numeric_limits<double>::infinity();
  for (size_t k = 0; k < m_results.size(); ++k)
    if (m_results[k].success && m_results[k].objective_value < min_ofn) {
      min_ofn = m_results[k].objective_value;
      best = (int)k;
    }
  if (best < 0) throw std::runtime_error("All " + std::to_string(multistart) + " optimization attempts failed");
  commit(m_results[best].gamma);
}

// ---- commit (Kriging.cpp:2156-2202): the model of the best start is rebuilt on the device by one value-only
//      evaluation at gamma* (bit-reproducible; no n x n matrix leaves the GPU) ----
void Kriging::commit(const arma::vec& best_gamma) {
  const arma::uword n = m_X.n_rows, d = m_X.n_cols, p = m_F.n_cols;
  const NoiseModel nm = m_noise_model;
  const arma::vec v = reparam_from(best_gamma);
  m_theta = v.head(d);
  const bool has_extra = nm != NoiseModel::None;
  const double extra_param = has_extra ? v[d] : 0.0;
  double commit_extra = has_extra ? extra_param : 1.0;
  // the committed model is the one the last fit_ofn(best_gamma) built, and _logLikelihood overrides the optimiser's
  // extra parameter there when it is fixed (Kriging.cpp:221-234)
  if (m_objective == "LL") {
    if (nm == NoiseModel::Heterogeneous && !m_est_sigma2) commit_extra = m_sigma2;
    else if (nm == NoiseModel::Nugget && !m_est_sigma2 && !m_est_nugget) commit_extra = m_sigma2 / (m_sigma2 + m_nugget);
  }
  double SSE;
  arma::vec betahat;
  model_scalars(m_theta, commit_extra, &SSE, &betahat);
  check(lkgpu_commit_model(m_h));
  m_est_theta = true;
  m_commit_extra = commit_extra;
  m_is_empty = false;
  if (m_est_beta) m_beta = betahat;
  if (nm == NoiseModel::Nugget) {
    m_alpha = extra_param;
    if (m_est_sigma2) {
      if (m_est_nugget) {
        const double tv = SSE / n;
        m_sigma2 = m_alpha * tv;
        if (m_objective == "LMP") m_sigma2 = m_sigma2 * n / (double)(n - p - 2);
        m_nugget = m_sigma2 / m_alpha - m_sigma2;
      } else {
        m_sigma2 = m_nugget * m_alpha / (1.0 - m_alpha);
      }
    } else if (m_est_nugget) {
      m_nugget = m_sigma2 * (1.0 - m_alpha) / m_alpha;
    }
  } else if (nm == NoiseModel::Heterogeneous) {
    if (m_est_sigma2) m_sigma2 = extra_param;
  } else if (m_est_sigma2) {
    m_sigma2 = SSE / n;
    if (m_objective == "LMP") m_sigma2 = SSE / (double)(n - p);
  }
  push_params();
}

void Kriging::need_model() {
  if (m_is_empty || !m_h) throw std::runtime_error("Kriging model is not fitted");
  // make sure the live model on the device is the committed one (objective calls may have replaced it)
  check(lkgpu_restore_model(m_h));
  m_have_scalars = false;
}

// ---- update (reference src/lib/Kriging.cpp:2425-2660) ----
void Kriging::update(const arma::vec& y_u, const arma::vec& noise_u, const arma::mat& X_u, bool refit) {
  if (m_noise_model != NoiseModel::Heterogeneous)
    throw std::runtime_error("update(y, noise, X) requires NoiseModel::Heterogeneous");
  if (m_is_empty || !m_h) throw std::runtime_error("Kriging model is not fitted");
  if (y_u.n_elem != X_u.n_rows)
    throw std::runtime_error("Dimension of new data should be the same:\n X: (" + std::to_string(X_u.n_rows) + "x" +
                             std::to_string(X_u.n_cols) + "), y: (" + std::to_string(y_u.n_elem) + ")");
  if (noise_u.n_elem != y_u.n_elem) throw std::runtime_error("noise_u must have the same length as y_u");
  // a new fit on the joined, de-normalised data, started from the current parameters (:2645-2658)
  const arma::vec y_all = arma::join_cols(m_y * m_scaleY + m_centerY, y_u);
  const arma::vec noise_all = arma::join_cols(m_noise * m_scaleY * m_scaleY, noise_u);
  arma::mat Xd = m_X;
  Xd.each_row() %= m_scaleX;
  Xd.each_row() += m_centerX;
  const arma::mat X_all = arma::join_cols(Xd, X_u);
  Parameters prm;
  prm.sigma2 = m_sigma2 * m_scaleY * m_scaleY;
  prm.is_sigma2_estim = m_est_sigma2;
  prm.theta = arma::mat(m_theta.t() % m_scaleX);
  prm.is_theta_estim = m_est_theta;
  if (!m_est_beta) prm.beta = m_beta * m_scaleY;
  prm.is_beta_estim = m_est_beta;
  const std::string optim = refit ? m_optim : "none", objective_name = m_objective, regmodel = m_regmodel;
  fit_impl(y_all, &noise_all, X_all, regmodel, m_normalize, optim, objective_name, prm);
}

void Kriging::update(const arma::vec& y_u, const arma::mat& X_u, bool refit) {
  if (m_is_empty || !m_h) throw std::runtime_error("Kriging model is not fitted");
  const arma::uword d = m_X.n_cols;
  if (y_u.n_elem != X_u.n_rows)
    throw std::runtime_error("Dimension of new data should be the same:\n X: (" + std::to_string(X_u.n_rows) + "x" +
                             std::to_string(X_u.n_cols) + "), y: (" + std::to_string(y_u.n_elem) + ")");
  if (X_u.n_cols != d)
    throw std::runtime_error("Dimension of new data should be the same:\n X: (...x" + std::to_string(d) +
                             "), new X: (...x" + std::to_string(X_u.n_cols) + ")");
  if (m_noise_model == NoiseModel::Heterogeneous)
    throw std::runtime_error("update(y, noise, X) requires a noise vector for NoiseModel::Heterogeneous");
  const NoiseModel nm = m_noise_model;
  m_used_block = false;
  if (refit && m_optim != "none" && nm == NoiseModel::Nugget) {
    // Nugget refit: a new fit on the de-normalised joined data (:2443-2468)
    const arma::vec y_all = arma::join_cols(m_y * m_scaleY + m_centerY, y_u);
    arma::mat Xd = m_X;
    Xd.each_row() %= m_scaleX;
    Xd.each_row() += m_centerX;
    const arma::mat X_all = arma::join_cols(Xd, X_u);
    Parameters prm;
    if (!(m_est_beta && m_est_nugget && m_est_sigma2 && m_est_theta)) {
      prm.sigma2 = m_sigma2 * m_scaleY * m_scaleY;
      prm.is_sigma2_estim = m_est_sigma2;
      prm.theta = arma::mat(m_theta.t() % m_scaleX);
      prm.is_theta_estim = m_est_theta;
      prm.nugget = m_nugget * m_scaleY * m_scaleY;
      prm.is_nugget_estim = m_est_nugget;
      if (!m_est_beta) {
        prm.beta = m_beta * m_scaleY;
        prm.is_beta_estim = false;
      }
    }
    const std::string optim = m_optim, objective_name = m_objective, regmodel = m_regmodel;
    fit_impl(y_all, nullptr, X_all, regmodel, m_normalize, optim, objective_name, prm);
    return;
  }

  // ---- extend the data with the model's own normalisation; the device keeps the committed factor ----
  need_model();
  arma::mat Xn_u = X_u;
  Xn_u.each_row() -= m_centerX;
  Xn_u.each_row() /= m_scaleX;
  const arma::vec yn_u = (y_u - m_centerY) / m_scaleY;
  const arma::mat F_u = regression_model_matrix(m_regmodel, Xn_u);
  check(lkgpu_append_data(m_h, (int)Xn_u.n_rows, Xn_u.memptr(), yn_u.memptr(), F_u.memptr(), nullptr));
  m_have_scalars = false;
  m_X = arma::join_cols(m_X, Xn_u);
  m_y = arma::join_cols(m_y, yn_u);
  m_F = arma::join_cols(m_F, F_u);
  const arma::uword n = m_X.n_rows, p = m_F.n_cols;
  const double extra = nm == NoiseModel::Nugget ? m_alpha : 1.0;
  double SSE;
  arma::vec betahat;

  if (!(refit && m_optim != "none")) {
    // update_no_refit_impl (KrigingImpl.cpp:576-625): make_Model(m_theta) -- update_eligible -> block extension
    model_scalars(m_theta, extra, &SSE, &betahat);
    m_used_block = lkgpu_last_eval_was_update(m_h) != 0;
    check(lkgpu_commit_model(m_h));
    m_commit_extra = extra;
    if (m_est_beta) m_beta = betahat;
    if (m_est_sigma2) m_sigma2 = SSE / n;
    push_params();
    return;
  }

  // ---- warm restart (:2470-2623): a single L-BFGS-B run from the current theta on the extended data ----
  const int obj = objective_id(m_objective);
  arma::vec theta_lower(d), theta_upper(d);
  check(lkgpu_theta_bounds(m_h, config.theta_lower_factor, config.theta_upper_factor,
                           config.variogram_bounds_heuristic, theta_lower.memptr(), theta_upper.memptr()));
  arma::vec gamma_start = reparam_to(m_theta);
  arma::vec gamma_lower = arma::min(gamma_start, reparam_to(theta_lower));
  arma::vec gamma_upper = arma::max(gamma_start, reparam_to(theta_upper));
  model_scalars(m_theta, extra, &SSE, &betahat);  // the warm-up populate_Model at m_theta (:2544-2547)
  m_used_block = lkgpu_last_eval_was_update(m_h) != 0;
  const double sign = obj == LKGPU_OBJ_LOO ? 1.0 : -1.0;
  const double nn = (double)n * (double)n;
  lbfgsb::Optimizer optimizer{(unsigned int)d};
  optimizer.iprint = -1;
  optimizer.max_iter = config.max_iteration;
  optimizer.pgtol = obj == LKGPU_OBJ_LOO ? config.gradient_tolerance / nn : config.gradient_tolerance;
  optimizer.factr = obj == LKGPU_OBJ_LOO ? config.objective_rel_tolerance / 1e-13 / nn
                                         : config.objective_rel_tolerance / 1e-13;
  std::vector<int> bounds_type(d, 2);
  arma::vec gamma_tmp = gamma_start;
  optimizer.minimize(
      [&](const arma::vec& x, arma::vec& grad) -> double {
        const arma::vec v = reparam_from(x);
        arma::vec g;
        const double val = objective(obj, v, &g);
        grad = sign * reparam_deriv(v, g);
        return sign * val;
      },
      gamma_tmp, gamma_lower.memptr(), gamma_upper.memptr(), bounds_type.data());
  m_theta = reparam_from(gamma_tmp).head(d);
  m_est_theta = true;
  // (the reference commits whatever model the optimiser's last evaluation left in km; here the model at the returned
  //  point is rebuilt by one value-only evaluation -- the same point unless the line search failed)
  model_scalars(m_theta, extra, &SSE, &betahat);
  check(lkgpu_commit_model(m_h));
  m_commit_extra = extra;
  if (m_est_beta) m_beta = betahat;
  if (m_est_sigma2) m_sigma2 = m_objective == "LMP" ? SSE / (double)(n - p) : SSE / n;
  push_params();
}

arma::vec Kriging::gamma_full(const arma::vec& theta) const {
  const arma::uword d = m_X.n_cols;
  if (theta.n_elem == d && m_noise_model != NoiseModel::None) {
    arma::vec g = theta;
    g.resize(d + 1);
    g[d] = m_noise_model == NoiseModel::Nugget ? m_alpha : m_sigma2;
    return g;
  }
  return theta;
}

std::tuple<double, arma::vec> Kriging::logLikelihoodFun(const arma::vec& theta, bool return_grad) {
  arma::vec g;
  const double v = objective(LKGPU_OBJ_LL, gamma_full(theta), return_grad ? &g : nullptr);
  return {v, g};
}
std::tuple<double, arma::vec> Kriging::leaveOneOutFun(const arma::vec& theta, bool return_grad) {
  if (m_noise_model != NoiseModel::None)
    throw std::invalid_argument("LOO objective not supported for Nugget/Heterogeneous noise modes");
  arma::vec g;
  const double v = objective(LKGPU_OBJ_LOO, theta, return_grad ? &g : nullptr);
  return {v, g};
}
std::tuple<double, arma::vec> Kriging::logMargPostFun(const arma::vec& theta, bool return_grad) {
  if (m_noise_model == NoiseModel::Heterogeneous)
    throw std::invalid_argument("LMP objective not supported for Heterogeneous noise mode");
  arma::vec g;
  const double v = objective(LKGPU_OBJ_LMP, gamma_full(theta), return_grad ? &g : nullptr);
  return {v, g};
}

// ---- predict mean / stdev (Kriging.cpp:2240-2285 -> KrigingImpl.cpp:145-243) ----
std::tuple<arma::vec, arma::vec> Kriging::predict(const arma::mat& X_n, bool return_stdev) {
  need_model();
  const arma::uword d = m_X.n_cols, n_o = m_F.n_rows, p = m_F.n_cols, m = X_n.n_rows;
  if (X_n.n_cols != d)
    throw std::runtime_error("Predict locations have wrong dimension: " + std::to_string(X_n.n_cols) + " instead of " +
                             std::to_string(d));
  arma::mat Xn = X_n;
  Xn.each_row() -= m_centerX;
  Xn.each_row() /= m_scaleX;
  const arma::mat Fn = regression_model_matrix(m_regmodel, Xn);
  const double lmp_scale = m_objective == "LMP" ? (double)(n_o - p) / ((double)(n_o - p) - 2.0) : 1.0;
  double factor = 1.0, var_scale = m_sigma2 * lmp_scale;
  if (m_noise_model == NoiseModel::Nugget) {
    factor = m_alpha;
    var_scale = m_sigma2 * lmp_scale / m_alpha;
  }
  arma::vec mean(m), var(m);
  check(lkgpu_predict(m_h, (int)m, Xn.memptr(), Fn.memptr(), m_beta.memptr(), factor, mean.memptr(),
                      return_stdev ? var.memptr() : nullptr));
  mean = m_centerY + m_scaleY * mean;
  arma::vec sd;
  if (return_stdev) {
    var.transform([](double x) { return (x != x || x < 0.0) ? 0.0 : x; });
    sd = arma::sqrt(var * var_scale * m_scaleY * m_scaleY);
  }
  return {mean, sd};
}

arma::mat Kriging::T() {
  need_model();
  arma::mat out(m_X.n_rows, m_X.n_rows);
  check(lkgpu_export(m_h, LKGPU_EXPORT_L, out.memptr()));
  return out;
}
arma::mat Kriging::M() {
  need_model();
  arma::mat out(m_X.n_rows, m_F.n_cols);
  check(lkgpu_export(m_h, LKGPU_EXPORT_FSTAR, out.memptr()));
  return out;
}
arma::vec Kriging::z() {
  need_model();
  arma::vec out(m_X.n_rows);
  check(lkgpu_export(m_h, LKGPU_EXPORT_Z, out.memptr()));
  return out;
}
arma::mat Kriging::circ() {
  need_model();
  arma::mat out(m_F.n_cols, m_F.n_cols);
  check(lkgpu_export(m_h, LKGPU_EXPORT_RSTAR, out.memptr()));
  return out;
}

}  // namespace lkgpu
