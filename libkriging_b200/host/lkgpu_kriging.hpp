// lkgpu_kriging.hpp -- C++ host side above the C ABI of include/lkgpu.h.
//
// The reference's host is C++: it keeps an Armadillo-facing API and drives the lbfgsb_cpp optimiser loop on the
// CPU (src/lib/include/libKriging/Kriging.hpp:113-274, src/lib/Kriging.cpp:1591-2215).  This class mirrors that
// surface for the fit / objective / predict path -- same method names, argument meaning and error texts -- and
// sends every objective evaluation to the device through lkgpu_objective_fun.  It uses the same two third-party
// dependencies as the reference's host, from where the reference vendors them (dependencies/armadillo-code,
// dependencies/lbfgsb_cpp), so the L-BFGS-B iterates are produced by the very code the reference runs.
// Everything below the objective call is liblkgpu.so; nothing here touches CUDA.
#pragma once
#include <armadillo>
#include <optional>
#include <string>
#include <tuple>
#include <vector>

namespace lkgpu {

class ShardComm;  // lkgpu_comm.hpp

struct OptimConfig {  // Optim:: statics of the reference (src/lib/Optim.cpp:39-141)
  bool reparametrize = true;
  double theta_lower_factor = 0.02, theta_upper_factor = 10.0;
  bool variogram_bounds_heuristic = true;
  int max_restart = 10, max_iteration = 20;
  double gradient_tolerance = 1e-3, objective_rel_tolerance = 1e-3;
};

struct KrigingParameters {  // Kriging::Parameters of the reference (Kriging.hpp:71-83)
  std::optional<double> sigma2;
  bool is_sigma2_estim = true;
  std::optional<arma::mat> theta;
  bool is_theta_estim = true;
  std::optional<arma::vec> beta;
  bool is_beta_estim = true;
  std::optional<double> nugget;
  bool is_nugget_estim = true;
};

class Kriging {
 public:
  enum class NoiseModel { None = 0, Nugget = 1, Heterogeneous = 2 };  // Kriging.hpp:45-49
  using Parameters = KrigingParameters;
  struct StartResult {
    int start_index = -1;
    bool success = false;
    double objective_value = 0.0;
    arma::vec gamma;
    int n_eval = 0, retries = 0;
    std::string error_message;
  };

  explicit Kriging(const std::string& kernel, NoiseModel noise_model = NoiseModel::None, int device = 0);
  ~Kriging();
  Kriging(const Kriging&) = delete;
  Kriging& operator=(const Kriging&) = delete;

  OptimConfig config;
  // multistart sharding: this process runs the starts {s : s mod world == rank}; the caller
  // exchanges start_results() (16 B per start) and calls commit(gamma*) on every rank.
  void set_shard(int rank, int world) { m_rank = rank; m_world = world; }
  // Sharded fit with the exchange done here (lkgpu_comm.hpp: one process per GPU, TCP star around rank 0): with more
  // starts than processes the start indices come from a shared ticket counter (dynamic queue), otherwise process r
  // takes {s : s mod world == r}; after the starts every process holds ALL start_results(), applies the reference's
  // argmin and commits the same model.  `comm` must outlive the fit calls.
  void set_comm(ShardComm* comm);
  int local_n_eval() const { return m_local_n_eval; }           // evaluations this process ran in the last fit
  const std::vector<int>& local_starts() const { return m_local_starts; }  // start indices this process ran
  // Multistart rows in flight on this process's GPU (0 = by size: 8 for n <= 3072, 4 for n <= 8192, else 1).  One
  // engine handle and one host thread per row in flight; the L-BFGS-B code itself (not thread-safe: f2c statics)
  // runs under a process-wide mutex that is released for the duration of every objective evaluation.
  void set_concurrent_starts(int k) { m_concurrent_starts = k; }
  int last_concurrency() const { return m_last_concurrency; }

  void fit(const arma::vec& y, const arma::mat& X, const std::string& regmodel = "constant", bool normalize = false,
           const std::string& optim = "BFGS", const std::string& objective = "LL", const Parameters& parameters = Parameters());
  void fit(const arma::vec& y, const arma::vec& noise, const arma::mat& X, const std::string& regmodel = "constant",
           bool normalize = false, const std::string& optim = "BFGS", const std::string& objective = "LL",
           const Parameters& parameters = Parameters());

  // Kriging::update (src/lib/Kriging.cpp:2425-2660): block extension of the committed factor (refit = false), warm
  // restart (refit = true), a new fit for the Nugget refit and for the Heterogeneous overload.
  void update(const arma::vec& y_u, const arma::mat& X_u, bool refit = true);
  void update(const arma::vec& y_u, const arma::vec& noise_u, const arma::mat& X_u, bool refit = true);
  bool last_update_used_block_extension() const { return m_used_block; }

  std::tuple<double, arma::vec> logLikelihoodFun(const arma::vec& theta, bool return_grad);
  std::tuple<double, arma::vec> leaveOneOutFun(const arma::vec& theta, bool return_grad);
  std::tuple<double, arma::vec> logMargPostFun(const arma::vec& theta, bool return_grad);
  double logLikelihood() { return std::get<0>(logLikelihoodFun(m_theta, false)); }
  double leaveOneOut() { return std::get<0>(leaveOneOutFun(m_theta, false)); }
  double logMargPost() { return std::get<0>(logMargPostFun(m_theta, false)); }
  std::tuple<arma::vec, arma::vec> predict(const arma::mat& X_n, bool return_stdev);

  const std::string& kernel() const { return m_kernel; }
  const arma::vec& theta() const { return m_theta; }
  double sigma2() const { return m_sigma2; }
  double nugget() const { return m_nugget; }
  const arma::vec& beta() const { return m_beta; }
  arma::mat T();
  arma::mat M();
  arma::vec z();
  arma::mat circ();
  const std::vector<StartResult>& start_results() const { return m_results; }
  int n_eval() const { return m_n_eval; }
  void commit(const arma::vec& best_gamma);  // Kriging.cpp:2156-2202

 private:
  void fit_impl(const arma::vec& y, const arma::vec* noise, const arma::mat& X, const std::string& regmodel,
                bool normalize, const std::string& optim, const std::string& objective, const Parameters& prm);
  double objective(int obj, const arma::vec& gamma, arma::vec* grad);
  double objective_on(void* h, int obj, const arma::vec& gamma, arma::vec* grad) const;
  int concurrency(int n_starts, arma::uword n) const;
  void model_scalars(const arma::vec& theta, double extra, double* SSE, arma::vec* betahat);
  void push_params();
  void need_model();
  arma::vec gamma_full(const arma::vec& theta) const;
  arma::vec reparam_to(const arma::vec& v) const;
  arma::vec reparam_from(const arma::vec& g) const;
  arma::vec reparam_deriv(const arma::vec& v, const arma::vec& grad) const;
  double sigma2_variogram() const;
  void close();

  std::string m_kernel, m_objective = "LL", m_regmodel = "constant";
  NoiseModel m_noise_model;
  int m_device, m_rank = 0, m_world = 1, m_concurrent_starts = 0, m_last_concurrency = 1;
  ShardComm* m_comm = nullptr;
  int m_local_n_eval = 0;
  std::vector<int> m_local_starts;
  void* m_h = nullptr;
  bool m_is_empty = true, m_normalize = false;
  bool m_est_beta = true, m_est_sigma2 = true, m_est_nugget = true, m_est_theta = true, m_used_block = false;
  std::string m_optim = "BFGS";
  arma::mat m_X, m_F;
  arma::vec m_y, m_noise, m_theta, m_beta;
  arma::rowvec m_centerX, m_scaleX;
  double m_centerY = 0.0, m_scaleY = 1.0;
  double m_sigma2 = 1.0, m_nugget = 0.0, m_alpha = 1.0, m_commit_extra = 1.0;
  bool m_have_scalars = false;
  arma::vec m_scalars_theta, m_scalars_beta;
  double m_scalars_extra = 0.0, m_scalars_SSE = 0.0;
  std::vector<StartResult> m_results;
  int m_n_eval = 0;
};

arma::mat regression_model_matrix(const std::string& regmodel, const arma::mat& X);  // Trend.cpp:34-93

}  // namespace lkgpu
