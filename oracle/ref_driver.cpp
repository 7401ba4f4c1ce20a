// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Thin command-line driver around the UNMODIFIED reference implementation
// (libKriging, compiled from the sources where they lie under /root/reference
// by oracle/build_ref.sh into oracle/_ref/).  It only calls the reference's
// public API (src/lib/include/libKriging/Kriging.hpp:113-274):
//   Kriging::fit, logLikelihoodFun, leaveOneOutFun, logMargPostFun, predict.
// It is used (a) to pin oracle/kriging_oracle.py, (b) to generate the fixtures
// in tests/golden/, (c) as the CPU baseline of bench.py (kind "reference").
//
// With update_n=<n_u> (+ Xu.bin, yu.bin, noiseu.bin; update_refit=0|1) the fitted model is then extended by
// Kriging::update before the outputs are written.
// Usage: ref_driver <workdir>
//   <workdir>/cfg.txt      key=value lines (see parse below)
//   <workdir>/X.bin        n*d float64, column-major
//   <workdir>/y.bin        n float64
//   <workdir>/noise.bin    n float64 (heterogeneous only)
//   <workdir>/theta.bin    nt*d float64 column-major (eval point / start points)
//   <workdir>/Xn.bin       m*d float64 (optional, predict)
// Output: JSON on stdout; matrices (if dump=1) as <workdir>/out_*.bin.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "libKriging/Kriging.hpp"
#include "libKriging/LinearAlgebra.hpp"
#include "libKriging/Optim.hpp"
#include "libKriging/Trend.hpp"

static std::vector<double> read_bin(const std::string& path, size_t count) {
  std::vector<double> v(count);
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
  size_t got = fread(v.data(), sizeof(double), count, f);
  fclose(f);
  if (got != count) { fprintf(stderr, "short read %s (%zu of %zu)\n", path.c_str(), got, count); exit(2); }
  return v;
}
static void write_bin(const std::string& path, const double* p, size_t count) {
  FILE* f = fopen(path.c_str(), "wb");
  fwrite(p, sizeof(double), count, f);
  fclose(f);
}
static bool exists(const std::string& p) { std::ifstream f(p); return f.good(); }
static void jvec(std::ostream& os, const char* key, const arma::vec& v) {
  os << "\"" << key << "\": [";
  os.precision(17);
  for (arma::uword i = 0; i < v.n_elem; i++) os << (i ? ", " : "") << std::scientific << v[i];
  os << "]";
}
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: ref_driver <workdir>\n"); return 2; }
  std::string wd = argv[1];
  std::map<std::string, std::string> cfg;
  {
    std::ifstream f(wd + "/cfg.txt");
    std::string line;
    while (std::getline(f, line)) {
      auto eq = line.find('=');
      if (eq == std::string::npos) continue;
      cfg[line.substr(0, eq)] = line.substr(eq + 1);
    }
  }
  auto gets = [&](const char* k, const char* def) { return cfg.count(k) ? cfg[k] : std::string(def); };
  auto geti = [&](const char* k, int def) { return cfg.count(k) ? atoi(cfg[k].c_str()) : def; };
  auto getd = [&](const char* k, double def) { return cfg.count(k) ? atof(cfg[k].c_str()) : def; };

  const int n = geti("n", 0), d = geti("d", 0);
  const std::string mode = gets("mode", "eval");            // eval | fit
  const std::string kernel = gets("kernel", "gauss");
  const std::string noise_model = gets("noise_model", "none");
  const std::string objective = gets("objective", "LL");
  const std::string regmodel = gets("regmodel", "constant");
  const std::string optim = gets("optim", "none");
  const bool normalize = geti("normalize", 0) != 0;
  const int nt = geti("ntheta", 1);
  const int reps = geti("reps", 1);
  const int want_grad = geti("grad", 1);
  const int dump = geti("dump", 0);
  const int rcond_check = geti("rcond_check", 1);
  LinearAlgebra::check_chol_rcond(rcond_check != 0);
  if (cfg.count("num_nugget")) LinearAlgebra::set_num_nugget(getd("num_nugget", 1e-10));

  arma::mat X(read_bin(wd + "/X.bin", (size_t)n * d).data(), n, d);
  arma::vec y(read_bin(wd + "/y.bin", n).data(), n);
  arma::vec noise;
  if (noise_model == "hetero") noise = arma::vec(read_bin(wd + "/noise.bin", n).data(), n);

  Kriging::NoiseModel nm = noise_model == "nugget" ? Kriging::NoiseModel::Nugget
                           : noise_model == "hetero" ? Kriging::NoiseModel::Heterogeneous
                                                     : Kriging::NoiseModel::None;
  Kriging k(kernel, nm);
  Kriging::Parameters prm;
  if (exists(wd + "/theta.bin")) prm.theta = arma::mat(read_bin(wd + "/theta.bin", (size_t)nt * d).data(), nt, d);
  if (cfg.count("sigma2")) { prm.sigma2 = getd("sigma2", 1.0); prm.is_sigma2_estim = geti("est_sigma2", 0) != 0; }
  if (cfg.count("nugget")) { prm.nugget = getd("nugget", 0.0); prm.is_nugget_estim = geti("est_nugget", 0) != 0; }
  // fixed trend coefficients (Parameters::beta with is_beta_estim = false): <workdir>/beta.bin, beta_n entries
  if (geti("beta_n", 0) > 0) {
    const int bn = geti("beta_n", 0);
    prm.beta = arma::vec(read_bin(wd + "/beta.bin", bn).data(), bn);
    prm.is_beta_estim = false;
  }
  Trend::RegressionModel rm = Trend::fromString(regmodel);

  std::ostringstream js;
  js << "{";
  double t0 = now_s();
  if (nm == Kriging::NoiseModel::Heterogeneous)
    k.fit(y, noise, X, rm, normalize, mode == "fit" ? optim : "none", objective, prm);
  else
    k.fit(y, X, rm, normalize, mode == "fit" ? optim : "none", objective, prm);
  double t_fit = now_s() - t0;
  js << "\"fit_s\": " << t_fit << ", ";

  if (mode == "eval") {
    // evaluation point: gamma = [theta, extra] (extra = alpha | sigma2) from gamma.bin, else theta row 0
    arma::vec gamma;
    int gd = d + (nm == Kriging::NoiseModel::None ? 0 : 1);
    if (exists(wd + "/gamma.bin")) gamma = arma::vec(read_bin(wd + "/gamma.bin", gd).data(), gd);
    else gamma = prm.theta.value().row(0).t();
    std::vector<double> times;
    double val = 0; arma::vec grad;
    for (int r = 0; r < reps; r++) {
      double t1 = now_s();
      std::tuple<double, arma::vec> res;
      if (objective == "LL") res = k.logLikelihoodFun(gamma, want_grad != 0, false);
      else if (objective == "LOO") res = k.leaveOneOutFun(gamma, want_grad != 0, false);
      else res = k.logMargPostFun(gamma, want_grad != 0, false);
      times.push_back(now_s() - t1);
      val = std::get<0>(res); grad = std::get<1>(res);
    }
    js << "\"eval_s_all\": [";
    for (size_t i = 0; i < times.size(); i++) js << (i ? ", " : "") << times[i];
    js << "], ";
    std::sort(times.begin(), times.end());
    js.precision(17);
    js << "\"value\": " << std::scientific << val << ", ";
    jvec(js, "grad", grad); js << ", ";
    js << "\"eval_s_median\": " << times[times.size() / 2] << ", \"eval_s_min\": " << times[0] << ", ";
    if (objective == "LOO" && geti("loovec", 0)) {
      auto lv = k.leaveOneOutVec(gamma);
      jvec(js, "loo_mean", std::get<0>(lv)); js << ", ";
      jvec(js, "loo_sd", std::get<1>(lv)); js << ", ";
    }
  }
  // Kriging::update (src/lib/Kriging.cpp:2425-2660) with n_u further observations
  const int n_u = geti("update_n", 0);
  if (n_u > 0) {
    arma::mat Xu(read_bin(wd + "/Xu.bin", (size_t)n_u * d).data(), n_u, d);
    arma::vec yu(read_bin(wd + "/yu.bin", n_u).data(), n_u);
    const bool refit = geti("update_refit", 0) != 0;
    double t1 = now_s();
    if (nm == Kriging::NoiseModel::Heterogeneous) {
      arma::vec nu(read_bin(wd + "/noiseu.bin", n_u).data(), n_u);
      k.update(yu, nu, Xu, refit);
    } else {
      k.update(yu, Xu, refit);
    }
    js << "\"update_s\": " << (now_s() - t1) << ", ";
  }
  js.precision(17);
  jvec(js, "theta", k.theta()); js << ", ";
  jvec(js, "beta", k.beta()); js << ", ";
  js << "\"sigma2\": " << std::scientific << k.sigma2() << ", \"nugget\": " << k.nugget() << ", ";
  if (mode == "fit") {
    double ll = (objective == "LOO") ? k.leaveOneOut() : (objective == "LMP" ? k.logMargPost() : k.logLikelihood());
    js << "\"objective_at_fit\": " << ll << ", ";
  }
  if (exists(wd + "/Xn.bin")) {
    int m = geti("m", 0);
    arma::mat Xn(read_bin(wd + "/Xn.bin", (size_t)m * d).data(), m, d);
    auto pr = k.predict(Xn, true, false, false);
    jvec(js, "pred_mean", std::get<0>(pr)); js << ", ";
    jvec(js, "pred_sd", std::get<1>(pr)); js << ", ";
  }
  if (mode == "fit" || n_u > 0) {
    js << "\"LL_at_model\": " << std::scientific << k.logLikelihood() << ", ";
  }
  if (dump) {
    write_bin(wd + "/out_T.bin", k.T().memptr(), k.T().n_elem);
    write_bin(wd + "/out_M.bin", k.M().memptr(), k.M().n_elem);
    write_bin(wd + "/out_z.bin", k.z().memptr(), k.z().n_elem);
  }
  js << "\"n\": " << n << ", \"d\": " << d << "}";
  std::cout << js.str() << std::endl;
  return 0;
}
