// lkgpu_host_driver.cpp -- command-line front end of the C++ host (lkgpu::Kriging), same work-directory protocol
// and JSON output as oracle/ref_driver.cpp so that the tests can put the two side by side:
//   <workdir>/cfg.txt (key=value), X.bin (n*d column-major), y.bin, noise.bin, theta.bin (nt*d), gamma.bin, Xn.bin (m*d)
// Sharded fit: launched once per GPU with RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment
// (LOCAL_RANK or LKGPU_HOST_DEVICE picks the device) the processes share the multistart rows over lkgpu::ShardComm;
// every process prints its own JSON line (same model on all of them, plus its rank and the starts it ran).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "../../include/lkgpu.h"
#include "lkgpu_comm.hpp"
#include "lkgpu_kriging.hpp"

static std::vector<double> read_bin(const std::string& path, size_t count) {
  std::vector<double> v(count);
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
  const size_t got = fread(v.data(), sizeof(double), count, f);
  fclose(f);
  if (got != count) { fprintf(stderr, "short read %s\n", path.c_str()); exit(2); }
  return v;
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
  if (argc < 2) { fprintf(stderr, "usage: lkgpu_host_driver <workdir>\n"); return 2; }
  const std::string wd = argv[1];
  std::map<std::string, std::string> cfg;
  {
    std::ifstream f(wd + "/cfg.txt");
    std::string line;
    while (std::getline(f, line)) {
      auto eq = line.find('=');
      if (eq != std::string::npos) cfg[line.substr(0, eq)] = line.substr(eq + 1);
    }
  }
  auto gets = [&](const char* k, const char* def) { return cfg.count(k) ? cfg[k] : std::string(def); };
  auto geti = [&](const char* k, int def) { return cfg.count(k) ? atoi(cfg[k].c_str()) : def; };
  auto getd = [&](const char* k, double def) { return cfg.count(k) ? atof(cfg[k].c_str()) : def; };
  const int n = geti("n", 0), d = geti("d", 0);
  const std::string mode = gets("mode", "eval"), kernel = gets("kernel", "gauss"), noise_model = gets("noise_model", "none");
  const std::string objective = gets("objective", "LL"), regmodel = gets("regmodel", "constant"), optim = gets("optim", "none");
  const bool normalize = geti("normalize", 0) != 0;
  const int nt = geti("ntheta", 1), want_grad = geti("grad", 1);
  int device = geti("device", 0);
  try {
    std::unique_ptr<lkgpu::ShardComm> comm = lkgpu::ShardComm::from_env();
    if (comm) {
      if (const char* dv = getenv("LKGPU_HOST_DEVICE")) device = atoi(dv);
      else if (const char* lr = getenv("LOCAL_RANK")) device = atoi(lr);
    }
    arma::mat X(read_bin(wd + "/X.bin", (size_t)n * d).data(), n, d);
    arma::vec y(read_bin(wd + "/y.bin", n).data(), n);
    arma::vec noise;
    if (noise_model == "hetero") noise = arma::vec(read_bin(wd + "/noise.bin", n).data(), n);
    using NM = lkgpu::Kriging::NoiseModel;
    const NM nm = noise_model == "nugget" ? NM::Nugget : noise_model == "hetero" ? NM::Heterogeneous : NM::None;
    lkgpu::Kriging k(kernel, nm, device);
    if (cfg.count("concurrent_starts")) k.set_concurrent_starts(geti("concurrent_starts", 0));
    if (comm) k.set_comm(comm.get());
    lkgpu::Kriging::Parameters prm;
    if (geti("beta_n", 0) > 0) {  // fixed trend coefficients (Parameters::beta, is_beta_estim = false)
      const int bn = geti("beta_n", 0);
      prm.beta = arma::vec(read_bin(wd + "/beta.bin", bn).data(), bn);
      prm.is_beta_estim = false;
    }
    if (exists(wd + "/theta.bin")) prm.theta = arma::mat(read_bin(wd + "/theta.bin", (size_t)nt * d).data(), nt, d);
    if (cfg.count("sigma2")) { prm.sigma2 = getd("sigma2", 1.0); prm.is_sigma2_estim = geti("est_sigma2", 0) != 0; }
    if (cfg.count("nugget")) { prm.nugget = getd("nugget", 0.0); prm.is_nugget_estim = geti("est_nugget", 0) != 0; }
    std::ostringstream js;
    js << "{";
    // CUDA start-up of this process (driver initialisation + context on `device`) is not part of the fit: timed apart
    {
      const double tc = now_s();
      unsigned long long free_b = 0, total_b = 0;
      lkgpu_mem_info(device, &free_b, &total_b);
      js << "\"cuda_init_s\": " << (now_s() - tc) << ", ";
    }
    if (comm) comm->barrier();  // sharded fit: the processes start their fits together
    const double t0 = now_s();
    if (nm == NM::Heterogeneous) k.fit(y, noise, X, regmodel, normalize, mode == "fit" ? optim : "none", objective, prm);
    else k.fit(y, X, regmodel, normalize, mode == "fit" ? optim : "none", objective, prm);
    js << "\"fit_s\": " << (now_s() - t0) << ", \"n_eval\": " << k.n_eval() << ", \"concurrent_starts\": "
       << k.last_concurrency() << ", ";
    if (comm) {
      js << "\"rank\": " << comm->rank() << ", \"world\": " << comm->world() << ", \"device\": " << device
         << ", \"local_n_eval\": " << k.local_n_eval() << ", \"local_starts\": [";
      for (size_t i = 0; i < k.local_starts().size(); ++i) js << (i ? ", " : "") << k.local_starts()[i];
      js << "], ";
    }
    if (mode == "eval") {
      const int gd = d + (nm == NM::None ? 0 : 1);
      arma::vec gamma = exists(wd + "/gamma.bin") ? arma::vec(read_bin(wd + "/gamma.bin", gd).data(), gd)
                                                   : arma::vec(prm.theta.value().row(0).t());
      std::tuple<double, arma::vec> res;
      if (objective == "LL") res = k.logLikelihoodFun(gamma, want_grad != 0);
      else if (objective == "LOO") res = k.leaveOneOutFun(gamma, want_grad != 0);
      else res = k.logMargPostFun(gamma, want_grad != 0);
      js.precision(17);
      js << "\"value\": " << std::scientific << std::get<0>(res) << ", ";
      jvec(js, "grad", std::get<1>(res)); js << ", ";
    }
    // Kriging::update with n_u further observations (same protocol as oracle/ref_driver.cpp)
    const int n_u = geti("update_n", 0);
    if (n_u > 0) {
      arma::mat Xu(read_bin(wd + "/Xu.bin", (size_t)n_u * d).data(), n_u, d);
      arma::vec yu(read_bin(wd + "/yu.bin", n_u).data(), n_u);
      const bool refit = geti("update_refit", 0) != 0;
      const double t1 = now_s();
      if (nm == NM::Heterogeneous) {
        arma::vec nu(read_bin(wd + "/noiseu.bin", n_u).data(), n_u);
        k.update(yu, nu, Xu, refit);
      } else {
        k.update(yu, Xu, refit);
      }
      js << "\"update_s\": " << (now_s() - t1) << ", \"used_block_extension\": "
         << (k.last_update_used_block_extension() ? 1 : 0) << ", ";
    }
    js.precision(17);
    jvec(js, "theta", k.theta()); js << ", ";
    jvec(js, "beta", k.beta()); js << ", ";
    js << "\"sigma2\": " << std::scientific << k.sigma2() << ", \"nugget\": " << k.nugget() << ", ";
    if (mode == "fit") {
      const double ll = objective == "LOO" ? k.leaveOneOut() : (objective == "LMP" ? k.logMargPost() : k.logLikelihood());
      js << "\"objective_at_fit\": " << ll << ", ";
    }
    if (mode == "fit" || n_u > 0) js << "\"LL_at_model\": " << std::scientific << k.logLikelihood() << ", ";
    if (exists(wd + "/Xn.bin")) {
      const int m = geti("m", 0);
      arma::mat Xn(read_bin(wd + "/Xn.bin", (size_t)m * d).data(), m, d);
      auto pr = k.predict(Xn, true);
      jvec(js, "pred_mean", std::get<0>(pr)); js << ", ";
      jvec(js, "pred_sd", std::get<1>(pr)); js << ", ";
    }
    js << "\"n\": " << n << ", \"d\": " << d << "}";
    std::cout << js.str() << std::endl;
    if (comm) comm->barrier();  // leave together (rank 0 serves the others until they are done)
  } catch (const std::exception& e) {
    std::cout << "{\"error\": \"" << e.what() << "\"}" << std::endl;
    return 1;
  }
  return 0;
}
