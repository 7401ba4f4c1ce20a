// oracle/ref_nested_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Driver around the UNMODIFIED reference's NestedKriging (src/lib/NestedKriging.cpp, compiled where it lies by
// oracle/build_ref.sh): fits a NestedKriging with a Random partition and prints, as JSON, the partition it drew, the
// unified hyper-parameters (NestedKriging.cpp:277-331) and, per sub-model, theta / sigma2 / beta / log-likelihood
// after the closed-form re-fit.  Generates tests/golden/refgen_nested.json.
// Usage: ref_nested_driver <workdir>     cfg.txt: n, d, groups, kernel, seed ; X.bin (col-major), y.bin
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "libKriging/NestedKriging.hpp"

static std::vector<double> read_bin(const std::string& path, size_t count) {
  std::vector<double> v(count);
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
  size_t got = fread(v.data(), sizeof(double), count, f);
  fclose(f);
  if (got != count) { fprintf(stderr, "short read %s\n", path.c_str()); exit(2); }
  return v;
}
static void jvec(std::ostream& os, const char* key, const arma::vec& v) {
  os << "\"" << key << "\": [";
  os.precision(17);
  for (arma::uword i = 0; i < v.n_elem; i++) os << (i ? ", " : "") << std::scientific << v[i];
  os << "]";
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::string wd = argv[1];
  std::map<std::string, std::string> cfg;
  {
    std::ifstream f(wd + "/cfg.txt");
    std::string line;
    while (std::getline(f, line)) {
      auto eq = line.find('=');
      if (eq != std::string::npos) cfg[line.substr(0, eq)] = line.substr(eq + 1);
    }
  }
  const int n = atoi(cfg["n"].c_str()), d = atoi(cfg["d"].c_str()), p = atoi(cfg["groups"].c_str());
  const int seed = cfg.count("seed") ? atoi(cfg["seed"].c_str()) : 123;
  arma::mat X(read_bin(wd + "/X.bin", (size_t)n * d).data(), n, d);
  arma::vec y(read_bin(wd + "/y.bin", n).data(), n);
  Kriging::Parameters prm;
  if (cfg.count("theta0")) prm.theta = arma::mat(1, d, arma::fill::value(atof(cfg["theta0"].c_str())));
  NestedKriging nk(y, X, cfg["kernel"], p, NestedKriging::Aggregation::PoE, NestedKriging::Partition::Random, seed,
                   Trend::RegressionModel::Constant, cfg.count("optim") ? cfg["optim"] : "BFGS", "LL", prm);
  std::ostringstream js;
  js << "{";
  jvec(js, "theta", nk.theta());
  js.precision(17);
  js << ", \"sigma2\": " << std::scientific << nk.sigma2() << ", \"beta0\": " << nk.beta0() << ", \"groups\": [";
  for (arma::uword g = 0; g < nk.nb_groups(); ++g) {
    js << (g ? ", [" : "[");
    const arma::uvec& idx = nk.groups()[g];
    for (arma::uword i = 0; i < idx.n_elem; ++i) js << (i ? "," : "") << idx[i];
    js << "]";
  }
  js << "], \"submodels\": [";
  for (arma::uword g = 0; g < nk.nb_groups(); ++g) {
    const Kriging& k = nk.submodel(g);
    js << (g ? ", {" : "{");
    jvec(js, "theta", k.theta());
    js << ", ";
    jvec(js, "beta", k.beta());
    js << ", \"sigma2\": " << std::scientific << k.sigma2() << ", \"LL\": " << const_cast<Kriging&>(k).logLikelihood() << "}";
  }
  js << "]}";
  std::cout << js.str() << std::endl;
  return 0;
}
