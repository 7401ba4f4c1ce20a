// objective.inl -- host-side assembly of the reference's objective value and gradient from the
// device reductions of one lkgpu_eval.  Included in engine.cu's anonymous namespace.
// Mirrors, formula by formula:
//   Kriging::_logLikelihood  (src/lib/Kriging.cpp:214-341)
//   Kriging::_leaveOneOut    (src/lib/Kriging.cpp:353-468)
//   Kriging::_logMargPost    (src/lib/Kriging.cpp:488-648)
// gamma = [theta] (None) | [theta, alpha] (Nugget) | [theta, sigma2] (Heterogeneous).
// grad_out (may be null) has gamma_n entries.

static double objective_fun(Engine& e, int objective, const double* gamma, int gamma_n, double* grad_out,
                            lkgpu_out* user_out) {
  const int d = e.d, n = e.n, p = e.p;
  if (gamma_n != d && gamma_n != d + 1) throw LkError{"lkgpu_objective_fun: gamma must have d or d+1 entries"};
  if (gamma_n == d + 1 && e.noise_model == LKGPU_NOISE_NONE)
    throw LkError{"lkgpu_objective_fun: NoiseModel::None takes exactly d parameters"};
  const bool want_grad = grad_out != nullptr;
  std::vector<double> t1(d), t2(d), og(d), beta(p);
  lkgpu_out local;
  lkgpu_out* out = user_out ? user_out : &local;
  memset(out, 0, sizeof(*out));
  out->t1 = t1.data();
  out->t2 = t2.data();
  out->obj_grad = og.data();
  out->betahat = beta.data();

  // extra parameter selection (Kriging.cpp:221-234)
  double extra;
  if (gamma_n > d) extra = gamma[d];
  else extra = (e.noise_model == LKGPU_NOISE_NUGGET) ? e.alpha0 : e.sigma2;
  if (objective == LKGPU_OBJ_LL) {
    if (e.noise_model == LKGPU_NOISE_HETERO && !e.est_sigma2) extra = e.sigma2;
    else if (e.noise_model == LKGPU_NOISE_NUGGET && !e.est_sigma2 && !e.est_nugget)
      extra = e.sigma2 / (e.sigma2 + e.nugget);
  }
  if (grad_out)
    for (int k = 0; k < gamma_n; ++k) grad_out[k] = 0.0;

  double value = 0.0;
  const double PI = 3.14159265358979323846;
  if (objective == LKGPU_OBJ_LL) {
    e.eval(LKGPU_OBJ_LL, gamma, extra, want_grad ? 1 : 0, out);
    const double sumlog = out->sum_log_diagL, SSE = out->SSEstar;
    double s2g, ll;
    if (e.noise_model == LKGPU_NOISE_NUGGET) {
      double a = extra, s2 = e.sigma2, nug = e.nugget;
      if (e.est_sigma2) {
        if (e.est_nugget) {
          const double var = SSE / n;
          s2 = a * var;
          nug = (1.0 - a) * var;
        } else {
          s2 = e.nugget * a / (1.0 - a);
        }
      } else if (e.est_nugget) {
        nug = e.sigma2 * (1.0 - a) / a;
      }
      const double tv = s2 + nug;
      ll = -0.5 * (n * std::log(2 * PI * tv) + 2 * sumlog + SSE / tv);
      s2g = tv;
    } else if (e.noise_model == LKGPU_NOISE_HETERO) {
      const double s2 = e.est_sigma2 ? extra : e.sigma2;
      ll = -0.5 * (n * std::log(2 * PI * s2) + 2 * sumlog + SSE / s2);
      s2g = s2;
    } else if (e.est_sigma2) {
      s2g = SSE / n;
      ll = -0.5 * (n * std::log(2 * PI * s2g) + 2 * sumlog + n);
    } else {
      s2g = e.sigma2;
      ll = -0.5 * (n * std::log(2 * PI * s2g) + 2 * sumlog + SSE / s2g);
    }
    value = ll;
    if (want_grad) {
      for (int k = 0; k < d; ++k) grad_out[k] = (t1[k] / s2g + t2[k]) / 2.0;
      if (gamma_n > d) {
        if (e.noise_model == LKGPU_NOISE_NUGGET) {
          const double a = extra;
          if (e.est_sigma2 && e.est_nugget) {
            // dRdv = R / alpha, diag 0
            const double term1 = -(out->sum_offdiag_xRx / a) / s2g;
            const double term2 = out->sum_offdiag_RinvR / a;
            grad_out[d] = -0.5 * (term1 + term2);
          } else if (e.est_sigma2 && !e.est_nugget) {
            // dRdv = R / alpha, diag 1
            const double xRx = out->sum_offdiag_xRx / a + out->sum_x2;
            const double tr = out->sum_offdiag_RinvR / a + out->trace_Rinv;
            const double term1 = -xRx / (s2g * s2g);
            const double term2 = tr / s2g;
            grad_out[d] = -0.5 * (term1 + term2) * e.nugget / (1.0 - a) / (1.0 - a);
          } else {
            grad_out[d] = 0.0;
          }
        } else if (e.noise_model == LKGPU_NOISE_HETERO) {
          if (!e.est_sigma2) {
            grad_out[d] = 0.0;
          } else {
            const double s2 = extra, s2sq = s2 * s2;
            grad_out[d] = -0.5 * (n / s2 - out->sum_noise_Rinv / s2sq + out->sum_noise_x2 / (s2sq * s2) - SSE / s2sq);
          }
        }
      }
    }
  } else if (objective == LKGPU_OBJ_LOO) {
    e.eval(LKGPU_OBJ_LOO, gamma, extra, want_grad ? 1 : 0, out);
    value = out->loo;
    if (want_grad)
      for (int k = 0; k < d; ++k) grad_out[k] = og[k];
  } else if (objective == LKGPU_OBJ_LMP) {
    const double alpha = (e.noise_model == LKGPU_NOISE_NUGGET) ? (gamma_n > d ? gamma[d] : e.alpha0) : 1.0;
    const bool analytic = want_grad && e.est_sigma2;
    e.eval(LKGPU_OBJ_LMP, gamma, alpha, analytic ? 1 : 0, out);
    double sigma2;
    if (e.noise_model == LKGPU_NOISE_NUGGET) {
      if (e.est_sigma2 && e.est_nugget) sigma2 = out->S2 / (n - p);
      else if (e.est_sigma2 || e.est_nugget) sigma2 = e.sigma2 / alpha;
      else sigma2 = e.sigma2 + e.nugget;
    } else if (e.est_sigma2) {
      sigma2 = out->S2 / (n - p);
    } else {
      sigma2 = e.sigma2;
    }
    const double logS2 = std::log(sigma2 * (n - p));
    const double lml = -out->sum_log_diagL - out->sum_log_diagLX - (n - p) / 2.0 * logS2;
    const double a = 0.2;
    const double nroot = std::pow((double)n, 1.0 / d);
    const double b = 1.0 / nroot * (a + d);
    double t = 0.0;
    std::vector<double> CL(d);
    for (int k = 0; k < d; ++k) {
      CL[k] = (e.xmax[k] - e.xmin[k]) / nroot;
      t += CL[k] / gamma[k];
    }
    if (e.noise_model == LKGPU_NOISE_NUGGET) t += (1.0 - alpha) / alpha;
    const double lprior = -b * t + a * std::log(t);
    value = lml + lprior;
    if (want_grad) {
      if (e.est_sigma2) {
        for (int k = 0; k < d; ++k) {
          const double ans = (t1[k] / sigma2 + t2[k]) / 2.0;
          grad_out[k] = ans - (a * CL[k] / t - b * CL[k]) / (gamma[k] * gamma[k]);
        }
        if (e.noise_model == LKGPU_NOISE_NUGGET && gamma_n > d) {
          if (e.est_sigma2 || e.est_nugget) {
            // gradR_d = R / alpha, diag 0
            const double ans_d = -0.5 * (out->sum_offdiag_RinvR / alpha) + (out->sum_offdiag_xRx / alpha) / (2.0 * sigma2);
            grad_out[d] = ans_d - (a / t - b) / (alpha * alpha);
          } else {
            grad_out[d] = 0.0;
          }
        }
      } else {
        // fixed sigma2: forward differences, as the reference does (Kriging.cpp:633-644)
        const double eps = 1e-6;
        std::vector<double> ge(gamma, gamma + gamma_n);
        for (int k = 0; k < d; ++k) {
          ge[k] = gamma[k] + eps;
          const double v = objective_fun(e, LKGPU_OBJ_LMP, ge.data(), gamma_n, nullptr, nullptr);
          grad_out[k] = (v - value) / eps;
          ge[k] = gamma[k];
        }
        if (e.noise_model == LKGPU_NOISE_NUGGET && gamma_n > d) grad_out[d] = 0.0;
        // leave the handle's model at gamma (the finite-difference probes moved it)
        lkgpu_out scratch;
        memset(&scratch, 0, sizeof(scratch));
        e.eval(LKGPU_OBJ_LMP, gamma, alpha, 0, &scratch);
      }
    }
  } else {
    throw LkError{"lkgpu_objective_fun: unknown objective"};
  }
  if (user_out) {
    // the scratch arrays die with this frame
    user_out->t1 = user_out->t2 = user_out->obj_grad = user_out->betahat = nullptr;
  }
  return value;
}
