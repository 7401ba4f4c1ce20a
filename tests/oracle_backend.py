"""TEST-ONLY objective backend: lets the CPU test-suite drive the product's host-side fit loop
(libkriging_b200/kriging.py: bounds, starts, reparametrisation, L-BFGS-B, restarts, argmin, commit) with
the oracle as the objective provider, so the host logic is checked against the reference's fits without a
GPU.  The product never constructs this class (tests/ may use oracle/, the package may not)."""
import numpy as np

from oracle import kriging_oracle as ko


class OracleBackend:
    def __init__(self, X, y, F, kernel, noise_model, noise, device=0):
        self.pb = ko.Problem(X=np.asarray(X), y=np.asarray(y), F=np.asarray(F), kernel=kernel,
                             noise_model=noise_model, noise=noise)
        self.model = None
        self.n_objective_calls = 0

    def set_params(self, est_sigma2, sigma2, est_nugget, nugget, alpha):
        pb = self.pb
        pb.est_sigma2, pb.sigma2, pb.est_nugget, pb.nugget, pb.alpha = est_sigma2, sigma2, est_nugget, nugget, alpha

    def theta_bounds(self, lower_factor, upper_factor, heuristic):
        return ko.theta_bounds(self.pb.X, self.pb.y, lower_factor, upper_factor, heuristic)

    def sigma2_variogram(self):
        return ko.sigma2_variogram(self.pb.X, self.pb.y)

    def objective(self, name, gamma, want_grad):
        self.n_objective_calls += 1
        if name == "LL":
            return ko.log_likelihood(self.pb, gamma, want_grad)
        if name == "LOO":
            return ko.leave_one_out(self.pb, gamma, want_grad)
        return ko.log_marg_post(self.pb, gamma, want_grad)

    def model_scalars(self, theta, extra):
        self.theta = np.asarray(theta, float)
        self.model = ko.populate_model(self.pb, theta, extra if self.pb.noise_model != "none" else None)
        return self.model.SSEstar, self.model.betahat

    def export(self, which):
        m = self.model
        return {"L": m.L, "Fstar": m.Fstar, "Estar": m.Estar, "Rstar": m.Rstar}[which]

    def predict(self, Xn, Fn, beta, r_on_factor):
        pb = self.pb
        saved = pb.alpha
        pb.alpha = r_on_factor
        mean, sd = ko.predict(pb, self.theta, 1.0, Xn, Fn, self.model)
        pb.alpha = saved
        return mean, sd * sd

    def close(self):
        pass
