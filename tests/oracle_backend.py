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

    def set_fixed_beta(self, beta):
        self.fixed_beta = None if beta is None else np.asarray(beta, float).copy()

    def theta_bounds(self, lower_factor, upper_factor, heuristic):
        return ko.theta_bounds(self.pb.X, self.pb.y, lower_factor, upper_factor, heuristic)

    def sigma2_variogram(self):
        return ko.sigma2_variogram(self.pb.X, self.pb.y)

    def objective(self, name, gamma, want_grad):
        self.n_objective_calls += 1
        try:
            if name == "LL":
                return ko.log_likelihood(self.pb, gamma, want_grad)
            if name == "LOO":
                return ko.leave_one_out(self.pb, gamma, want_grad)
            return ko.log_marg_post(self.pb, gamma, want_grad)
        finally:
            self.pb.kept = None  # like the device engine: the kept factor is consumed by the first evaluation

    def model_scalars(self, theta, extra):
        self.theta = np.asarray(theta, float)
        self.extra = extra if self.pb.noise_model != "none" else None
        self.model = ko.populate_model(self.pb, theta, self.extra)
        self.used_block_update = self.model.used_block_update
        self.pb.kept = None  # like the device engine: the kept factor is consumed by the first evaluation
        return self.model.SSEstar, self.model.betahat

    def commit(self):
        self.committed = (self.model, self.theta, getattr(self, "extra", None))

    def restore(self):
        self.model, self.theta, self.extra = self.committed

    def append_data(self, X_u, y_u, F_u, noise_u=None):
        """Kriging::update, data side: the committed model (self.model at self.theta) becomes the kept factor."""
        pb = self.pb
        if self.model is not None:
            pb.kept = ko.KeptModel(T=self.model.L, R=self.model.R, theta=np.asarray(self.theta, float).copy(),
                                   extra=getattr(self, "extra", None))
        pb.X = np.vstack([pb.X, X_u])
        pb.y = np.concatenate([pb.y, y_u])
        pb.F = np.vstack([pb.F, F_u])
        if pb.noise is not None and noise_u is not None:
            pb.noise = np.concatenate([pb.noise, noise_u])
        self.model = None

    def export(self, which):
        m = self.model
        fb = getattr(self, "fixed_beta", None)
        z = m.Estar if fb is None else m.ystar - m.Fstar @ fb
        return {"L": m.L, "Fstar": m.Fstar, "Estar": m.Estar, "Rstar": m.Rstar, "z": z}[which]

    def predict(self, Xn, Fn, beta, r_on_factor):
        pb = self.pb
        saved = pb.alpha
        pb.alpha = r_on_factor
        mean, sd = ko.predict(pb, self.theta, 1.0, Xn, Fn, self.model, fixed_beta=getattr(self, "fixed_beta", None))
        pb.alpha = saved
        return mean, sd * sd

    def close(self):
        pass
