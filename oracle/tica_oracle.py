"""oracle/tica_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

NumPy float64 restatement of the reference tICA estimator's arithmetic
(msmbuilder/decomposition/tica.py in /root/reference).  It is the parity oracle
and the "port" CPU baseline for the covariance half of the hot path; it travels
to the GPU box (the reference's own tica.py cannot: /root/reference is absent
there and ``import msmbuilder`` needs mdtraj).

Pinned in tests/test_oracle_tica.py against
  * the reference's OWN tica.py loaded verbatim (oracle/ref_loader.py) on seeded
    inputs, when /root/reference is present, and
  * tests/golden/tica_*.npz written by oracle/gen_golden.py from that verbatim
    reference run, everywhere else.

Line map (reference tica.py):
  state / _initialize ............ :113-165
  _fit ........................... :401-424  (f64 cast :402, short-sequence skip
                                              :410-412, counters :414-415,
                                              six accumulations :417-422)
  means_ / offset_correlation_ /
  covariance_ .................... :228-259  (two_N :229,:236,:245)
  rao_blackwell_ledoit_wolf ...... :492-524
  _solve ......................... :167-199  (eigh on the top-k index range;
                                              `eigvals=` became `subset_by_index=`
                                              in SciPy >= 1.14, same LAPACK call)
  transform ...................... :312-354
  score .......................... :426-467
"""
import warnings

import numpy as np
import scipy.linalg


def rao_blackwell_ledoit_wolf(S, n):
    """tica.py:492-524."""
    p = len(S)
    assert S.shape == (p, p)
    alpha = (n - 2) / (n * (n + 2))
    beta = ((p + 1) * n - 2) / (n * (n + 2))
    trace_S2 = np.sum(S * S)
    U = ((p * trace_S2 / np.trace(S) ** 2) - 1)
    rho = min(alpha + beta / U, 1)
    F = (np.trace(S) / p) * np.eye(p)
    return (1 - rho) * S + rho * F, rho


class TicaOracle(object):
    """Same public surface as the reference class, trimmed to the arithmetic."""

    def __init__(self, n_components=None, lag_time=1, shrinkage=None,
                 kinetic_mapping=False, commute_mapping=False):
        if kinetic_mapping and commute_mapping:
            raise ValueError("Can't have both kinetic mapping and commute mapping.")
        self.n_components = n_components
        self.lag_time = lag_time
        self.shrinkage = shrinkage
        self.shrinkage_ = None
        self.kinetic_mapping = kinetic_mapping
        self.commute_mapping = commute_mapping
        self.n_features = None
        self.n_observations_ = None
        self.n_sequences_ = None
        self._initialized = False
        self._is_dirty = True
        self._eigenvalues_ = None
        self._eigenvectors_ = None

    # -- accumulation ------------------------------------------------------
    def _initialize(self, n_features):
        if self._initialized:
            return
        if self.n_components is None:
            self.n_components = n_features
        self.n_features = n_features
        self.n_observations_ = 0
        self.n_sequences_ = 0
        z2 = lambda: np.zeros((n_features, n_features))
        z1 = lambda: np.zeros(n_features)
        self._outer_0_to_T_lagged = z2()
        self._sum_0_to_TminusTau = z1()
        self._sum_tau_to_T = z1()
        self._sum_0_to_T = z1()
        self._outer_0_to_TminusTau = z2()
        self._outer_offset_to_T = z2()
        self._initialized = True

    def partial_fit(self, X):
        X = np.asarray(np.atleast_2d(X), dtype=np.float64)
        tau = self.lag_time
        self._initialize(X.shape[1])
        if not len(X) > tau:
            warnings.warn("length of data (%d) is too short for the lag time (%d)"
                          % (len(X), tau))
            return self
        self.n_observations_ += X.shape[0]
        self.n_sequences_ += 1
        head, tail = X[:-tau], X[tau:]
        self._outer_0_to_T_lagged += np.dot(head.T, tail)
        self._sum_0_to_TminusTau += head.sum(axis=0)
        self._sum_tau_to_T += tail.sum(axis=0)
        self._sum_0_to_T += X.sum(axis=0)
        self._outer_0_to_TminusTau += np.dot(head.T, head)
        self._outer_offset_to_T += np.dot(tail.T, tail)
        self._is_dirty = True
        return self

    def fit(self, sequences):
        self._initialized = False
        for X in sequences:
            self.partial_fit(X)
        if self.n_sequences_ == 0:
            raise ValueError("All sequences were shorter than the lag time, %d"
                             % self.lag_time)
        return self

    # -- moments -----------------------------------------------------------
    def _two_N(self):
        return 2 * (self.n_observations_ - self.lag_time * self.n_sequences_)

    @property
    def means_(self):
        return (self._sum_0_to_TminusTau + self._sum_tau_to_T) / float(self._two_N())

    @property
    def offset_correlation_(self):
        term = (self._outer_0_to_T_lagged + self._outer_0_to_T_lagged.T) / self._two_N()
        mu = self.means_
        return term - np.outer(mu, mu)

    @property
    def covariance_(self):
        term = (self._outer_0_to_TminusTau + self._outer_offset_to_T) / self._two_N()
        mu = self.means_
        S = term - np.outer(mu, mu)
        if self.shrinkage is None:
            sigma, self.shrinkage_ = rao_blackwell_ledoit_wolf(S, n=self.n_observations_)
        else:
            self.shrinkage_ = self.shrinkage
            p = self.n_features
            F = (np.trace(S) / p) * np.eye(p)
            sigma = (1 - self.shrinkage) * S + self.shrinkage * F
        return sigma

    # -- eigensolve ----------------------------------------------------------
    def _solve(self):
        if not self._is_dirty and len(self._eigenvalues_) >= self.n_components:
            return
        if self.n_observations_ == 0:
            raise RuntimeError("The model must be fit() before use.")
        lhs = self.offset_correlation_
        rhs = self.covariance_
        if not np.allclose(lhs, lhs.T):
            raise RuntimeError("offset correlation matrix is not symmetric")
        if not np.allclose(rhs, rhs.T):
            raise RuntimeError("correlation matrix is not symmetric")
        lo, hi = self.n_features - self.n_components, self.n_features - 1
        vals, vecs = scipy.linalg.eigh(lhs, b=rhs, subset_by_index=(lo, hi))
        order = np.argsort(vals)[::-1]
        self._eigenvalues_ = vals[order]
        self._eigenvectors_ = vecs[:, order]
        self._is_dirty = False

    @property
    def eigenvalues_(self):
        self._solve()
        return self._eigenvalues_[:self.n_components]

    @property
    def eigenvectors_(self):
        self._solve()
        return self._eigenvectors_[:, :self.n_components]

    @property
    def components_(self):
        return self.eigenvectors_[:, 0:self.n_components].T

    @property
    def timescales_(self):
        self._solve()
        return -1. * self.lag_time / np.log(self._eigenvalues_[:self.n_components])

    @property
    def score_(self):
        self._solve()
        return self._eigenvalues_[:self.n_components].sum()

    # -- projection ----------------------------------------------------------
    def transform(self, sequences):
        out = []
        for X in sequences:
            X = np.asarray(np.atleast_2d(X))
            X = X - self.means_
            Y = np.dot(X, self.components_.T)
            if self.kinetic_mapping:
                Y *= self.eigenvalues_
            if self.commute_mapping:
                ts = self.timescales_
                reg = 0.5 * ts * np.tanh(np.pi * ((ts - self.lag_time) / self.lag_time) + 1)
                Y *= np.sqrt(reg / 2)
                Y = np.nan_to_num(Y)
            out.append(Y)
        return out

    def score(self, sequences):
        assert self._initialized
        V = self.eigenvectors_
        m2 = TicaOracle(shrinkage=self.shrinkage, n_components=self.n_components,
                        lag_time=self.lag_time)
        for X in sequences:
            m2.partial_fit(X)
        num = V.T.dot(m2.offset_correlation_).dot(V)
        den = V.T.dot(m2.covariance_).dot(V)
        try:
            return np.trace(num.dot(np.linalg.inv(den)))
        except np.linalg.LinAlgError:
            return np.nan

    # -- helpers for parity tests ---------------------------------------------
    def packed_moments(self):
        """[C_tau | C_00 | C_tautau | S_0 | S_tau | S | n_obs | n_seq] as one f64 vector
        (the layout of the device accumulator, include/msmb200.h)."""
        return np.concatenate([
            self._outer_0_to_T_lagged.ravel(), self._outer_0_to_TminusTau.ravel(),
            self._outer_offset_to_T.ravel(), self._sum_0_to_TminusTau,
            self._sum_tau_to_T, self._sum_0_to_T,
            [float(self.n_observations_), float(self.n_sequences_)]])
