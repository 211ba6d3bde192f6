"""tICA with the covariance accumulation on a B200.

Drop-in for ``msmbuilder.decomposition.tICA`` (msmbuilder/decomposition/tica.py:26):
same constructor, same public methods and fitted attributes, same *private*
accumulator names (subclasses and ``score`` in the reference depend on them,
tica.py:123-148,456-461).  The six streaming accumulations of ``_fit``
(tica.py:417-422) run on the GPU through the C ABI
``msmb200_tica_accumulate`` (include/msmb200.h); everything downstream of the
D x D sufficient statistics -- moments, shrinkage, the generalised symmetric
eigenproblem -- stays on the host in float64 exactly as in the reference
(tica.py:167-259,492-524).  State is plain NumPy, so estimators pickle and
``sklearn.clone`` like the originals.
"""
from __future__ import print_function, division, absolute_import

import ctypes
import warnings

import numpy as np
import scipy.linalg
from sklearn.base import TransformerMixin

from ..base import BaseEstimator
from ..utils import check_iter_of_sequences, is_tensor
from .. import _lib

__all__ = ['tICA']


def _eigh_top(lhs, rhs, lo, hi):
    """scipy.linalg.eigh on the index range [lo, hi] (tica.py:188-189).  The
    reference's ``eigvals=`` keyword was renamed ``subset_by_index`` in SciPy 1.5
    and removed in 1.14; both spell the same LAPACK call."""
    try:
        return scipy.linalg.eigh(lhs, b=rhs, subset_by_index=(lo, hi))
    except TypeError:  # pragma: no cover - very old SciPy
        return scipy.linalg.eigh(lhs, b=rhs, eigvals=(lo, hi))


class tICA(BaseEstimator, TransformerMixin):
    """Time-structure Independent Component Analysis (tICA), GPU-accumulated.

    Finds the linear combinations of the input features that decorrelate most
    slowly at the chosen lag time, by solving the generalised eigenproblem
    ``C v = lambda * Sigma v`` for the symmetrised time-lagged correlation
    matrix ``C`` and the covariance ``Sigma``.

    Parameters
    ----------
    n_components : int, None
        Number of slow components kept.  ``None`` keeps all of them.
    lag_time : int
        Delay, in frames, between the two time slices that are correlated.
    shrinkage : float, default=None
        Covariance shrinkage intensity in [0, 1].  ``None`` selects it
        analytically with the Rao-Blackwellised Ledoit-Wolf estimator.
    kinetic_mapping : bool, default=False
        Scale the projection by the eigenvalues (kinetic map).
    commute_mapping : bool, default=False
        Scale the projection by regularised timescales (commute map).
    engine : {'auto', 'simt_f64', 'umma_3xf16', 'umma_6xbf16', 'umma_3xbf16', 'umma_3xtf32', 'umma_tf32'}, default='auto'
        Which device kernel accumulates the covariance matrices.  'auto' uses the
        tcgen05 tensor-core kernel with the error-compensated fp16 split of the
        per-feature power-of-two scaled frames (~2^-22 per product) when the shape
        allows it (float32, n_features a multiple of 32 between 64 and 256) and the
        float64 CUDA-core kernel otherwise.  A value more than 2^6 times the largest
        magnitude of the scale sample (1024 frames spread over the call), or a
        non-finite one, makes the call redo itself on the stream with the float64
        kernel, so outliers cost time, not accuracy.  'umma_6xbf16' is ~2^-24 per
        product over the full float32 range, 'umma_3xbf16' ~2^-16, unbiased.
    devices : None, 'all' or list of CUDA ordinals, default=None
        GPUs ``fit`` deals whole sequences to from ONE process (a thread per GPU,
        the packed float64 accumulators added on the host): the whole box without
        torchrun.  None = the current device.

    Attributes
    ----------
    components_ : array-like, shape (n_components, n_features)
        Projection vectors, slowest first.
    offset_correlation_ : array-like, shape (n_features, n_features)
        Symmetrised time-lagged correlation matrix.
    eigenvalues_ : array-like, shape (n_components,)
        Autocorrelation of each component at the lag time.
    eigenvectors_ : array-like, shape (n_features, n_components)
    means_ : array, shape (n_features,)
    n_observations_ : int
        Frames seen so far (all ``fit`` / ``partial_fit`` calls).
    n_sequences_ : int
        Sequences seen so far.
    timescales_ : array-like, shape (n_components,)
        Implied timescales in units of frames.
    """

    def __init__(self, n_components=None, lag_time=1, shrinkage=None,
                 kinetic_mapping=False, commute_mapping=False, engine='auto', devices=None):
        self.n_components = n_components
        self.devices = devices
        self.lag_time = lag_time
        self.shrinkage = shrinkage
        self.shrinkage_ = None
        self.kinetic_mapping = kinetic_mapping
        self.commute_mapping = commute_mapping
        self.engine = engine
        if self.kinetic_mapping and self.commute_mapping:
            raise ValueError("Can't have both kinetic mapping and "
                             "commute mapping. Please only use one.")
        self.n_features = None
        self.n_observations_ = None
        self.n_sequences_ = None

        self._initialized = False

        # sufficient statistics, names as in the reference (tica.py:123-137)
        self._outer_0_to_T_lagged = None     # sum_t x_t x_{t+tau}^T
        self._sum_0_to_TminusTau = None      # sum_{t<n-tau} x_t
        self._sum_tau_to_T = None            # sum_{t>=tau} x_t
        self._sum_0_to_T = None              # sum_t x_t
        self._outer_0_to_TminusTau = None    # sum_{t<n-tau} x_t x_t^T
        self._outer_offset_to_T = None       # sum_{t>=tau} x_t x_t^T

        self._components_ = None
        self._eigenvectors_ = None
        self._eigenvalues_ = None
        self._is_dirty = True

    # ------------------------------------------------------------------ state
    def _initialize(self, n_features):
        if self._initialized:
            return
        if self.n_components is None:
            self.n_components = n_features
        self.n_features = n_features
        self.n_observations_ = 0
        self.n_sequences_ = 0
        self._outer_0_to_T_lagged = np.zeros((n_features, n_features))
        self._sum_0_to_TminusTau = np.zeros(n_features)
        self._sum_tau_to_T = np.zeros(n_features)
        self._sum_0_to_T = np.zeros(n_features)
        self._outer_0_to_TminusTau = np.zeros((n_features, n_features))
        self._outer_offset_to_T = np.zeros((n_features, n_features))
        self._initialized = True

    # ---------------------------------------------------------- eigenproblem
    def _solve(self):
        if not self._is_dirty:
            # n_components may have been raised after the last solve
            if len(self._eigenvalues_) >= self.n_components:
                return
        if self.n_observations_ == 0:
            raise RuntimeError('The model must be fit() before use.')

        lhs = self.offset_correlation_
        rhs = self.covariance_

        if not np.allclose(lhs, lhs.T):
            raise RuntimeError('offset correlation matrix is not symmetric')
        if not np.allclose(rhs, rhs.T):
            raise RuntimeError('correlation matrix is not symmetric')

        vals, vecs = _eigh_top(lhs, rhs, self.n_features - self.n_components,
                               self.n_features - 1)
        ind = np.argsort(vals)[::-1]
        self._eigenvalues_ = vals[ind]
        self._eigenvectors_ = vecs[:, ind]
        self._is_dirty = False

    @property
    def score_(self):
        """Training GMRQ: the sum of the first ``n_components`` eigenvalues."""
        self._solve()
        return self._eigenvalues_[:self.n_components].sum()

    @property
    def eigenvectors_(self):
        self._solve()
        return self._eigenvectors_[:, :self.n_components]

    @property
    def eigenvalues_(self):
        self._solve()
        return self._eigenvalues_[:self.n_components]

    @property
    def timescales_(self):
        self._solve()
        return -1. * self.lag_time / np.log(self._eigenvalues_[:self.n_components])

    @property
    def components_(self):
        return self.eigenvectors_[:, 0:self.n_components].T

    def _two_N(self):
        return 2 * (self.n_observations_ - self.lag_time * self.n_sequences_)

    @property
    def means_(self):
        return (self._sum_0_to_TminusTau + self._sum_tau_to_T) / float(self._two_N())

    @property
    def offset_correlation_(self):
        two_N = self._two_N()
        term = (self._outer_0_to_T_lagged + self._outer_0_to_T_lagged.T) / two_N
        means = self.means_
        return term - np.outer(means, means)

    @property
    def covariance_(self):
        two_N = self._two_N()
        term = (self._outer_0_to_TminusTau + self._outer_offset_to_T) / two_N
        means = self.means_
        S = term - np.outer(means, means)
        if self.shrinkage is None:
            sigma, self.shrinkage_ = rao_blackwell_ledoit_wolf(S, n=self.n_observations_)
        else:
            self.shrinkage_ = self.shrinkage
            p = self.n_features
            F = (np.trace(S) / p) * np.eye(p)
            sigma = (1 - self.shrinkage) * S + self.shrinkage * F
        return sigma

    # --------------------------------------------------------------- fitting
    def fit(self, sequences, y=None):
        """Fit the model with a collection of sequences (not online: previous
        state is discarded).

        Parameters
        ----------
        sequences: list of array-like, each of shape (n_samples_i, n_features)
            NumPy arrays, or torch tensors on the host or on the GPU.
        y : None
            Ignored

        Returns
        -------
        self : object
        """
        self._initialized = False
        check_iter_of_sequences(sequences, max_iter=3)  # input may be lazy
        from .._device import resolve_devices
        devs = resolve_devices(self.devices)
        if devs is not None:
            self._fit_many_on_devices(sequences, devs)
        else:
            self._fit_many(sequences)
        if self.n_sequences_ == 0:
            raise ValueError('All sequences were shorter than '
                             'the lag time, %d' % self.lag_time)
        return self

    def partial_fit(self, X):
        """Update the model with one more sequence X, shape (n_samples, n_features)."""
        self._fit(X)
        return self

    def _fit(self, X):
        self._fit_many([X])

    # batches of at most this many bytes of HOST data are staged on the device: the upload of
    # batch i+1 (side stream, _device.HostUploader) overlaps K1 of batch i (compute stream)
    _stage_bytes = 1 << 30

    def _fit_many(self, sequences):
        """Device accumulation of any number of sequences (tica.py:401-424 per
        sequence).  Host arrays are staged in bounded batches; device tensors are
        consumed in place."""
        import torch
        from .. import _device as dev

        batch, batch_bytes = [], 0
        state = {"acc": None}

        def flush():
            # every batch ADDS into one device accumulator; nothing is read back (and nothing
            # synchronises) until the last batch has been enqueued
            if batch:
                state["acc"] = self._accumulate_device(batch, acc=state["acc"])
                del batch[:]

        for X in sequences:
            if is_tensor(X):
                if X.ndim == 1:
                    X = X.unsqueeze(0)
                if X.dtype not in (torch.float32, torch.float64):
                    X = X.to(torch.float64)
                # device tensors count too: a lazy stream (io.NumpyDirStream) must not
                # pile its whole dataset up in `batch`
                nbytes = X.numel() * X.element_size()
            else:
                X = np.atleast_2d(np.asarray(X))
                if X.dtype not in (np.float32, np.float64):
                    X = X.astype(np.float64)   # tica.py:402 widens everything
                nbytes = X.nbytes
            if X.ndim != 2:
                raise ValueError('sequences must be a list of sequences')
            n, d = int(X.shape[0]), int(X.shape[1])
            if d > n:
                warnings.warn("The number of features (%d) is greater than the length of the "
                              "data (%d). The covariance matrix is not guaranteed to be "
                              "positive definite." % (d, n))
            self._initialize(d)
            if d != self.n_features:
                raise ValueError("sequence has %d features, model has %d" % (d, self.n_features))
            if not n > self.lag_time:
                warnings.warn("length of data (%d) is too short for the lag time (%d)"
                              % (n, self.lag_time))
                continue
            if batch and batch_bytes + nbytes > self._stage_bytes:
                flush()
                batch_bytes = 0
            batch.append(X)
            batch_bytes += nbytes
        flush()
        if state["acc"] is not None:
            self._add_packed(state["acc"].cpu().numpy())       # the one synchronisation of a fit

    def _fit_many_on_devices(self, sequences, devs):
        """`devices=`: whole sequences are dealt to the GPUs (longest first), one thread per GPU
        stages and accumulates its share exactly like _fit_many, and the packed float64
        accumulators are added on the host in device order (the all-reduce of the torchrun path,
        tica.py:417-422 being a plain sum over sequences)."""
        import threading
        import torch
        from .. import parallel as P
        good = []
        for X in sequences:
            if is_tensor(X):
                X = X.detach()
                if X.ndim == 1:
                    X = X.unsqueeze(0)
                if X.dtype not in (torch.float32, torch.float64):
                    X = X.to(torch.float64)
            else:
                X = np.atleast_2d(np.asarray(X))
                if X.dtype not in (np.float32, np.float64):
                    X = X.astype(np.float64)
            if X.ndim != 2:
                raise ValueError('sequences must be a list of sequences')
            n, d = int(X.shape[0]), int(X.shape[1])
            if d > n:
                warnings.warn("The number of features (%d) is greater than the length of the "
                              "data (%d). The covariance matrix is not guaranteed to be "
                              "positive definite." % (d, n))
            self._initialize(d)
            if d != self.n_features:
                raise ValueError("sequence has %d features, model has %d" % (d, self.n_features))
            if not n > self.lag_time:
                warnings.warn("length of data (%d) is too short for the lag time (%d)"
                              % (n, self.lag_time))
                continue
            good.append(X)
        if not good:
            return
        shares = [sh for sh in P.shard_sequences([int(x.shape[0]) for x in good], len(devs)) if sh]
        packed, errors = [None] * len(shares), []

        def work(r):
            try:
                with torch.cuda.device(devs[r]):
                    acc, batch, nbytes = None, [], 0
                    for i in shares[r]:
                        X = good[i]
                        b = X.numel() * X.element_size() if is_tensor(X) else X.nbytes
                        if batch and nbytes + b > self._stage_bytes:
                            acc = self._accumulate_device(batch, acc=acc)
                            batch, nbytes = [], 0
                        batch.append(X)
                        nbytes += b
                    if batch:
                        acc = self._accumulate_device(batch, acc=acc)
                    packed[r] = acc.cpu().numpy()
            except BaseException as e:      # noqa: B902
                errors.append(e)

        threads = [threading.Thread(target=work, args=(r,), name="msmb200-tica-%d" % devs[r])
                   for r in range(len(shares))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        total = packed[0]
        for extra in packed[1:]:
            total = total + extra
        self._add_packed(total)

    def _accumulate_device(self, seqs, acc=None):
        """Run K1 over `seqs` (host arrays are uploaded, device tensors used in
        place); ADDS their statistics to `acc` (a new zeroed one by default) and returns it: the
        packed float64 accumulator as a CUDA tensor (layout: include/msmb200.h).  Device work is
        enqueued on torch's current stream and not waited for; host arrays are uploaded through
        the pinned ring, which returns once the last chunk has been issued."""
        import torch
        from .. import _device as dev
        _lib.require_gpu()
        D = self.n_features
        dts = {torch.float64 if (s.dtype in (np.float64, torch.float64)) else torch.float32
               for s in seqs}
        dtype = torch.float64 if torch.float64 in dts else torch.float32
        np_dtype = np.float64 if dtype == torch.float64 else np.float32
        dev_seqs = [None] * len(seqs)
        host = [(i, dev.host_array(s).astype(np_dtype, copy=False))
                for i, s in enumerate(seqs) if not is_tensor(s)]
        if host:
            # one device block for the batch's host arrays, rows padded to 16-byte multiples
            # apart is not needed: D * 4 is the pitch and every slot starts on a row boundary
            rows = [a.shape[0] for _, a in host]
            block = torch.empty((sum(rows), D), dtype=dtype, device="cuda")
            pairs, o = [], 0
            for (i, a), n in zip(host, rows):
                dev_seqs[i] = block[o:o + n]
                pairs.append((a, dev_seqs[i]))
                o += n
            dev.uploader().upload(pairs)
        for i, s in enumerate(seqs):
            if is_tensor(s):
                t = s if s.dtype == dtype else s.to(dtype)
                t = t.cuda(non_blocking=True) if not t.is_cuda else t
                dev_seqs[i] = t.contiguous()
        n_seq = len(dev_seqs)
        ptrs = (ctypes.c_void_p * n_seq)(*[t.data_ptr() for t in dev_seqs])
        rows = (ctypes.c_int64 * n_seq)(*[int(t.shape[0]) for t in dev_seqs])
        lib = _lib.load()
        acc_len = lib.msmb200_tica_acc_len(D)
        if acc is None:
            acc = torch.zeros(acc_len, dtype=torch.float64, device="cuda")
        engine = _lib.ENGINES[self.engine]
        ws_bytes = lib.msmb200_tica_workspace_bytes(D, engine)
        ws = dev.workspace().get(("tica", D), ws_bytes)
        _lib.call("msmb200_tica_accumulate", ptrs, rows, n_seq, D, D, dev.dtype_id(dev_seqs[0]),
                  int(self.lag_time), engine, dev.ptr(acc), dev.ptr(ws), ws.numel(),
                  dev.stream_ptr())
        return acc

    def _add_packed(self, packed):
        """Fold one packed device accumulator (layout: include/msmb200.h) into the
        reference-named NumPy attributes."""
        D = self.n_features
        DD = D * D
        S = packed[3 * DD + 2 * D: 3 * DD + 3 * D]
        if not np.isfinite(S.sum()) or not np.isfinite(packed[:3 * DD]).all():
            raise ValueError("Input contains NaN, infinity or a value too large for "
                             "dtype('float64').")
        self._outer_0_to_T_lagged += packed[0:DD].reshape(D, D)
        self._outer_0_to_TminusTau += packed[DD:2 * DD].reshape(D, D)
        self._outer_offset_to_T += packed[2 * DD:3 * DD].reshape(D, D)
        self._sum_0_to_TminusTau += packed[3 * DD: 3 * DD + D]
        self._sum_tau_to_T += packed[3 * DD + D: 3 * DD + 2 * D]
        self._sum_0_to_T += S
        self.n_observations_ += int(round(packed[3 * DD + 3 * D]))
        self.n_sequences_ += int(round(packed[3 * DD + 3 * D + 1]))
        self._is_dirty = True

    # ------------------------------------------------------------ projection
    def transform(self, sequences):
        """Project each sequence onto the tICs.

        Parameters
        ----------
        sequences: list of array-like, each of shape (n_samples_i, n_features)

        Returns
        -------
        sequence_new : list of array-like, each of shape (n_samples_i, n_components)
            float64.  NumPy in -> NumPy out; CUDA tensor in -> CUDA tensor out.
        """
        import torch
        from .. import _device as dev
        check_iter_of_sequences(sequences, max_iter=3)
        _lib.require_gpu()
        k = int(self.n_components)
        means = torch.from_numpy(np.ascontiguousarray(self.means_)).cuda()
        comps = torch.from_numpy(np.ascontiguousarray(self.components_)).cuda()
        scale = None
        if self.kinetic_mapping:
            scale = np.array(self.eigenvalues_, dtype=np.float64)
        if self.commute_mapping:
            # same damping of fast timescales as tica.py:338-351
            ts = self.timescales_
            reg = 0.5 * ts * np.tanh(np.pi * ((ts - self.lag_time) / self.lag_time) + 1)
            cm = np.sqrt(reg / 2)
            scale = cm if scale is None else scale * cm
        d_scale = None if scale is None else torch.from_numpy(np.ascontiguousarray(scale)).cuda()

        out = []
        for X in sequences:
            on_device = is_tensor(X) and X.is_cuda
            t = dev.to_device(X)
            if t.shape[1] != self.n_features:
                raise ValueError("sequence has %d features, model has %d"
                                 % (t.shape[1], self.n_features))
            y = torch.empty((t.shape[0], k), dtype=torch.float64, device="cuda")
            _lib.call("msmb200_tica_transform", dev.ptr(t), int(t.shape[0]), int(t.shape[1]),
                      int(t.shape[1]), dev.dtype_id(t), dev.ptr(means), dev.ptr(comps),
                      dev.ptr(d_scale), k, dev.ptr(y), dev.stream_ptr())
            if self.commute_mapping:
                y = torch.nan_to_num(y)
            out.append(y if on_device else y.cpu().numpy())
        return out

    def partial_transform(self, features):
        """Project one sequence, shape (n_samples, n_features)."""
        return self.transform([features])[0]

    def fit_transform(self, sequences, y=None):
        """``fit(sequences)`` then ``transform(sequences)``."""
        self.fit(sequences)
        return self.transform(sequences)

    def score(self, sequences, y=None):
        """Generalised matrix Rayleigh quotient of this model's eigenvectors on
        new data (McGibbon & Pande, J. Chem. Phys. 142, 124105 (2015))."""
        assert self._initialized
        V = self.eigenvectors_
        m2 = self.__class__(shrinkage=self.shrinkage, n_components=self.n_components,
                            lag_time=self.lag_time)
        if hasattr(m2, 'engine'):
            m2.engine = self.engine
        for X in sequences:
            m2.partial_fit(X)
        numerator = V.T.dot(m2.offset_correlation_).dot(V)
        denominator = V.T.dot(m2.covariance_).dot(V)
        try:
            trace = np.trace(numerator.dot(np.linalg.inv(denominator)))
        except np.linalg.LinAlgError:
            trace = np.nan
        return trace

    def summarize(self):
        """Some summary information."""
        self.covariance_   # forces shrinkage_ to be computed
        return """time-structure based Independent Components Analysis (tICA)
-----------------------------------------------------------
n_components        : {n_components}
shrinkage           : {shrinkage}
lag_time            : {lag_time}
kinetic_mapping     : {kinetic_mapping}

Top 5 timescales :
{timescales}

Top 5 eigenvalues :
{eigenvalues}
""".format(n_components=self.n_components, lag_time=self.lag_time,
           shrinkage=self.shrinkage_, kinetic_mapping=self.kinetic_mapping,
           timescales=self.timescales_[:5], eigenvalues=self.eigenvalues_[:5])


def rao_blackwell_ledoit_wolf(S, n):
    """Rao-Blackwellised Ledoit-Wolf shrinkage of a sample covariance matrix
    (Chen, Wiesel & Hero, ICASSP 2009); same formula as tica.py:492-524.

    Returns
    -------
    sigma : array, shape=(p, p)
    shrinkage : float
    """
    p = len(S)
    assert S.shape == (p, p)

    alpha = (n - 2) / (n * (n + 2))
    beta = ((p + 1) * n - 2) / (n * (n + 2))

    trace_S2 = np.sum(S * S)
    U = ((p * trace_S2 / np.trace(S) ** 2) - 1)
    rho = min(alpha + beta / U, 1)

    F = (np.trace(S) / p) * np.eye(p)
    return (1 - rho) * S + rho * F, rho
