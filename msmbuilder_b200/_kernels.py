"""Device-level wrappers: CUDA tensors in, CUDA tensors out, one C-ABI call each.

Everything here launches on torch's current stream and returns without
synchronising unless a Python scalar is requested.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from . import _device as dev

_I64 = torch.int64


def _rows_tensor(rows):
    if rows is None:
        return None
    if isinstance(rows, torch.Tensor):
        return rows.to(device="cuda", dtype=_I64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int64)).cuda()


# --------------------------------------------------------------------------- K2
class KCentersState(object):
    """Per-shard state of a k-centers run on a (n, d) device tensor (vector
    metrics) or a centred (n, n_atoms, 3) tensor + traces (rmsd)."""

    def __init__(self, data, metric, traces=None, row_offset=0, max_centres=0):
        _lib.require_gpu()
        self.data = data
        self.rmsd = (metric == "rmsd")
        self.metric = None if self.rmsd else _lib.metric_id(metric)
        self.traces = traces
        self.n = int(data.shape[0])
        self.row_offset = int(row_offset)
        if self.rmsd:
            self.n_atoms = int(data.shape[1])
            self.row_elems = self.n_atoms * 3 + 1     # coords + trace
            self.dtype = _lib.F32
        else:
            self.d = int(data.shape[1])
            self.row_elems = self.d
            self.dtype = dev.dtype_id(data)
        lib = _lib.load()
        self.cand_bytes = int(lib.msmb200_candidate_bytes(self.row_elems, self.dtype))
        self.distances = torch.full((self.n,), float("inf"), dtype=torch.float64, device="cuda")
        self.labels = torch.zeros((self.n,), dtype=torch.int32, device="cuda")
        ws_bytes = int(lib.msmb200_kcenters_workspace_bytes(0))
        if self.rmsd:
            # the pruned RMSD pass keeps a compact list of the frames it has to look at
            ws_bytes = max(ws_bytes, int(lib.msmb200_rmsd_pass_workspace_bytes(self.n, int(max_centres))))
        self.max_centres = int(max_centres)
        self.ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
        self.local = torch.zeros(self.cand_bytes, dtype=torch.uint8, device="cuda")

    def payload_ptr(self, cand):
        return ctypes.c_void_p(cand.data_ptr() + 16)

    def seed(self, cand, local_row):
        """Write the seed centre (kcenters.py:84) into candidate buffer `cand`."""
        if self.rmsd:
            flat = self.data.reshape(self.n, -1)
            view = cand[16:16 + 4 * self.row_elems].view(torch.float32)
            view[:self.n_atoms * 3].copy_(flat[local_row])
            view[self.n_atoms * 3:].copy_(self.traces[local_row:local_row + 1])
            hdr = cand[:16].view(torch.float64)
            hdr[0] = float("inf")
            cand[8:16].view(torch.int64)[0] = self.row_offset + int(local_row)
        else:
            _lib.call("msmb200_candidate_from_row", dev.ptr(self.data), int(local_row), self.d,
                      self.d, self.dtype, self.row_offset, dev.ptr(cand), dev.stream_ptr())

    def run_pass(self, center_cand, label, out_cand=None, start=0, ring=None):
        """One pass against the centre stored in `center_cand`; this shard's
        farthest frame is written to `out_cand` (default: self.local).  `start`
        restricts the pass to rows [start, n) (RegularSpatial never looks back).
        `ring` (RMSD only): the (k + 1, cand_bytes) tensor whose slot j holds centre j and whose
        slot `label` is `center_cand` -- the pass then skips the frames the triangle inequality
        rules out (same result, see msmb200_rmsd_kcenters_pass_pruned)."""
        out = self.local if out_cand is None else out_cand
        start = int(start)
        n = self.n - start
        if n <= 0:
            return out
        data = self.data[start:]
        dist = self.distances[start:]
        lab = self.labels[start:]
        if (self.rmsd and ring is not None and start == 0 and 0 < int(label) <= self.max_centres
                and not os.environ.get("MSMB200_RMSD_NO_PRUNE")):
            _lib.call("msmb200_rmsd_kcenters_pass_pruned", dev.ptr(data), dev.ptr(self.traces),
                      n, self.n_atoms, dev.ptr(ring), int(ring.stride(0)), int(label),
                      dev.ptr(dist), dev.ptr(lab), self.row_offset,
                      dev.ptr(out), dev.ptr(self.ws), self.ws.numel(), dev.stream_ptr())
        elif self.rmsd:
            _lib.call("msmb200_rmsd_kcenters_pass", dev.ptr(data), dev.ptr(self.traces[start:]),
                      n, self.n_atoms, self.payload_ptr(center_cand), int(label),
                      dev.ptr(dist), dev.ptr(lab), self.row_offset + start,
                      dev.ptr(out), dev.ptr(self.ws), self.ws.numel(), dev.stream_ptr())
        else:
            _lib.call("msmb200_kcenters_pass", dev.ptr(data), n, self.d, self.d,
                      self.dtype, self.metric, self.payload_ptr(center_cand), int(label),
                      dev.ptr(dist), dev.ptr(lab), self.row_offset + start,
                      dev.ptr(out), dev.ptr(self.ws), self.ws.numel(), dev.stream_ptr())
        return out

    def select(self, gathered, n_cand, out_cand):
        _lib.call("msmb200_candidate_select", dev.ptr(gathered), int(n_cand), self.cand_bytes,
                  self.row_elems, self.dtype, dev.ptr(out_cand), dev.stream_ptr())


def kcenters_fit(data, n_clusters, metric, seed_index, traces=None, lookahead=True, stats=None):
    """Single-GPU Gonzalez k-centers (kcenters.py:79-102).  float32 (sq)euclidean input takes
    the look-ahead path (kcenters_fit_lookahead: several centres per read of the frames);
    everything else runs k passes enqueued back to back, the arg-max that names the next
    centre never leaving the device.  Both give the reference's centres and labels.

    Returns (cluster_ids int64[k] tensor, distances f64[n], labels i32[n])."""
    if traces is None and lookahead and lookahead_supported(data, metric):
        ids, _, distances, labels = kcenters_fit_lookahead(data, n_clusters, metric, seed_index,
                                                           stats=stats)
        return ids, distances, labels
    k = int(n_clusters)
    st = KCentersState(data, metric, traces=traces, max_centres=k)
    if stats is not None:
        stats["passes"] = k
    # ring of k candidate slots: slot i holds centre i (its index is cluster_ids_[i])
    ring = torch.zeros((k + 1, st.cand_bytes), dtype=torch.uint8, device="cuda")
    st.seed(ring[0], int(seed_index))
    for i in range(k):
        st.run_pass(ring[i], i, out_cand=ring[i + 1], ring=ring)
    ids = ring[:k, 8:16].contiguous().view(torch.int64).reshape(k)
    return ids, st.distances, st.labels


# ------------------------------------------------------------------ K2b (look-ahead)
LOOKAHEAD_T_CAP = 1024    # candidates kept per shard and pass (chain cost grows with it: +0.2 ms per doubling)
LOOKAHEAD_J_CAP = 16      # centres applied by one fused pass at most (fewer for wide frames: they
                          # share 15 KB of shared memory next to the cp.async frame ring)
CENTERS_HEADER = 32       # sizeof(CentersHeader)


def lookahead_supported(data, metric):
    """True when the fused look-ahead kernels take this input (float32, euclidean or
    sqeuclidean, 16-byte aligned rows); otherwise KCenters runs one pass per centre."""
    if metric not in ("euclidean", "sqeuclidean") or data.dtype != torch.float32 or data.dim() != 2:
        return False
    if data.data_ptr() % 16 or not data.is_contiguous():
        return False
    d = int(data.shape[1])
    return bool(_lib.load().msmb200_kcenters_lookahead_supported(d, d, _lib.F32, _lib.metric_id(metric)))


class LookaheadState(object):
    """Per-shard buffers of the look-ahead k-centers (csrc/kcenters_lookahead.cu)."""

    def __init__(self, data, metric, row_offset=0, t_cap=LOOKAHEAD_T_CAP, j_cap=None):
        _lib.require_gpu()
        if j_cap is None:
            # shared memory of a fused pass (two blocks per SM: 113 KB each): frame ring <= 96 KB, the lane
            # records 1.5 KB, the centres (padded to chunks of 4) get 15 KB
            fit = (15 * 1024) // (4 * int(data.shape[1]))
            j_cap = max(1, min(LOOKAHEAD_J_CAP, fit // 4 * 4 if fit >= 4 else fit))
        lib = _lib.load()
        self.data = data
        self.metric = _lib.metric_id(metric)
        self.n, self.d = int(data.shape[0]), int(data.shape[1])
        self.row_offset = int(row_offset)
        self.t_cap, self.j_cap = int(t_cap), int(j_cap)
        self.distances = torch.full((self.n,), float("inf"), dtype=torch.float64, device="cuda")
        self.labels = torch.zeros((self.n,), dtype=torch.int32, device="cuda")
        self.lane = torch.zeros(int(lib.msmb200_kcenters_lane_bytes(0)), dtype=torch.uint8, device="cuda")
        self.set_bytes = int(lib.msmb200_kcenters_set_bytes(self.d, self.t_cap))
        self.centers_bytes = int(lib.msmb200_kcenters_centers_bytes(self.d, self.j_cap))
        self.cset = torch.zeros(self.set_bytes, dtype=torch.uint8, device="cuda")
        self.centers = torch.zeros(self.centers_bytes, dtype=torch.uint8, device="cuda")

    # views into a centres blob
    def centers_ids(self, blob=None):
        b = self.centers if blob is None else blob
        return b[CENTERS_HEADER:CENTERS_HEADER + 8 * self.j_cap].view(torch.int64)

    def centers_rows(self, blob=None):
        b = self.centers if blob is None else blob
        off = CENTERS_HEADER + 8 * self.j_cap
        return b[off:off + 4 * self.j_cap * self.d].view(torch.float32).reshape(self.j_cap, self.d)

    def seed(self, global_row):
        """The seed centre (kcenters.py:84) as the first pending centre; on a rank that does not
        hold the row the blob is zeroed (the caller broadcasts the owner's)."""
        self.centers.zero_()
        local = int(global_row) - self.row_offset
        if 0 <= local < self.n:
            self.centers[:8].view(torch.int32)[0] = 1
            self.centers[:8].view(torch.int32)[1] = self.j_cap
            self.centers_ids()[0] = int(global_row)
            self.centers_rows()[0].copy_(self.data[local])
            return True
        return False

    def multi_pass(self, n_centers, label0, first):
        if self.n == 0:
            return
        _lib.call("msmb200_kcenters_multi_pass", dev.ptr(self.data), self.n, self.d, self.d,
                  _lib.F32, self.metric, dev.ptr(self.centers), int(n_centers), self.j_cap,
                  int(label0), 1 if first else 0, dev.ptr(self.distances), dev.ptr(self.labels),
                  self.row_offset, dev.ptr(self.lane), self.lane.numel(), dev.stream_ptr())

    def select(self):
        if self.n == 0:
            self.cset.zero_()
            self.cset[8:16].view(torch.float64)[0] = float("-inf")
            return self.cset
        _lib.call("msmb200_kcenters_select", dev.ptr(self.data), self.n, self.d, self.d,
                  self.row_offset, dev.ptr(self.lane), self.t_cap, dev.ptr(self.cset),
                  dev.stream_ptr())
        return self.cset

    def chain(self, sets, n_sets, k_remaining):
        _lib.call("msmb200_kcenters_chain", dev.ptr(sets), int(n_sets), self.set_bytes, self.d,
                  self.metric, int(k_remaining), self.j_cap, dev.ptr(self.centers), dev.stream_ptr())
        return int(self.centers[:4].view(torch.int32).item())     # the one host sync per chain


def kcenters_fit_lookahead(data, n_clusters, metric, seed_index, gather_sets=None, bcast=None,
                           row_offset=0, stats=None, state=None):
    """Gonzalez k-centers with look-ahead: the reference's centres, labels and distances
    (kcenters.py:79-102) in (number of chains + 1) reads of the frames instead of k.

    gather_sets(set_blob) -> (all_sets, n_sets) and bcast(blob) make it rank-collective
    (parallel.kcenters_fit_gpu); single GPU by default.  `state` is the per-shard engine
    (default: the CUDA LookaheadState; the gloo protocol tests pass a host mirror with the
    same five methods).

    Returns (cluster_ids int64[k] (global rows), centre rows (k, d), distances f64[n],
    labels i32[n]) as tensors on the state's device."""
    st = LookaheadState(data, metric, row_offset=row_offset) if state is None else state
    k = int(n_clusters)
    ids, rows = [], []
    st.seed(seed_index)
    if bcast is not None:
        bcast(st.centers)
    done, n_pending, passes = 0, 1, 0
    while True:
        timed = stats is not None and stats.get("time_passes") and torch.cuda.is_available()
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        st.multi_pass(n_pending, done, first=(done == 0))
        if timed:
            e1.record()
            stats.setdefault("pass_events", []).append((n_pending, e0, e1))
        passes += 1
        ids.append(st.centers_ids()[:n_pending].clone())
        rows.append(st.centers_rows()[:n_pending].clone())
        done += n_pending
        if done >= k:
            break
        sets, n_sets = st.select(), 1
        if gather_sets is not None:
            sets, n_sets = gather_sets(sets)
        n_pending = st.chain(sets, n_sets, k - done)
        if n_pending < 1:
            raise _lib.Msmb200Error("k-centers look-ahead found no next centre (empty input?)")
    if stats is not None:
        stats["passes"] = passes
    return torch.cat(ids), torch.cat(rows), st.distances, st.labels


def regular_spatial_fit(data, d_min, metric, traces=None):
    """RegularSpatial.fit (cluster/regularspatial.py:70-77): frame i becomes a centre
    when every centre found before it is farther than d_min.  The reference asks
    that question frame by frame (one `dist` call per frame against the centre
    list); here each new centre costs ONE streaming pass over the frames after it
    (running minimum, same arithmetic as K2) plus one scan for the first frame
    whose minimum still exceeds d_min -- identical centres, O(n_centres) passes.

    Returns the list of centre indices."""
    st = KCentersState(data, metric, traces=traces)
    # a frame with a NaN coordinate is never a centre in the reference (`np.all(d > d_min)` is False for
    # NaN, regularspatial.py:70-77); here its distance to every centre is NaN too, the strict `<` of the
    # pass would leave its running minimum at +inf and first_above would promote it: pin it to -inf
    bad = torch.isnan(data.reshape(st.n, -1)).any(dim=1)
    if st.n > 1 and bool(bad[1:].any()):
        st.distances[bad] = float("-inf")
    cand = torch.zeros(st.cand_bytes, dtype=torch.uint8, device="cuda")
    nxt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ids = [0]
    c = 0
    d_min = float(d_min)
    while True:
        st.seed(cand, c)
        st.run_pass(cand, len(ids) - 1, start=c + 1)
        _lib.call("msmb200_first_above", dev.ptr(st.distances), st.n, c + 1, d_min,
                  dev.ptr(nxt), dev.stream_ptr())
        c = int(nxt.item())
        if c < 0:
            break
        ids.append(c)
    return ids


# --------------------------------------------------------------------------- K3
def _same_width(X, Y, msg='X and Y must have the same number of columns'):
    """The C ABI takes ONE width for both operands: a mismatch would read out of bounds on the
    device.  The reference raises (libdistance.pyx:378,398 assert; :447,456 ValueError)."""
    if X.dim() != 2 or Y.dim() != 2 or int(X.shape[1]) != int(Y.shape[1]):
        raise ValueError(msg)


def assign_nearest(X, Y, metric, rows=None, want_min_dist=False):
    """labels (i32), min_dist (f64 or None), inertia (0-d f64 tensor)."""
    _lib.require_gpu()
    lib = _lib.load()
    rows_t = _rows_tensor(rows)
    n_out = int(X.shape[0]) if rows_t is None else int(rows_t.numel())
    labels = torch.empty((n_out,), dtype=torch.int32, device="cuda")
    min_dist = torch.empty((n_out,), dtype=torch.float64, device="cuda") if want_min_dist else None
    inertia = torch.zeros((1,), dtype=torch.float64, device="cuda")
    if n_out == 0:
        return labels, min_dist, inertia[0]
    if metric == "rmsd":
        raise TypeError("use rmsd_assign_nearest for metric='rmsd'")
    m = _lib.metric_id(metric)
    if X.dtype != Y.dtype:
        raise TypeError('X and y must be both float32 or float64')
    _same_width(X, Y)
    d, k = int(X.shape[1]), int(Y.shape[0])
    ws_bytes = int(lib.msmb200_assign_workspace_bytes(n_out, k, d))
    ws = dev.workspace().get("assign", ws_bytes)
    _lib.call("msmb200_assign_nearest", dev.ptr(X), int(X.shape[0]), d, d, dev.dtype_id(X),
              dev.ptr(Y), k, m, dev.ptr(rows_t), n_out if rows_t is not None else 0,
              dev.ptr(labels), dev.ptr(min_dist), dev.ptr(inertia), dev.ptr(ws), ws.numel(),
              dev.stream_ptr())
    return labels, min_dist, inertia[0]


def rmsd_assign_nearest(xyz, traces, Y, Y_traces, rows=None, want_min_dist=False):
    _lib.require_gpu()
    if int(xyz.shape[1]) != int(Y.shape[1]):
        raise ValueError("Input trajectories must have same number of atoms. found %d and %d."
                         % (int(xyz.shape[1]), int(Y.shape[1])))          # libdistance.pyx:320-322
    rows_t = _rows_tensor(rows)
    n_out = int(xyz.shape[0]) if rows_t is None else int(rows_t.numel())
    labels = torch.empty((n_out,), dtype=torch.int32, device="cuda")
    min_dist = torch.empty((n_out,), dtype=torch.float64, device="cuda") if want_min_dist else None
    inertia = torch.zeros((1,), dtype=torch.float64, device="cuda")
    if n_out:
        _lib.call("msmb200_rmsd_assign_nearest", dev.ptr(xyz), dev.ptr(traces),
                  int(xyz.shape[0]), int(xyz.shape[1]), dev.ptr(Y), dev.ptr(Y_traces),
                  int(Y.shape[0]), dev.ptr(rows_t), n_out if rows_t is not None else 0,
                  dev.ptr(labels), dev.ptr(min_dist), dev.ptr(inertia), dev.stream_ptr())
    return labels, min_dist, inertia[0]


# --------------------------------------------------------------------------- K4
def dist(X, y, metric, rows=None):
    _lib.require_gpu()
    if int(y.numel()) != int(X.shape[1]):
        raise ValueError('X and y must have the same number of columns')   # libdistance.pyx:378,398 assert
    rows_t = _rows_tensor(rows)
    n_out = int(X.shape[0]) if rows_t is None else int(rows_t.numel())
    out = torch.empty((n_out,), dtype=torch.float64, device="cuda")
    if n_out:
        _lib.call("msmb200_dist", dev.ptr(X), int(X.shape[0]), int(X.shape[1]), int(X.shape[1]),
                  dev.dtype_id(X), dev.ptr(y), _lib.metric_id(metric), dev.ptr(rows_t),
                  n_out if rows_t is not None else 0, dev.ptr(out), dev.stream_ptr())
    return out


def cdist(XA, XB, metric):
    _lib.require_gpu()
    _same_width(XA, XB, 'XA and XB must have the same number of columns')
    if XA.dtype != XB.dtype:
        raise TypeError('XA and XB must be both float32 or float64')
    out = torch.empty((int(XA.shape[0]), int(XB.shape[0])), dtype=torch.float64, device="cuda")
    if out.numel():
        _lib.call("msmb200_cdist", dev.ptr(XA), int(XA.shape[0]), dev.ptr(XB), int(XB.shape[0]),
                  int(XA.shape[1]), dev.dtype_id(XA), _lib.metric_id(metric), dev.ptr(out),
                  dev.stream_ptr())
    return out


def pdist(X, metric, rows=None):
    _lib.require_gpu()
    rows_t = _rows_tensor(rows)
    m = int(X.shape[0]) if rows_t is None else int(rows_t.numel())
    out = torch.empty((m * (m - 1) // 2,), dtype=torch.float64, device="cuda")
    if out.numel():
        _lib.call("msmb200_pdist", dev.ptr(X), int(X.shape[0]), int(X.shape[1]), int(X.shape[1]),
                  dev.dtype_id(X), _lib.metric_id(metric), dev.ptr(rows_t),
                  m if rows_t is not None else 0, dev.ptr(out), dev.stream_ptr())
    return out


def sumdist(X, metric, pairs):
    _lib.require_gpu()
    pairs_t = _rows_tensor(np.asarray(pairs).reshape(-1, 2) if not isinstance(pairs, torch.Tensor)
                           else pairs)
    out = torch.zeros((1,), dtype=torch.float64, device="cuda")
    _lib.call("msmb200_sumdist", dev.ptr(X), int(X.shape[0]), int(X.shape[1]), int(X.shape[1]),
              dev.dtype_id(X), _lib.metric_id(metric), dev.ptr(pairs_t),
              int(pairs_t.shape[0]), dev.ptr(out), dev.stream_ptr())
    return out[0]


# --------------------------------------------------------------------------- RMSD
def rmsd_center(xyz):
    """Centre frames IN PLACE (like Trajectory.center_coordinates, cluster/base.py:68)
    and return the float32 traces."""
    _lib.require_gpu()
    traces = torch.empty((int(xyz.shape[0]),), dtype=torch.float32, device="cuda")
    _lib.call("msmb200_rmsd_center", dev.ptr(xyz), int(xyz.shape[0]), int(xyz.shape[1]),
              dev.ptr(traces), dev.stream_ptr())
    return traces


def rmsd_dist(xyz, traces, y, y_trace, rows=None):
    _lib.require_gpu()
    rows_t = _rows_tensor(rows)
    n_out = int(xyz.shape[0]) if rows_t is None else int(rows_t.numel())
    out = torch.empty((n_out,), dtype=torch.float64, device="cuda")
    if n_out:
        _lib.call("msmb200_rmsd_dist", dev.ptr(xyz), dev.ptr(traces), int(xyz.shape[0]),
                  int(xyz.shape[1]), dev.ptr(y), ctypes.c_float(float(y_trace)), dev.ptr(rows_t),
                  n_out if rows_t is not None else 0, dev.ptr(out), dev.stream_ptr())
    return out


def rmsd_pdist(xyz, traces, rows=None):
    _lib.require_gpu()
    rows_t = _rows_tensor(rows)
    m = int(xyz.shape[0]) if rows_t is None else int(rows_t.numel())
    out = torch.empty((m * (m - 1) // 2,), dtype=torch.float64, device="cuda")
    if out.numel():
        _lib.call("msmb200_rmsd_pdist", dev.ptr(xyz), dev.ptr(traces), int(xyz.shape[0]),
                  int(xyz.shape[1]), dev.ptr(rows_t), m if rows_t is not None else 0,
                  dev.ptr(out), dev.stream_ptr())
    return out


# ------------------------------------------------------------------- host k-medoids
def _random_starts(n_clusters, n_elements, n_pass, random):
    """The initial assignments kmedoids.cc:314-383 would draw inside C: per pass
    n_clusters - 1 binomials (every cluster keeps one element) and one shuffle,
    on the same RandomState in the same order."""
    starts = np.zeros((n_pass, n_elements), dtype=np.int64)
    for p in range(n_pass):
        row = starts[p]
        left = n_elements - n_clusters
        k = 0
        for i in range(n_clusters - 1):
            j = int(random.binomial(float(left), 1.0 / (n_clusters - i)))
            left -= j
            j += k + 1
            row[k:j] = i
            k = j
        row[k:] = n_clusters - 1
        random.shuffle(row)
    return starts


def kmedoids(n_clusters, distmatrix, n_pass, clusterid=None, random_state=None):
    """_kmedoids.kmedoids (cluster/_kmedoids.pyx:23-107): n_pass == 0 starts from
    `clusterid`, n_pass >= 1 from random assignments (cluster/kmedoids.py:92-94)."""
    from sklearn.utils import check_random_state
    dm = np.ascontiguousarray(distmatrix, dtype=np.float64)
    n_elements = int(1 + np.sqrt(8 * len(dm) + 1) / 2.0)
    if len(dm) != (n_elements * (n_elements - 1) / 2):
        raise ValueError('len(distmatrix)=%s is not a valid size of a condensed distance '
                         'matrix, which should be of size (N*(N-1)/2) for some positive '
                         'integer, N' % len(dm))
    if n_clusters > n_elements:
        raise ValueError('Number of clusters requested (%d) greater than '
                         'number of elements (%d)' % (n_clusters, n_elements))
    if clusterid is not None and len(clusterid) != n_elements:
        raise ValueError('clusterid must be None or an array of length n_elements')
    if n_pass < 0:
        raise ValueError('n_pass must be greater than or equal to zero.')
    cid = np.zeros(n_elements, dtype=np.int64) if clusterid is None else \
        np.array(clusterid, dtype=np.int64, copy=True)
    random = check_random_state(random_state)
    err = ctypes.c_double(0.0)
    ifound = ctypes.c_int64(0)
    if n_pass >= 2:
        starts = _random_starts(int(n_clusters), n_elements, int(n_pass), random)
        _lib.call("msmb200_kmedoids_restarts", int(n_clusters), n_elements, dev.ptr(dm),
                  int(n_pass), dev.ptr(starts), dev.ptr(cid), ctypes.byref(err),
                  ctypes.byref(ifound))
    else:
        if n_pass == 1:           # in place on the random start (kmedoids.cc:165-166)
            cid = _random_starts(int(n_clusters), n_elements, 1, random)[0].copy()
        _lib.call("msmb200_kmedoids", int(n_clusters), n_elements, dev.ptr(dm), dev.ptr(cid),
                  ctypes.byref(err), ctypes.byref(ifound))
    return cid.astype(np.intp, copy=False), err.value, int(ifound.value)


def contigify_ids(ids):
    """_kmedoids.contigify_ids (cluster/_kmedoids.pyx:110-117)."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    keys = np.zeros(max(len(ids), 1), dtype=np.int64)
    n_keys = ctypes.c_int64(0)
    _lib.call("msmb200_contigify_ids", dev.ptr(ids), len(ids), dev.ptr(keys),
              ctypes.byref(n_keys))
    return ids.astype(np.intp, copy=False), {int(keys[r]): r for r in range(n_keys.value)}
