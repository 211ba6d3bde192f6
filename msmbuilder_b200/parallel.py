"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink).

The hot path shards by frames:

* tICA: whole sequences are dealt to ranks (longest-processing-time first); each
  rank accumulates its packed float64 statistics on its GPU and ONE
  ``all_reduce(SUM)`` of 3*D*D + 3*D + 2 doubles (1.58 MB at D = 256) merges them.
  The sufficient statistics are additive (tica.py:414-422 are all ``+=``), and
  n_observations / n_sequences ride in the same buffer.
* KCenters: the concatenated frames are sharded contiguously.  Look-ahead path
  (float32, (sq)euclidean): after each fused pass every rank all-gathers its
  candidate set (<= 512 frames + the bound tau on everything else, ~0.5 MB) and
  runs the same deterministic chain kernel on the union, so all ranks know the next
  J centres without a broadcast: ONE ``all_gather`` per chain, not per centre.
  Other metrics: every pass each rank writes its farthest frame {value, global
  index, row} into a candidate slot, ONE ``all_gather`` of those slots follows, and
  every rank deterministically selects (max value, then lowest global index ==
  np.argmax's first maximum, kcenters.py:97).
* assign_nearest: centres are broadcast once; labels stay with their shard.

The collectives are tiny and latency-bound, so they are plain NCCL calls on
device buffers.  The same code runs under the ``gloo`` backend with CPU tensors
when the compute callbacks are swapped for host ones (that is how the
world_size-2 CPU tests exercise the protocol).
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None

CAND_HEADER = 16   # sizeof(msmb200_candidate): double value, int64 index


def world():
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ----------------------------------------------------------------------------- communicators
# The k-centers loops need two collectives on small device blobs: an all-gather of equally sized
# byte blobs and a SUM all-reduce.  Two back ends provide them:
#   DistComm    one process per GPU under torchrun (NCCL over NVLink; gloo on CPU in the tests)
#   ThreadComm  one PROCESS, one Python thread per GPU (`devices=` of the estimators): the blobs are
#               exchanged with peer copies between two thread barriers.  This is what lets a stock
#               Pipeline / fit call use every GPU of the box without torchrun.
class DistComm(object):
    def __init__(self, group=None):
        self.group = group
        self.rank, self.ws = world()

    def all_gather_into(self, out, local):
        if self.ws == 1:
            out.copy_(local.reshape(-1))
            return
        dist.all_gather_into_tensor(out, local, group=self.group)

    def all_reduce_sum(self, t):
        if self.ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)


class ThreadGroup(object):
    """Shared state of `n` cooperating threads (one per device)."""

    def __init__(self, n):
        import threading
        self.n = int(n)
        self.barrier = threading.Barrier(self.n)
        self.slots = [None] * self.n

    def comm(self, rank):
        return ThreadComm(self, rank)

    def abort(self):
        self.barrier.abort()


def _sync_stream(t):
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()


class ThreadComm(object):
    def __init__(self, group, rank):
        self.g = group
        self.rank = int(rank)
        self.ws = group.n

    def all_gather_into(self, out, local):
        g = self.g
        if g.n == 1:
            out.copy_(local.reshape(-1))
            return
        _sync_stream(local)                        # my blob is complete before anyone reads it
        g.slots[self.rank] = local
        g.barrier.wait()
        n = local.numel()
        for r in range(g.n):
            out[r * n:(r + 1) * n].copy_(g.slots[r].reshape(-1), non_blocking=True)
        _sync_stream(out)                          # every blob has been read ...
        g.barrier.wait()                           # ... before its owner may overwrite it

    def all_reduce_sum(self, t):
        g = self.g
        if g.n == 1:
            return
        _sync_stream(t)
        g.slots[self.rank] = t
        g.barrier.wait()
        total = g.slots[0].to(t.device, copy=True)
        for r in range(1, g.n):                    # fixed order: every thread gets the same bits
            total += g.slots[r].to(t.device)
        _sync_stream(total)
        g.barrier.wait()
        t.copy_(total)


def _agree_max(values, group=None):
    """Element-wise MAX of a few Python ints over the ranks (device tensor under NCCL, CPU tensor
    under gloo); every rank returns the same list."""
    _, ws = world()
    vals = [int(v) for v in values]
    if ws == 1:
        return vals
    on_gpu = dist.get_backend(group) == "nccl"
    t = torch.tensor(vals, dtype=torch.int64, device="cuda" if on_gpu else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [int(v) for v in t.cpu().tolist()]


def shard_sequences(lengths, world_size):
    """Longest-processing-time assignment of whole sequences to ranks.
    Returns a list (per rank) of sequence indices, each in ascending order."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    owned = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        owned[r].append(i)
        loads[r] += int(lengths[i])
    return [sorted(o) for o in owned]


def shard_rows(n_total, world_size):
    """Contiguous, near-equal row ranges [(start, stop)] per rank."""
    base, extra = divmod(int(n_total), world_size)
    out, start = [], 0
    for r in range(world_size):
        stop = start + base + (1 if r < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def allreduce_packed(packed, group=None):
    """In-place SUM all-reduce of a packed float64 tICA accumulator (device tensor
    under NCCL, CPU tensor under gloo)."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def select_candidate_host(gathered, n_cand, cand_bytes):
    """Host mirror of msmb200_candidate_select for CPU (gloo) runs: gathered is a
    uint8 tensor of n_cand slots; returns the winning slot index."""
    buf = gathered.numpy().reshape(n_cand, cand_bytes)
    values = buf[:, :8].copy().view(np.float64).reshape(-1)
    idx = buf[:, 8:16].copy().view(np.int64).reshape(-1)
    win = 0
    for r in range(1, n_cand):
        if values[r] > values[win] or (values[r] == values[win] and idx[r] < idx[win]):
            win = r
    return win


def kcenters_fit_distributed(n_clusters, cand_bytes, seed_fn, pass_fn, select_fn, alloc_fn,
                             group=None, pass_ring=False, comm=None):
    """Rank-collective Gonzalez loop.

    seed_fn(cand)                      fill `cand` with the seed centre on the rank
                                       that owns it (value must be +inf there) and
                                       with value = -inf elsewhere
    pass_fn(center_cand, label, out)   one local pass; writes the local farthest
                                       frame into `out`
    select_fn(gathered, n, out)        choose the winner among n gathered slots
    alloc_fn(n_bytes)                  zeroed uint8 buffer on the compute device
    pass_ring                          also hand the ring of chosen centres to pass_fn
                                       (pass_fn(center_cand, label, out, ring): the pruned RMSD pass)

    Returns the (k, cand_bytes) ring of chosen centres (slot i = centre i).
    """
    if comm is None:
        comm = DistComm(group)
    ws = comm.ws
    k = int(n_clusters)
    ring = alloc_fn((k + 1) * cand_bytes).reshape(k + 1, cand_bytes)
    local = alloc_fn(cand_bytes)
    gathered = alloc_fn(ws * cand_bytes)

    def exchange(dst):
        if ws == 1:
            dst.copy_(local)
            return
        comm.all_gather_into(gathered, local)
        select_fn(gathered, ws, dst)

    seed_fn(local)
    exchange(ring[0])
    for i in range(k):
        if pass_ring:
            pass_fn(ring[i], i, local, ring)
        else:
            pass_fn(ring[i], i, local)
        exchange(ring[i + 1])
    return ring


def lookahead_collectives(group=None, comm=None):
    """The two collectives of the look-ahead k-centers (_kernels.kcenters_fit_lookahead):

    gather_sets(local_set) -> (all_sets, world_size)   ONE all-gather per chain of every rank's
                                                       candidate set (uint8 blob, same size everywhere)
    bcast(blob)                                        the seed centre: only the rank that holds
                                                       the seed row wrote a non-zero blob, so a SUM
                                                       all-reduce of the bytes hands it to everyone

    Works on whatever device the blobs live on (NCCL: cuda, gloo: cpu)."""
    if comm is None:
        comm = DistComm(group)
    ws = comm.ws
    gathered = {}

    def gather_sets(local_set):
        if ws == 1:
            return local_set, 1
        if "buf" not in gathered:
            gathered["buf"] = torch.empty(ws * local_set.numel(), dtype=local_set.dtype,
                                          device=local_set.device)
        comm.all_gather_into(gathered["buf"], local_set)
        return gathered["buf"], ws

    def bcast(blob):
        if ws > 1:
            comm.all_reduce_sum(blob)

    return gather_sets, bcast


def broadcast_centers(centers, src=0, group=None):
    _, ws = world()
    if ws > 1:
        dist.broadcast(centers, src=src, group=group)
    return centers


def kcenters_fit_gpu(data_local, row_offset, n_clusters, metric, seed_global, traces=None,
                     group=None, lookahead=True, stats=None, comm=None):
    """KCenters over frame shards: `data_local` is this rank's contiguous slice of
    the concatenated frames, starting at global row `row_offset`.

    Returns (cluster_ids int64[k] (global indices, device), distances f64[n_local],
    labels i32[n_local], centres ring (k+1, cand_bytes) uint8)."""
    from . import _kernels as K
    if comm is None:
        comm = DistComm(group)
    # every rank must take the same path: the shape test is a function of (d, dtype, metric) and of
    # the local base address alignment, which FrameStore / torch allocations always satisfy
    if traces is None and lookahead and K.lookahead_supported(data_local, metric):
        gather_sets, bcast = lookahead_collectives(comm=comm)
        # the chain kernel works on the UNION of the ranks' candidate sets (one block; its cost grows with
        # the union): every rank keeps its share of the single-GPU cap, the union stays ~1024 candidates and
        # tau (the largest per-rank cut) stays near the 1024th largest value overall
        t_cap = max(64, K.LOOKAHEAD_T_CAP // max(1, comm.ws))
        state = K.LookaheadState(data_local, metric, row_offset=row_offset, t_cap=t_cap)
        ids, rows, distances, labels = K.kcenters_fit_lookahead(
            data_local, n_clusters, metric, seed_global, gather_sets=gather_sets, bcast=bcast,
            row_offset=row_offset, stats=stats, state=state)
        k = int(n_clusters)
        cand_bytes = CAND_HEADER + 4 * int(data_local.shape[1])
        ring = torch.zeros((k + 1, cand_bytes), dtype=torch.uint8, device="cuda")
        ring[:k, 8:16] = ids.view(torch.uint8).reshape(k, 8)
        ring[:k, CAND_HEADER:] = rows.contiguous().view(torch.uint8).reshape(k, -1)
        return ids, distances, labels, ring
    if stats is not None:
        stats["passes"] = int(n_clusters)
    st = K.KCentersState(data_local, metric, traces=traces, row_offset=row_offset,
                         max_centres=int(n_clusters))

    def alloc(nbytes):
        return torch.zeros(int(nbytes), dtype=torch.uint8, device="cuda")

    def seed_fn(cand):
        local = int(seed_global) - int(row_offset)
        if 0 <= local < st.n:
            st.seed(cand, local)
        else:
            cand.zero_()
            cand[:8].view(torch.float64)[0] = float("-inf")

    def pass_fn(center_cand, label, out, ring):
        st.run_pass(center_cand, label, out_cand=out, ring=ring)

    ring = kcenters_fit_distributed(n_clusters, st.cand_bytes, seed_fn, pass_fn, st.select,
                                    alloc, comm=comm, pass_ring=True)
    k = int(n_clusters)
    ids = ring[:k, 8:16].contiguous().view(torch.int64).reshape(k)
    return ids, st.distances, st.labels, ring


# ----------------------------------------------------------------- estimator level
def tica_fit_sharded(est, local_sequences, group=None):
    """``tICA.fit`` under torchrun: every rank passes ITS sequences; after one
    all-reduce every rank holds the same fitted estimator (tica.py:261-290)."""
    import warnings
    est._initialized = False
    local_sequences = list(local_sequences)
    # a rank may hold no sequence at all (fewer sequences than ranks): the width is agreed on
    # first, every rank initialises, and an empty rank contributes a zero accumulator
    widths = {int(X.shape[1]) for X in local_sequences}
    if len(widths) > 1:
        raise ValueError("sequences have different numbers of features: %s" % sorted(widths))
    D = _agree_max([max(widths) if widths else 0], group=group)[0]
    if D == 0:
        raise ValueError("no sequences on any rank")
    if widths and max(widths) != D:
        raise ValueError("sequence has %d features, other ranks have %d" % (max(widths), D))
    est._initialize(D)
    seqs = []
    for X in local_sequences:
        if not int(X.shape[0]) > est.lag_time:
            warnings.warn("length of data (%d) is too short for the lag time (%d)"
                          % (int(X.shape[0]), est.lag_time))
            continue
        seqs.append(X)
    lib_len = 3 * D * D + 3 * D + 2
    if seqs:
        acc = est._accumulate_device(seqs)
    else:
        acc = torch.zeros(lib_len, dtype=torch.float64, device="cuda")
    allreduce_packed(acc, group=group)
    est._add_packed(acc.cpu().numpy())
    if est.n_sequences_ == 0:
        raise ValueError('All sequences were shorter than the lag time, %d' % est.lag_time)
    return est


def kcenters_fit_sharded(est, local_sequences, row_offset, n_total, group=None):
    """``KCenters.fit`` under torchrun: `local_sequences` are this rank's slice of
    the global frame order starting at global row `row_offset`.  Sets the usual
    fitted attributes; labels_/distances_ cover the local sequences only."""
    from sklearn.utils import check_random_state
    from ._device import FrameStore
    local_sequences = list(local_sequences)
    if not local_sequences:
        # an empty shard still takes part in every collective: an (0, ...) block of the agreed shape
        raise ValueError("kcenters_fit_sharded: this rank holds no sequence; pass an empty "
                         "(0, n_features) array so the frame width is known")
    store = FrameStore(local_sequences)
    seed = check_random_state(est.random_state).randint(0, int(n_total))   # same draw on every rank
    traces = None
    data = store.data
    if est.metric == 'rmsd':
        from . import _kernels as K
        data = data.clone()
        traces = K.rmsd_center(data)
    ids, distances, labels, ring = kcenters_fit_gpu(data, row_offset, est.n_clusters, est.metric,
                                                    seed, traces=traces, group=group)
    k = int(est.n_clusters)
    est.cluster_ids_ = [int(c) for c in ids.cpu().numpy()]
    row_elems = int(np.prod(data.shape[1:]))       # not data[0]: the shard may hold no frame
    es = data.element_size()
    cent = ring[:k, CAND_HEADER:CAND_HEADER + row_elems * es].contiguous().view(data.dtype)
    est.cluster_centers_ = cent.reshape((k,) + tuple(data.shape[1:])).cpu().numpy()
    from ._device import to_host
    est.labels_ = store.split(to_host(labels, torch.int64))
    est.distances_ = store.split(to_host(distances))
    local_sum = distances.sum().reshape(1)
    _, ws = world()
    if ws > 1:
        dist.all_reduce(local_sum, op=dist.ReduceOp.SUM, group=group)
    est.inertia_ = float(local_sum.item())
    return est
