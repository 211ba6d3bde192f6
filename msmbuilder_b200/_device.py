"""Device-memory plumbing (PyTorch is used for allocation, streams and copies only)."""
import ctypes

import os

import numpy as np
import torch

from . import _lib
from .utils import is_tensor


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or host) address of a tensor / ndarray, or NULL."""
    if t is None:
        return ctypes.c_void_p(0)
    if is_tensor(t):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def dtype_id(t):
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float64:
        return _lib.F64
    raise TypeError("frames must be float32 or float64 on the device, got %s" % t.dtype)


def _float_dtype(dt):
    """Reference policy (kcenters.py:80-82, tica.py:402 modulo storage width):
    float32 and float64 are kept, everything else is widened to float64."""
    if is_tensor(dt):
        dt = dt.dtype
    if dt in (torch.float32, np.dtype(np.float32), np.float32):
        return torch.float32
    return torch.float64


def to_device(X, ndim=2):
    """One array-like -> contiguous CUDA tensor (float32 or float64)."""
    _lib.require_gpu()
    if is_tensor(X):
        want = _float_dtype(X.dtype)
        t = X if X.dtype == want else X.to(want)
        if not t.is_cuda:
            t = t.cuda(non_blocking=True)
        t = t.contiguous()
    else:
        a = np.asarray(X)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a).cuda()
    if ndim == 2 and t.ndim == 1:
        t = t.unsqueeze(0)
    return t


class HostUploader(object):
    """Pageable host arrays -> device memory at (close to) the PCIe rate.

    ``tensor.cuda()`` on pageable memory is one driver thread copying through a small staging
    buffer (a few GB/s).  Here a few Python threads fill a ring of PINNED chunks with
    ``np.copyto`` (the GIL is released for the copy) while earlier chunks are already on their
    way with ``cudaMemcpyAsync`` on a side stream; the compute stream only waits for the last
    chunk.  This is what a Pipeline caller holds: plain NumPy arrays (cluster/base.py:55-58 and
    tica.py:401-403 both start from them)."""

    CHUNK = 32 << 20
    SLOTS = 12
    # one host thread copies ~10 GB/s into the pinned ring: 4 threads capped the upload at 41 GB/s of
    # the box's 55 GB/s pinned H2D rate (r2o bench e2e); leave a few cores to the caller
    THREADS = max(4, min(8, (os.cpu_count() or 8) - 4))

    def __init__(self):
        self._pinned = None
        self._pool = None
        self._stream = {}

    def _ring(self):
        if self._pinned is None:
            self._pinned = torch.empty(self.SLOTS * self.CHUNK, dtype=torch.uint8).pin_memory()
        return self._pinned

    def upload(self, pairs):
        """pairs: [(C-contiguous ndarray, CUDA tensor of the same byte size)].  Asynchronous with
        respect to the host only in its tail: returns when every chunk has been ISSUED; the current
        stream is made to wait for the copies."""
        import threading
        from concurrent.futures import ThreadPoolExecutor
        jobs = []
        for src, dst in pairs:
            if src.nbytes == 0:
                continue
            sb = src.reshape(-1).view(np.uint8)
            db = dst.reshape(-1).view(torch.uint8)
            if sb.nbytes != db.numel():
                raise ValueError("upload: size mismatch")
            for o in range(0, sb.nbytes, self.CHUNK):
                jobs.append((sb[o:o + self.CHUNK], db[o:o + self.CHUNK]))
        if not jobs:
            return
        cur = torch.cuda.current_stream()
        devi = torch.cuda.current_device()
        if devi not in self._stream:
            self._stream[devi] = torch.cuda.Stream()
        cs = self._stream[devi]
        cs.wait_stream(cur)                      # the destination blocks may still be in use there
        if len(jobs) == 1 and jobs[0][0].nbytes < (1 << 20):
            jobs[0][1].copy_(torch.from_numpy(jobs[0][0]), non_blocking=False)
            return
        ring = self._ring()
        ring_np = ring.numpy()
        if self._pool is None:
            self._pool = ThreadPoolExecutor(self.THREADS, thread_name_prefix="msmb200-h2d")
        issued = [threading.Event() for _ in jobs]
        events = [None] * len(jobs)

        def fill(k):
            if k >= self.SLOTS:                  # the slot's previous chunk must have left the host
                issued[k - self.SLOTS].wait()
                events[k - self.SLOTS].synchronize()
            slot = (k % self.SLOTS) * self.CHUNK
            n = jobs[k][0].nbytes
            np.copyto(ring_np[slot:slot + n], jobs[k][0])
            return slot, n

        futs = [self._pool.submit(fill, k) for k in range(len(jobs))]
        try:
            with torch.cuda.stream(cs):
                for k, f in enumerate(futs):
                    slot, n = f.result()
                    jobs[k][1].copy_(ring[slot:slot + n], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    events[k] = ev
                    issued[k].set()
        finally:
            for k in range(len(jobs)):           # never leave a worker waiting on a failed upload
                if events[k] is None:
                    events[k] = torch.cuda.Event()
                    events[k].record(cs)
                issued[k].set()
        cur.wait_stream(cs)
        # the ring is reused by the next call: its chunks must have been read by then
        events[-1].synchronize()


_UPLOADERS = {}


def uploader():
    """One uploader (pinned ring, copy threads, side stream) per device: the worker threads of a
    multi-GPU fit (`devices=`) upload to their own GPU at the same time."""
    d = torch.cuda.current_device()
    up = _UPLOADERS.get(d)
    if up is None:
        up = _UPLOADERS.setdefault(d, HostUploader())
    return up


def resolve_devices(devices):
    """`devices` of the estimators -> list of CUDA ordinals, or None for "the current device".
    None / a single device keep the one-GPU path; 'all' = every visible GPU."""
    if devices is None:
        return None
    if isinstance(devices, str):
        if devices != 'all':
            raise ValueError("devices must be None, 'all' or a sequence of CUDA ordinals")
        devs = list(range(torch.cuda.device_count()))
    else:
        devs = [int(d) for d in devices]
    if len(set(devs)) != len(devs) or any(d < 0 or d >= torch.cuda.device_count() for d in devs):
        raise ValueError("devices=%r: need distinct ordinals below torch.cuda.device_count()=%d"
                         % (devices, torch.cuda.device_count()))
    return devs if len(devs) > 1 else None


def host_array(a):
    """ndarray in the dtype the device path keeps it in (float32 / float64, C-contiguous)."""
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


class FrameStore(object):
    """The concatenation of a list of sequences as ONE (N, ...) device tensor.

    Replaces MultiSequenceClusterMixin._concat's host np.concatenate
    (msmbuilder/cluster/base.py:55-58): host arrays are copied straight into
    their slot of the device buffer; device tensors that already sit back to
    back in one allocation are adopted without a copy.
    """

    def __init__(self, sequences):
        _lib.require_gpu()
        seqs = list(sequences)
        if len(seqs) == 0:
            raise ValueError("no sequences")
        self.lengths = [int(s.shape[0]) for s in seqs]
        inner = tuple(seqs[0].shape[1:])
        for s in seqs:
            if tuple(s.shape[1:]) != inner:
                raise ValueError("all sequences must have the same trailing shape")
        dts = {_float_dtype(s.dtype) for s in seqs}
        dtype = torch.float64 if torch.float64 in dts else torch.float32
        self.offsets = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)
        n_total = int(self.offsets[-1])
        row = int(np.prod(inner)) if inner else 1

        adopted = None
        if all(is_tensor(s) and s.is_cuda and s.dtype == dtype and s.is_contiguous()
               for s in seqs) and n_total > 0:
            base = seqs[0].data_ptr()
            es = seqs[0].element_size()
            if all(s.data_ptr() == base + int(o) * row * es
                   for s, o in zip(seqs, self.offsets[:-1])):
                try:
                    adopted = seqs[0].as_strided((n_total,) + inner,
                                                 torch.empty((n_total,) + inner, device="meta").stride())
                except RuntimeError:
                    adopted = None
        if adopted is not None:
            self.data = adopted
        else:
            self.data = torch.empty((n_total,) + inner, dtype=dtype, device="cuda")
            np_dtype = np.float64 if dtype == torch.float64 else np.float32
            host_pairs = []
            for s, o, n in zip(seqs, self.offsets[:-1], self.lengths):
                if n == 0:
                    continue
                dst = self.data[int(o):int(o) + n]
                if is_tensor(s):
                    dst.copy_(s, non_blocking=True)
                else:
                    host_pairs.append((host_array(s).astype(np_dtype, copy=False), dst))
            uploader().upload(host_pairs)        # pageable arrays: pinned ring + copy threads
        self.n = n_total
        self.inner = inner

    def split(self, concat):
        """cluster/base.py:76-77 on host arrays or device tensors."""
        return [concat[int(o): int(o) + n] for o, n in zip(self.offsets[:-1], self.lengths)]


def device_sequences(sequences):
    """Upload a list of host sequences ONCE: returns CUDA tensors that sit back to
    back in one allocation, so tICA.fit / transform and every clusterer's
    ``_concat`` (cluster/base.py:55-58 in the reference) use them without a
    further copy.  Host arrays are copied slot by slot (pinned memory goes at
    PCIe rate); CUDA tensors are adopted when already contiguous."""
    store = FrameStore(sequences)
    return store.split(store.data)


PINNED_RESULT_LIMIT = 256 << 20


def to_host(t, dtype=None):
    """Device tensor -> NumPy array through PINNED host memory (torch's caching host allocator
    keeps the blocks, so repeated calls pay no cudaHostAlloc): a pageable ``.cpu()`` of the
    labels / distances of a few million frames runs at ~2 GB/s and was 20 % of an end-to-end
    KCenters.fit; this runs at PCIe rate.  `dtype`: convert ON THE DEVICE first (e.g. the
    int32 labels to the reference's intp) instead of a second pass over the host copy.
    The returned array owns its (pinned) storage."""
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    t = t.contiguous()
    if t.numel() * t.element_size() < (1 << 20) or not t.is_cuda:
        return t.cpu().numpy()
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    arr = out.numpy()
    # labels_ / distances_ live as long as the estimator: above PINNED_RESULT_LIMIT bytes they move on to
    # pageable memory (one host memcpy) so that a fit on hundreds of millions of frames does not keep
    # gigabytes page-locked; smaller results keep their pinned storage (no second pass over them)
    if arr.nbytes > PINNED_RESULT_LIMIT:
        arr = arr.copy()
    return arr


class Workspace(object):
    """Grow-only device scratch buffers keyed by (purpose, device, stream): two calls on
    different CUDA streams or devices never share scratch (they would race on it)."""

    def __init__(self):
        self._bufs = {}

    def get(self, key, nbytes):
        nbytes = max(int(nbytes), 256)
        key = (key, torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
            self._bufs[key] = buf
        return buf


_WS = None


def workspace():
    global _WS
    if _WS is None:
        _WS = Workspace()
    return _WS
