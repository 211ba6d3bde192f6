"""Streaming a directory of .npy sequences onto the GPU.

The reference keeps a featurised dataset as a directory of ``%08d.npy`` files
(``NumpyDirDataset``, msmbuilder/dataset.py:290-334) and feeds it to
``estimator.fit`` as a lazy iterable; every clusterer then makes one more host
copy of everything (``np.concatenate``, cluster/base.py:58).  ``NumpyDirStream``
is that iterable for the device path:

  * iterating it yields CUDA tensors; a reader thread stays ``prefetch`` files
    ahead (file -> pinned staging buffer -> H2D on a side stream), so disk, PCIe
    and the kernels of the consumer (``tICA.fit`` consumes sequence by sequence)
    overlap;
  * ``to_device()`` uploads every file straight into its slot of ONE device
    allocation (shapes come from the .npy headers, nothing is concatenated on
    the host); the clusterers call it instead of their ``_concat`` copy.

No estimator state lives here; this is plumbing around ``cudaMemcpyAsync``.
"""
from __future__ import absolute_import, print_function, division

import os
import re
import threading

import numpy as np

from . import _lib

__all__ = ['NumpyDirStream', 'save_sequences']

_ITEM_FORMAT = '%08d.npy'
_ITEM_RE = re.compile(r'(\d{8})\.npy$')


def save_sequences(path, sequences):
    """Write sequences as ``path/%08d.npy`` (the layout of dataset.py:305-325)."""
    os.makedirs(path, exist_ok=True)
    for i, x in enumerate(sequences):
        np.save(os.path.join(path, _ITEM_FORMAT % i), np.asarray(x))
    return path


class _Staging(object):
    """`count` pinned host buffers handed out round-robin; a buffer is reused only
    after the copy that last read it has finished (its event)."""

    def __init__(self, count):
        self.bufs = [None] * count
        self.events = [None] * count
        self.i = 0

    def take(self, nbytes):
        import torch
        j = self.i
        self.i = (self.i + 1) % len(self.bufs)
        if self.events[j] is not None:
            self.events[j].synchronize()
        if self.bufs[j] is None or self.bufs[j].numel() < nbytes:
            self.bufs[j] = torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()
        return j, self.bufs[j]

    def drain(self):
        """Wait until no copy reads any buffer (they are about to be released: the
        pinned allocator does not know about copies issued from NumPy views)."""
        for ev in self.events:
            if ev is not None:
                ev.synchronize()


class NumpyDirStream(object):
    """Lazy, re-iterable collection of sequences stored as .npy files.

    Parameters
    ----------
    source : str or list of str
        A directory holding ``%08d.npy`` files (read in numeric order), or an
        explicit list of .npy paths.
    prefetch : int
        How many sequences the reader thread may be ahead of the consumer.
    """

    def __init__(self, source, prefetch=2):
        if isinstance(source, (list, tuple)):
            self.files = [str(f) for f in source]
        else:
            root = os.path.expanduser(str(source))
            names = sorted((int(m.group(1)), fn) for fn in os.listdir(root)
                           for m in [_ITEM_RE.match(fn)] if m)
            self.files = [os.path.join(root, fn) for _, fn in names]
        if not self.files:
            raise ValueError('no .npy sequences in %r' % (source,))
        self.prefetch = max(int(prefetch), 1)

    def __len__(self):
        return len(self.files)

    def keys(self):
        return list(range(len(self.files)))

    def _open(self, i):
        a = np.load(self.files[i], mmap_mode='r')
        if a.dtype not in (np.float32, np.float64):
            a = np.asarray(a, dtype=np.float64)       # tica.py:402 / kcenters.py:80-82
        return a

    def shapes(self):
        """(shape, dtype) of every sequence from the file headers only."""
        out = []
        for i in range(len(self.files)):
            a = self._open(i)
            out.append((tuple(a.shape), a.dtype))
        return out

    # ------------------------------------------------------------------ streaming
    def _upload(self, a, staging, copy_stream, dst=None):
        """file/array -> pinned -> device (async on copy_stream); returns (tensor, event)."""
        import torch
        a = np.ascontiguousarray(a) if not a.flags['C_CONTIGUOUS'] else a
        nbytes = a.nbytes
        j, pin = staging.take(nbytes)
        host = pin[:nbytes].numpy().view(a.dtype).reshape(a.shape)
        np.copyto(host, a)                             # page cache / disk -> pinned
        with torch.cuda.stream(copy_stream):
            if dst is None:
                dst = torch.empty(a.shape, dtype=torch.from_numpy(host[:0]).dtype, device="cuda")
            dst.copy_(torch.from_numpy(host), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staging.events[j] = ev
        return dst, ev

    def __iter__(self):
        import queue
        import torch
        _lib.require_gpu()
        device = torch.cuda.current_device()
        copy_stream = torch.cuda.Stream()
        staging = _Staging(self.prefetch + 1)
        q = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def reader():
            try:
                torch.cuda.set_device(device)
                for i in range(len(self.files)):
                    if stop.is_set():
                        return
                    q.put(self._upload(self._open(i), staging, copy_stream))
                q.put(None)
            except BaseException as e:           # surfaced in the consumer
                q.put(e)

        th = threading.Thread(target=reader, name="msmb200-npy-reader", daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                t, ev = item
                cur = torch.cuda.current_stream()
                cur.wait_event(ev)
                t.record_stream(cur)
                yield t
        finally:
            stop.set()
            while th.is_alive():                 # unblock a reader stuck on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                th.join(timeout=0.01)
            staging.drain()

    # ------------------------------------------------------------------ one allocation
    def to_device(self):
        """All sequences as CUDA tensors lying back to back in one device buffer
        (what ``_device.FrameStore`` adopts without a copy)."""
        import torch
        _lib.require_gpu()
        shapes = self.shapes()
        inner = shapes[0][0][1:]
        for s, _ in shapes:
            if s[1:] != inner:
                raise ValueError('all sequences must have the same trailing shape')
        dtype = torch.float64 if any(dt == np.float64 for _, dt in shapes) else torch.float32
        np_dtype = np.float64 if dtype == torch.float64 else np.float32
        lengths = [s[0] for s, _ in shapes]
        offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
        data = torch.empty((int(offsets[-1]),) + tuple(inner), dtype=dtype, device="cuda")
        copy_stream = torch.cuda.Stream()
        copy_stream.wait_stream(torch.cuda.current_stream())   # `data` may reuse a block still in use
        staging = _Staging(self.prefetch + 1)
        last = None
        for i, n in enumerate(lengths):
            if n == 0:
                continue
            a = self._open(i)
            if a.dtype != np_dtype:
                a = np.asarray(a, dtype=np_dtype)
            _, last = self._upload(a, staging, copy_stream, dst=data[int(offsets[i]):int(offsets[i + 1])])
        if last is not None:
            torch.cuda.current_stream().wait_event(last)
        staging.drain()
        return [data[int(o):int(o) + n] for o, n in zip(offsets[:-1], lengths)]
