"""Host stand-in for msmbuilder_b200._kernels.LookaheadState (test infrastructure).

It models exactly the information flow of csrc/kcenters_lookahead.cu -- per-lane top three of the
running minima (the two largest are entries, the third bounds the rest of the lane), candidate cut at
the T-th largest entry, the bound tau on every frame that is not a candidate, the certified chain --
with NumPy arithmetic through the oracle, so the *logic* of the look-ahead (what may be certified, tie
handling, degenerate inputs) is pinned to the reference loop on the CPU, independently of the CUDA
kernels (tests/test_gpu_lookahead.py pins those)."""
import numpy as np
import torch


class HostLookahead(object):
    """Host stand-in for _kernels.LookaheadState (same five methods, NumPy arithmetic through the
    oracle): rows are dealt to `n_lanes` strided lanes that keep their three largest running minima
    (first row first among equals), exactly the information the CUDA pass leaves behind."""

    def __init__(self, Xl, row_offset, n_lanes=7, t_cap=5, j_cap=4):
        from oracle import libdistance_oracle as lo
        self.lo = lo
        self.X, self.row_offset = Xl, row_offset
        self.n, self.d = Xl.shape
        self.n_lanes, self.t_cap, self.j_cap = n_lanes, t_cap, j_cap
        self.distances = torch.full((self.n,), float("inf"), dtype=torch.float64)
        self.labels = torch.zeros((self.n,), dtype=torch.int32)
        # centres blob: [n, ids[j_cap], rows[j_cap * d]] as float64 (ids are exact in a double here)
        self.centers = torch.zeros(1 + j_cap + j_cap * self.d, dtype=torch.float64)
        # set blob: [count, tau, val[t_cap], idx[t_cap], rows[t_cap * d]]
        self.set_len = 2 + 2 * t_cap + t_cap * self.d

    def centers_ids(self):
        return self.centers[1:1 + self.j_cap].to(torch.int64)

    def centers_rows(self):
        return self.centers[1 + self.j_cap:].reshape(self.j_cap, self.d).to(torch.float32)

    def seed(self, global_row):
        self.centers.zero_()
        local = global_row - self.row_offset
        if 0 <= local < self.n:
            self.centers[0] = 1
            self.centers[1] = global_row
            self.centers[1 + self.j_cap:1 + self.j_cap + self.d] = torch.from_numpy(self.X[local].astype(np.float64))

    def multi_pass(self, n_centers, label0, first):
        dist_np, lab_np = self.distances.numpy(), self.labels.numpy()
        rows = self.centers_rows().numpy()
        for j in range(n_centers):
            if self.n == 0:
                break
            dv = self.lo.dist(self.X, rows[j], "euclidean")
            m = dv < dist_np
            dist_np[m] = dv[m]
            lab_np[m] = label0 + j

    def select(self):
        out = torch.zeros(self.set_len, dtype=torch.float64)
        dist_np = self.distances.numpy()
        entries, thirds, tops = [], [-np.inf], []
        for lane in range(self.n_lanes):
            rows = np.arange(lane, self.n, self.n_lanes)
            if len(rows) == 0:
                continue
            v = dist_np[rows]
            order = sorted(range(len(rows)), key=lambda r: (-v[r], r))     # strict '>' updates: first row first
            tops.append((v[order[0]], self.row_offset + int(rows[order[0]])))
            for r in order[:2]:
                entries.append((v[r], self.row_offset + int(rows[r])))
            if len(rows) > 2:
                thirds.append(v[order[2]])
        if not tops:
            out[1] = -np.inf
            return out
        best = min(tops, key=lambda t: (-t[0], t[1]))
        vals = sorted((t[0] for t in entries if t[0] > 0), reverse=True)
        cut = vals[self.t_cap - 1] if len(vals) >= self.t_cap else 0.0
        cands = [t for t in entries if t[0] > 0 and t[0] > cut]
        if best not in cands:
            cands.append(best)
        out[0] = len(cands)
        out[1] = max(max(thirds), cut)
        for c, (v, gi) in enumerate(cands):
            out[2 + c] = v
            out[2 + self.t_cap + c] = gi
            o = 2 + 2 * self.t_cap + c * self.d
            out[o:o + self.d] = torch.from_numpy(self.X[gi - self.row_offset].astype(np.float64))
        return out

    def chain(self, sets, n_sets, k_remaining):
        sets = sets.reshape(n_sets, self.set_len).numpy()
        vals, idx, rows = [], [], []
        tau = -np.inf
        for s in sets:
            c = int(s[0])
            tau = max(tau, s[1])
            vals += list(s[2:2 + c])
            idx += [int(i) for i in s[2 + self.t_cap:2 + self.t_cap + c]]
            rows += [s[2 + 2 * self.t_cap + i * self.d:2 + 2 * self.t_cap + (i + 1) * self.d].astype(np.float32)
                     for i in range(c)]
        vals = np.array(vals)
        rows = np.array(rows, dtype=np.float32).reshape(len(vals), self.d)
        self.centers.zero_()
        steps = 0
        while steps < min(k_remaining, self.j_cap) and len(vals):
            order = sorted(range(len(vals)), key=lambda c: (-vals[c], idx[c]))
            b = order[0]
            if steps > 0 and not vals[b] > tau:
                break
            self.centers[1 + steps] = idx[b]
            o = 1 + self.j_cap + steps * self.d
            self.centers[o:o + self.d] = torch.from_numpy(rows[b].astype(np.float64))
            steps += 1
            vals = np.minimum(vals, self.lo.dist(rows, rows[b], "euclidean"))
        self.centers[0] = steps
        return steps
