"""A CPU stand-in for ``heat_b200.engine.CudaEngine`` built on the oracle.  TEST ONLY: it lets the gloo
world_size-2 tests drive the *host* logic of KMeans.fit / predict (sharding, allreduce, sticky
convergence flag, n_iter bookkeeping) without a GPU.  The product never imports this module."""
import numpy as np
import torch

from oracle import kmeans_oracle as orc


class CheckerEngine:
    def __init__(self, device):
        self.device = device
        self.comm = None
        self.steps = 0

    def init_comm(self, comm):
        self.comm = comm

    def row_workspace(self, n_local):
        return None

    def allreduce_f64(self, buf):
        if self.comm is not None and self.comm.is_distributed():
            from heat_b200.communication import IN_PLACE

            self.comm.Allreduce(IN_PLACE, buf)

    def _partials(self, x, c):
        k, d = c.shape
        lab = orc.assign_to_cluster(x, c).view(-1) if x.shape[0] else torch.zeros(0, dtype=torch.int64)
        part = torch.zeros(k, d + 1, dtype=torch.float64)
        if x.shape[0]:
            part[:, :d].index_add_(0, lab, x.double())
            part[:, d] += torch.bincount(lab, minlength=k).double()
        return part.view(-1)

    def lloyd_accumulate(self, x, c, partials, labels=None, path="auto", row_ws=None):
        partials.copy_(self._partials(x, c))

    def lloyd_finalize(self, partials, c_in, c_out, use_tol, tol_cmp, shift2, state):
        if int(state[0]):
            return
        k, d = c_in.shape
        p = partials.view(k, d + 1)
        div = p[:, d].clamp(min=1).to(torch.float32).to(torch.float64)
        new = (p[:, :d] / div.view(-1, 1)).to(c_in.dtype)
        s = ((c_in - new) ** 2).sum()
        c_out.copy_(new)
        shift2.copy_(s)
        state[1] += 1
        if use_tol and bool(s <= torch.tensor(tol_cmp, dtype=torch.float32).to(s.dtype)):
            state[0] = 1

    def lloyd_run(self, x, c, c_prev, use_tol, tol_cmp, shift2, state, allreduce, iters, path="auto", row_ws=None):
        for _ in range(iters):
            self.lloyd_step(x, c, c_prev, use_tol, tol_cmp, shift2, state, allreduce)

    def lloyd_step(self, x, c, c_prev, use_tol, tol_cmp, shift2, state, allreduce, labels=None, path="auto",
                   row_ws=None):
        self.steps += 1
        if int(state[0]):
            return
        part = self._partials(x, c)
        if allreduce:
            self.allreduce_f64(part)
        c_prev.copy_(c)
        self.lloyd_finalize(part, c.clone(), c, use_tol, tol_cmp, shift2, state)

    def cdist(self, x, y, out, quadratic_expansion, sqrt=True):
        out.copy_(orc.cdist(x, y, quadratic_expansion))

    def pairwise(self, x, y, out, metric="euclidean", expand=False, sigma=1.0):
        if x.shape[0] and y.shape[0]:
            out.copy_(orc.pairwise(x, y, metric, expand, sigma))

    def assign(self, x, c, labels, fv=None, path="auto", row_ws=None):
        if x.shape[0] == 0:
            if fv is not None:
                fv.zero_()
            return
        lab, mins = orc.assign_to_cluster(x, c, eval_functional_value=True)
        labels.copy_(lab.to(labels.dtype).view(labels.shape))
        if fv is not None:
            fv[0] = float((mins.double() ** 2).sum())

    # -- other consumers (KMedians / KMedoids / kNN): numpy restatement of the device protocol of include/hkmeans.h ----
    def assign_l1(self, x, c, labels, fv=None):
        from oracle import consumers_oracle as con

        if x.shape[0] == 0:
            if fv is not None:
                fv.zero_()
            return
        lab, mins = con.assign_l1(x, c)
        labels.copy_(lab.to(labels.dtype).view(labels.shape))
        if fv is not None:
            fv[0] = float(mins.double().sum())

    def kmex_update(self, c, flag, atol, partials=None, medians=None, counts=None, rtol=1e-5):
        k, d = c.shape
        old = c.clone()
        if partials is not None:
            p = partials.view(k, d + 1)
            has = p[:, d] > 0
            new = (p[:, :d] / p[:, d].clamp(min=1).view(-1, 1)).to(c.dtype)
        else:
            has, new = counts > 0, medians
        c[has] = new[has]
        flag[0] = int(torch.allclose(c, old, atol=atol, rtol=rtol))

    def cluster_medians(self, x, labels, k, allsum=None, drop_zero_rows=True, lower=False):
        n, d = x.shape
        bits = 32 if x.dtype == torch.float32 else 64
        raw = x.contiguous().numpy().view(np.uint32 if bits == 32 else np.uint64).astype(np.uint64)
        top = np.uint64(1) << np.uint64(bits - 1)
        full = np.uint64((1 << bits) - 1)
        key = np.where(raw & top != 0, raw ^ full, raw ^ top)
        keep = (x != 0).any(dim=1).numpy() if drop_zero_rows else np.ones(n, dtype=bool)
        lab = labels.reshape(-1).numpy()
        prefix = np.zeros((2, k, d), dtype=np.uint64)
        remaining = np.zeros((2, k, d), dtype=np.int64)
        counts = None
        for p in range(bits // 8):
            shift = np.uint64(bits - 8 * (p + 1))
            hist = np.zeros((2, k, d, 256), dtype=np.int64)
            for j in range(k):
                kj = key[keep & (lab == j)]
                if kj.shape[0] == 0:
                    continue
                digit = ((kj >> shift) & np.uint64(255)).astype(np.int64)
                for w in range(2):
                    for f in range(d):
                        sel = np.ones(kj.shape[0], dtype=bool) if p == 0 else (kj[:, f] >> (shift + np.uint64(8))) == prefix[w, j, f]
                        hist[w, j, f] = np.bincount(digit[sel, f], minlength=256)
            ht = torch.from_numpy(hist)
            if allsum is not None:
                allsum(ht)
            hist = ht.numpy()
            if p == 0:
                counts = hist[0, :, 0, :].sum(axis=1)
                remaining[0] = (np.maximum(counts - 1, 0) // 2).reshape(k, 1)
                remaining[1] = (counts // 2).reshape(k, 1)
            cum = np.cumsum(hist, axis=3)
            digit = (cum > remaining[..., None]).argmax(axis=3)
            before = np.where(digit > 0, np.take_along_axis(cum, np.maximum(digit - 1, 0)[..., None], axis=3)[..., 0], 0)
            remaining = remaining - before
            prefix = (prefix << np.uint64(8)) | digit.astype(np.uint64)
        dec = np.where(prefix & top != 0, prefix ^ top, prefix ^ full)
        vals = dec.astype(np.uint32 if bits == 32 else np.uint64).view(np.float32 if bits == 32 else np.float64)
        lo, hi = torch.from_numpy(vals[0].copy()), torch.from_numpy(vals[1].copy())
        frac = torch.from_numpy(np.where((counts % 2 == 0) & (not lower), 0.5, 0.0)).to(x.dtype).view(k, 1)
        return lo + (hi - lo) * frac, torch.from_numpy(counts.copy())

    def nearest_rows_l1(self, x, p, row_base):
        k = p.shape[0]
        bd = torch.full((k,), float("inf"), dtype=torch.float64)
        bi = torch.full((k,), torch.iinfo(torch.int64).max, dtype=torch.int64)
        if x.shape[0]:
            dist = orc.manhattan_fast(x, p)
            m = torch.min(dist, dim=0)
            bd, bi = m.values.double(), m.indices + row_base
        return bd, bi

    def topk_rows(self, dmat, kk):
        v, i = torch.topk(dmat, kk, dim=1, largest=False)
        return v, i

    def knn_vote(self, idx, y):
        return torch.argmax(y[idx.flatten()].reshape(idx.shape + (y.shape[1],)).sum(dim=1), dim=1)
