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

    def assign(self, x, c, labels, fv=None, path="auto", row_ws=None):
        if x.shape[0] == 0:
            if fv is not None:
                fv.zero_()
            return
        lab, mins = orc.assign_to_cluster(x, c, eval_functional_value=True)
        labels.copy_(lab.to(labels.dtype).view(labels.shape))
        if fv is not None:
            fv[0] = float((mins.double() ** 2).sum())
