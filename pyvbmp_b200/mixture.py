"""Mixture / GaussianMixtureModel with the reference's interface (dists/Mixture.py:5-127,
models/GaussianMixtureModel.py:6-16).  The EM loop stays host Python; update_assignments is ONE fused
kernel sequence (K1 prep + K2 E-step with the logsumexp / responsibilities / NA / logZ epilogue) when
the mixture axis is the last batch dim of a NormalInverseWishart with a vector event.
"""
from __future__ import annotations

import torch

from . import _lib, _shapes, sharding
from .dirichlet import Dirichlet
from .niw import NormalInverseWishart
from .normal_gamma import NormalGamma


def fused_update_assignments(self, X, fallback=None):
    """dists/Mixture.py:38-45 on the CUDA path.  Also used as the method patch install() puts on the
    reference's Mixture class (``fallback`` = the reference's own method, taken for anything this does not fuse)."""
    dist = self.dist
    other = fallback if fallback is not None else generic_update_assignments
    fusable = (isinstance(dist, (NormalInverseWishart, NormalGamma)) and self.event_dim == 1 and dist.event_dim == 1
               and (fallback is None or dist.mu.is_cuda))
    if not fusable:
        return other(self, X)
    Xv = X.view(X.shape[:-dist.event_dim] + self.event_dim * (1,) + dist.event_shape)
    plan = dist._plan(Xv)
    if not plan.k_is_batch:
        return other(self, X)
    dev = dist.mu.device
    Xc = _lib.f32(Xv, dev).reshape(plan.N, plan.GX, dist.dim)
    if isinstance(dist, NormalGamma):          # diagonal precision: the streaming E-step with the same softmax epilogue
        mu, tau, cst = dist._prep(plan, logprior=self.pi.loggeomean())
        p, logZn, NA, logZ = _lib.diag_estep(Xc, plan.N, plan.GX, _shapes.idx_tensor(plan.xg, dev), mu, tau, cst,
                                             plan.G, plan.K, dist.dim, 1)
    else:
        W, m, cst, info, Dp = dist._prep(plan, logprior=self.pi.loggeomean())
        p, logZn, NA, logZ = _lib.estep(Xc, None, plan.N, plan.GX, _shapes.idx_tensor(plan.xg, dev), W, m, cst,
                                        plan.G, plan.K, Dp, 1)
    self.p = p.view(plan.sample_shape + plan.lead + (plan.K,))
    self.logZ_n = logZn.view(plan.sample_shape + plan.lead)
    self.NA = NA.view(plan.lead + (plan.K,))
    self.logZ = logZ.view(plan.lead)


def generic_update_assignments(self, X):
    """Any other event / batch structure: logits from dist.Elog_like (CUDA), softmax glue in torch."""
    log_p = self.Elog_like(X)
    dims = list(range(-self.event_dim, 0))
    logZ = torch.logsumexp(log_p, dim=dims)
    self.p = (log_p - logZ.view(logZ.shape + self.event_dim * (1,))).exp()
    sample_dim = self.p.ndim - self.batch_dim - self.event_dim
    self.NA = self.p.sum(list(range(sample_dim)))
    self.logZ = logZ.sum(list(range(sample_dim)))


class Mixture():

    def __init__(self, dist, event_shape, prior_parms={'alpha': torch.tensor(0.5)}):
        """dists/Mixture.py:8-19."""
        assert dist.batch_shape[-len(event_shape):] == event_shape
        self.event_shape = event_shape
        self.event_dim = len(event_shape)
        self.batch_shape = dist.batch_shape[:-len(event_shape)]
        self.batch_dim = len(self.batch_shape)

        self.pi = Dirichlet(event_shape=event_shape, batch_shape=self.batch_shape, prior_parms=prior_parms)
        self.dist = dist
        self.logZ = torch.tensor(-torch.inf, requires_grad=False)
        self.ELBO_last = torch.tensor(-torch.inf)

    def to_event(self, n):
        if n == 0:
            return self
        self.event_dim = self.event_dim + n
        self.event_shape = self.batch_shape[-n:] + self.event_shape
        self.batch_shape = self.batch_shape[:-n]
        self.pi.to_event(n)
        self.dist.to_event(n)
        return self

    def to(self, device):
        self.pi.to(device)
        self.dist.to(device)
        self.logZ = self.logZ.to(device)
        self.ELBO_last = self.ELBO_last.to(device)
        return self

    update_assignments = fused_update_assignments

    def update_parms(self, X, lr=1.0):
        """dists/Mixture.py:47-49."""
        self.pi.ss_update(self.NA, lr=lr)
        self.update_dist(X, lr=lr)

    def raw_update(self, X, iters=1, lr=1.0, verbose=False):
        self.update(X, iters=iters, lr=lr, verbose=verbose)

    def update(self, X, iters=1, lr=1.0, verbose=False):
        """dists/Mixture.py:54-62: E-step, ELBO with pre-M-step parameters, M-step.

        X may live in host memory (ideally pinned): the first iteration streams it through the device in row chunks,
        the H2D copy of chunk i+1 overlapping the E-step + Gram kernels of chunk i (same arithmetic, same results).  The
        reference's loop passes the same X to every iteration, so with iters > 1 the chunks land in ONE device tensor (when it
        fits beside the responsibilities) and iterations 2.. run on the resident rows: the host link is crossed once per
        call, not once per iteration."""
        if isinstance(X, torch.Tensor) and not X.is_cuda and self.dist.mu.is_cuda:
            keep = iters > 1 and self._rows_fit(X)
            ELBO = self._streamed_iteration(X, lr, keep=keep)
            if verbose:
                print('Percent Change in ELBO:   ', (ELBO - self.ELBO_last) / self.ELBO_last.abs() * 100.0)
            self.ELBO_last = ELBO
            if iters == 1:
                return
            if keep:
                X = self._stream_state.pop("resident")
                iters -= 1                                     # ... and fall through to the device loop below
            else:
                for i in range(iters - 1):
                    ELBO = self._streamed_iteration(X, lr)
                    if verbose:
                        print('Percent Change in ELBO:   ', (ELBO - self.ELBO_last) / self.ELBO_last.abs() * 100.0)
                    self.ELBO_last = ELBO
                return
        for i in range(iters):
            self.update_assignments(X)
            if sharding.enabled():
                ELBO = self._sharded_m_step(X, lr)
            else:
                ELBO = self.ELBO()
                self.update_parms(X, lr)
            if verbose:
                print('Percent Change in ELBO:   ', (ELBO - self.ELBO_last) / self.ELBO_last.abs() * 100.0)
            self.ELBO_last = ELBO

    STREAM_ROWS = 1 << 19      # rows per streamed chunk (128 MiB of X at d = 64)
    STREAM_FIRST = 1 << 16     # rows of the first chunk (its copy, 16 MiB at d = 64, is the only exposed one)
    STREAM_GROWTH = 1.3        # chunk i+1 / chunk i while ramping up

    def _rows_fit(self, Xh):
        """Room for a device copy of the rows beside the responsibilities, their operand images and the workspaces?"""
        dev = self.dist.mu.device
        K = self.dist.batch_shape[-1] if self.dist.batch_dim else 1
        need = Xh.numel() * 4 + 3 * Xh.shape[0] * K * 4 + (2 << 30)
        return torch.cuda.mem_get_info(dev)[0] > need

    def _streamed_iteration(self, Xh, lr, keep=False):
        """One EM iteration over host-resident rows: per chunk H2D (copy stream) -> K2 E-step -> K3 Gram on the compute
        stream; the statistics of the chunks are summed in a fixed order, then ELBO and the (optionally sharded) update.
        keep: the chunks land in one device tensor (left in self._stream_state["resident"]) instead of two staging buffers."""
        dist = self.dist
        dev = dist.mu.device
        fusable = (isinstance(dist, NormalInverseWishart) and self.event_dim == 1 and dist.event_dim == 1
                   and dist.batch_dim == 1 and Xh.ndim == 2)
        if not fusable:
            Xd = Xh.to(dev, non_blocking=True)
            self.update_assignments(Xd)
            if sharding.enabled():
                ELBO = self._sharded_m_step(Xd, lr)
            else:
                ELBO = self.ELBO()
                self.update_parms(Xd, lr)
            if keep:
                if getattr(self, "_stream_state", None) is None:
                    self._stream_state = {"key": None}
                self._stream_state["resident"] = Xd
            return ELBO
        N, d = Xh.shape
        K = dist.batch_shape[-1]
        plan = dist._plan(Xh[:1].view(1, 1, d))
        W, m, cst, info, Dp = dist._prep(plan, logprior=self.pi.loggeomean())
        xg = _shapes.idx_tensor((0,), dev)
        rows = min(self.STREAM_ROWS, max(N, 1))
        st = getattr(self, "_stream_state", None)
        if st is None or st["key"] != (N, d, K, str(dev)):
            st = {"key": (N, d, K, str(dev)), "copy": torch.cuda.Stream(dev),
                  "buf": None, "stage": None,
                  "free": [torch.cuda.Event() for _ in range(2)],
                  "p": torch.empty((N, K), dtype=torch.float32, device=dev),
                  "lz": torch.empty((N,), dtype=torch.float32, device=dev)}
            self._stream_state = st
        cur = torch.cuda.current_stream(dev)
        for e in st["free"]:
            e.record(cur)
        # Where the chunks land.  When a device copy of all rows fits, every chunk has its own place: the copies then run
        # back to back at the link's rate from the first microsecond, independent of the kernels (keep: that tensor is
        # handed on as the resident rows; otherwise it is a staging area reused by the next call).  Otherwise two staging
        # buffers alternate and the copy of chunk i+1 has to wait for the kernels of chunk i-1.
        if keep:
            full = torch.empty((N, d), dtype=torch.float32, device=dev)
        elif st["stage"] is not None or self._rows_fit(Xh):
            if st["stage"] is None:
                st["stage"] = torch.empty((N, d), dtype=torch.float32, device=dev)
            full = st["stage"]
        else:
            full = None
            if st["buf"] is None:
                st["buf"] = [torch.empty((rows, d), dtype=torch.float32, device=dev) for _ in range(2)]
        Gs = NA = logZ = None
        # Chunk sizes grow geometrically from STREAM_FIRST rows: only the first, small copy is exposed.  The growth factor
        # must stay below (kernel time per row) / (copy time per row) — 6.2 / 4.7 ns at cfg2 over PCIe 5 x16 — or every
        # chunk of the ramp waits for its rows (doubling cost 1.5 ms per iteration at cfg2).
        bounds, a, size = [], 0, float(min(rows, self.STREAM_FIRST))
        while a < N:
            step = min(int(size) // 256 * 256 or int(size), rows)
            bounds.append((a, min(a + step, N)))
            a += step
            size = min(size * self.STREAM_GROWTH, float(rows))
        for i, (a, b) in enumerate(bounds):
            buf = full[a:b] if full is not None else st["buf"][i & 1][: b - a]
            ready = torch.cuda.Event()
            with torch.cuda.stream(st["copy"]):
                if full is None or i == 0:
                    # staging pair: the kernels of chunk i-2 are done with this buffer; full-size staging: the kernels of
                    # the previous call are (free[0] was recorded on the compute stream above)
                    st["copy"].wait_event(st["free"][i & 1])
                buf.copy_(Xh[a:b], non_blocking=True)
                ready.record(st["copy"])
            cur.wait_event(ready)
            pc, lzc, NAc, lZc = _lib.estep(buf.view(b - a, 1, d), None, b - a, 1, xg, W, m, cst, 1, K, Dp, 1,
                                           out=st["p"][a:b].view(b - a, 1, K), logZn=st["lz"][a:b].view(b - a, 1))
            Gc = _lib.gram(buf.view(b - a, 1, d), None, b - a, 1, xg, pc, 1, xg, 1, K, Dp)
            if full is None:
                st["free"][i & 1].record(cur)
            Gs = Gc if Gs is None else Gs + Gc                   # fixed chunk order: deterministic
            NA = NAc if NA is None else NA + NAc
            logZ = lZc if logZ is None else logZ + lZc
        if keep:
            full.record_stream(st["copy"])
            st["resident"] = full
        self.p, self.logZ_n = st["p"], st["lz"]
        self.NA, self.logZ = NA.view(K), logZ.view(())
        if sharding.enabled():
            Gs, self.logZ, self.NA = sharding.all_reduce_packed([Gs, self.logZ, self.NA])
        ELBO = self.ELBO()
        self.pi.ss_update(self.NA, lr=lr)
        dist._update_from_gram(Gs, plan, True, lr)
        return ELBO

    def _sharded_m_step(self, X, lr):
        """Rows are sharded over ranks: local Gram + ONE all-reduce of [Gram | logZ | NA], then the
        replicated update (identical on every rank).  ELBO still uses the pre-update parameters."""
        Xv = X.view(X.shape[:-self.dist.event_dim] + self.event_dim * (1,) + self.dist.event_shape)
        if isinstance(self.dist, NormalGamma):
            SExx, SEx, N = self.dist._stats(Xv, self.p)
            SExx, SEx, N, logZ, NA = sharding.all_reduce_packed([SExx.contiguous(), SEx.contiguous(), N.contiguous(),
                                                                 self.logZ, self.NA])
            self.logZ, self.NA = logZ, NA
            ELBO = self.ELBO()
            self.pi.ss_update(self.NA, lr=lr)
            self.dist.ss_update(SExx, SEx, N, lr)
            return ELBO
        G, plan = self.dist._gram(Xv, self.p)
        G, logZ, NA = sharding.all_reduce_packed([G, self.logZ, self.NA])
        self.logZ, self.NA = logZ, NA
        ELBO = self.ELBO()
        self.pi.ss_update(self.NA, lr=lr)
        self.dist._update_from_gram(G, plan, True, lr)
        return ELBO

    def update_dist(self, X, lr):
        """dists/Mixture.py:64-66."""
        Xv = X.view(X.shape[:-self.dist.event_dim] + self.event_dim * (1,) + self.dist.event_shape)
        self.dist.raw_update(Xv, self.p, lr)

    def Elog_like(self, X):
        """dists/Mixture.py:68-70."""
        X = X.view(X.shape[:-self.dist.event_dim] + self.event_dim * (1,) + self.dist.event_shape)
        return self.dist.Elog_like(X) + self.pi.loggeomean()

    def KLqprior(self):
        return self.dist.KLqprior().sum(list(range(-self.event_dim, 0))) + self.pi.KLqprior()

    def ELBO(self):
        return self.logZ - self.KLqprior()

    def assignment_pr(self):
        return self.p

    def assignment(self):
        return self.p.argmax(-1)

    def means(self):
        return self.dist.mean()

    def event_average_f(self, function_string, A=None, keepdim=False):
        f = getattr(self.dist, function_string)
        return self.event_average(f() if A is None else f(A), keepdim=keepdim)

    def average_f(self, function_string, A=None, keepdim=False):
        f = getattr(self.dist, function_string)
        return self.average(f() if A is None else f(A), keepdim=keepdim)

    def average(self, A, keepdim=False):
        return (A * self.p).sum(-1, keepdim)

    def event_average(self, A, keepdim=False):
        out = (A * self.p.view(self.p.shape + (1,) * self.dist.event_dim)).sum(-1 - self.dist.event_dim, keepdim)
        for i in range(self.event_dim - 1):
            out = out.sum(-self.dist.event_dim - 1, keepdim)
        return out

    def stable_logsumexp(self, x, dim=None, keepdim=False):
        """dists/Mixture.py:110-127 (helper kept for callers of the reference's interface)."""
        dims = (dim,) if isinstance(dim, int) else tuple(dim)
        return torch.logsumexp(x, dim=dims, keepdim=keepdim)


class GaussianMixtureModel(Mixture):
    def __init__(self, nc, dim, isotropic=False):
        """models/GaussianMixtureModel.py:7-12."""
        if isotropic is False:
            dist = NormalInverseWishart(event_shape=(dim,), batch_shape=(nc,), scale=1.0 / nc ** (1.0 / dim))
        else:
            dist = NormalGamma(event_shape=(dim,), batch_shape=(nc,), scale=1.0 / nc ** (1.0 / dim))
        super().__init__(dist, event_shape=(nc,))

    def initialize(self, data):
        """models/GaussianMixtureModel.py:14-16."""
        idx = torch.randint(data.shape[0], self.event_shape)
        self.dist.mu = data[idx, :]
