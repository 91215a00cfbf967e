#!/usr/bin/env python
"""bench.py — VB-EM sample·component updates/s (GMM d=64, K=256) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full EM iteration of GaussianMixtureModel NIW VB-EM (E-step + softmax + ELBO +
weighted Gram statistics [+ all-reduce] + NIW update) over the rank's rows.  At N=1 the workload is
BASELINE.json configs[1] (N=4 194 304, d=64, K=256, fp32); with N>1 ranks every rank owns the same
number of rows (weak scaling, sample-sharded, one all-reduce of the statistics per iteration).
Prints ONE JSON line on rank 0.  Besides the contract's keys the line carries
  * `secondary` (N=1): the rows next to the hot path, MixtureofLinearTransforms.update(pX, pY) / predict (SURVEY.md 8f,
    against the HBM peak), and BASELINE.json configs[2] (MixtureofLinearTransforms N=8 388 608, n=p=32, K=64) and configs[3]
    (ARHMM 4096 sequences x T=1024, d=16, K=32), each timed the same way (ms/step, per-kernel ms, roofline fraction);
  * `secondary.cfg1` (N=1): BASELINE.json configs[0] (two-moons, N=10 000, d=2, K=20) — the one configuration the reference
    runs at full size: the same call on the GPU and, like for like, by the unmodified reference on the host cores;
  * `cfg5` (N=8): BASELINE.json configs[4] at its stated size, 8 388 608 rows per GPU = 67 108 864 rows in total;
  * `e2e_iters20`: one public call update(X_pinned_host, iters=20) — the rows cross the host link once per call;
  * `replicas_bitwise_equal` (N>1): the posterior is bit-identical on every rank after the timed steps.
`--impl reference` times the reference's own CPU path on the host cores on a bounded sample of the same workload:
the UNMODIFIED reference when it is importable (baseline/_ref, `kind: "reference"`), else its restatement in
oracle/ (`kind: "port"`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "VB-EM sample*component updates/s (GMM d=64 K=256)"
UNIT = "sample*component updates/s"
D, K = 64, 256
ROWS_PER_GPU = 4_194_304
FLOPS_PER_UPDATE = 4 * D * D          # E 2d^2 + M 2d^2, each real multiply-add counted once
BYTES_PER_SAMPLE = 8 * D + 8 * K


def synth_rows(n, device, seed, k_true=256, d=D, chunk=32768):
    """cfg2 recipe: mu ~ 3 N(0,I), A_k = I + 0.3 randn/8, x = mu_z + A_z eps."""
    g = torch.Generator(device=device).manual_seed(4321)          # cluster parameters: same on every rank
    mu = 3.0 * torch.randn(k_true, d, generator=g, device=device)
    A = torch.eye(d, device=device) + 0.3 * torch.randn(k_true, d, d, generator=g, device=device) / 8
    g = torch.Generator(device=device).manual_seed(seed)           # rows: rank-offset stream
    X = torch.empty(n, d, device=device)
    for a in range(0, n, chunk):
        m = min(chunk, n - a)
        z = torch.randint(k_true, (m,), generator=g, device=device)
        e = torch.randn(m, d, 1, generator=g, device=device)
        X[a:a + m] = mu[z] + torch.bmm(A[z], e).squeeze(-1)
    return X


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx, self.first = [], None, gpu_index, 0

    def mark(self):
        """The timed region starts here: only samples from now on count (the process is started earlier, during warm-up, because
        nvidia-smi needs a few hundred ms before its first line)."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows, note = self.rows[self.first:], None
        if not rows and self.rows:               # a timed region shorter than the sampling period
            rows, note = self.rows[-1:], "no sample inside the timed region; the last one before it"
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def measure_tf32_peak(device, seconds=1.5):
    """cuBLAS TF32 8192^3 (same method as MEASURED_PEAKS.json's bf16 figure): burst and sustained TFLOP/s."""
    n = 8192
    a = torch.randn(n, n, device=device)
    b = torch.randn(n, n, device=device)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(3):
            a @ b
        torch.cuda.synchronize(device)
        best = 1e9
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); a @ b; e.record(); e.synchronize()
            best = min(best, s.elapsed_time(e))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(seconds * 1e3 / best))
        s.record()
        for _ in range(reps):
            a @ b
        e.record(); e.synchronize()
        sus = s.elapsed_time(e) / reps
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    fl = 2.0 * n ** 3
    return fl / best / 1e9, fl / sus / 1e9


def _reference_tree():
    """Path of an importable, unmodified pyVBMP tree (PYVBMP_REFERENCE, or the install under baseline/_ref), else None."""
    for p in (os.environ.get("PYVBMP_REFERENCE"), os.path.join(ROOT, "baseline", "_ref")):
        if p and os.path.isdir(os.path.join(p, "dists")) and os.path.isdir(os.path.join(p, "models")):
            return p
    return None


def cpu_iteration_rate(n_rows, iters, warmup, threads, seed=0):
    """Full EM iterations of the reference's CPU path on a bounded sample of the cfg2 workload.  The reference's algorithm is a
    broadcast multiply + sum over a materialised (N,K,d,d) temporary (dists/NormalInverseWishart.py:83,93), so 512 rows is
    the largest sample whose temporary fits comfortably (2 GiB).  Returns (updates/s, s/iteration, kind)."""
    torch.set_num_threads(threads)
    X = synth_rows(n_rows, torch.device("cpu"), 1234 + seed)
    ref = _reference_tree()
    if ref is not None:
        if ref not in sys.path:
            sys.path.insert(0, ref)
        import models as ref_models                      # the unmodified reference (no install(): its own torch code)
        torch.manual_seed(0)
        m = ref_models.GaussianMixtureModel(K, D)
        m.dist.mu = X[torch.randint(n_rows, (K,))].clone()
        step = lambda n: m.update(X, iters=n, lr=1.0, verbose=False)      # noqa: E731
        kind = "reference"
    else:
        from oracle import vbem_oracle as O
        torch.manual_seed(0)
        m = O.gmm_new(K, D)
        m["dist"]["mu"] = X[torch.randint(n_rows, (K,))].clone()
        step = lambda n: O.mixture_update(m, X, n, exact=True)             # noqa: E731
        kind = "port"
    step(max(warmup, 1))                                 # warm-up iterations (allocator, MKL threads)
    t0 = time.perf_counter()
    step(iters)
    dt = time.perf_counter() - t0
    return iters * n_rows * K / dt, dt / iters, kind


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_rows = args.ref_rows
    val, s_per, kind = cpu_iteration_rate(n_rows, args.steps, args.warmup, threads)
    what = ("the unmodified reference (models.GaussianMixtureModel.update on CPU torch)" if kind == "reference"
            else "oracle/ restatement of the reference's op order")
    sample = (f"{n_rows} rows of the cfg2 recipe per step (largest chunk whose (N,K,d,d) fp32 temporary fits "
              f"comfortably: {n_rows * K * D * D * 4 / 2**30:.1f} GiB); full EM iteration per step; {what}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "GaussianMixtureModel NIW VB-EM, d=64, K=256, fp32 (cfg2 recipe), CPU sample",
                   "rows_per_step": n_rows},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


_ALL_CPUS = set()      # the affinity mask this process started with (restored before the CPU baseline runs)


def numa_pin(dev_index):
    """Bind this process to the CPUs next to its GPU BEFORE any pinned host memory is allocated (first touch then places the
    staging buffers on the GPU's NUMA node).  Returns a short description for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = open(base + "/numa_node").read().strip()
        cpus = open(base + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        allowed = os.sched_getaffinity(0)
        _ALL_CPUS.update(allowed)
        ids &= allowed
        if ids and ids != allowed:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": cpus, "pinned": bool(ids and ids != allowed), "host_cpus": len(allowed)}
    except Exception as e:                       # containers without sysfs access: report and carry on
        return {"error": str(e)[:80]}


def time_steps(fn, steps, warmup, dev, barrier):
    """W warm-up + K timed steps, CUDA events on the current stream, per-C-ABI-call events collected on the side."""
    from pyvbmp_b200 import _lib
    for _ in range(warmup):
        fn()
    barrier()
    _lib.profile_begin(16 * steps)
    n0 = _lib.lib().vbmp_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    barrier()
    prof = _lib.profile_end()
    kern = {}
    for name, evs in prof.items():
        tt = sum(a.elapsed_time(b) for a, b in evs)
        kern[name] = {"calls": len(evs), "ms_total": tt, "ms_avg": tt / max(len(evs), 1), "ms_per_step": tt / steps}
    return ev0.elapsed_time(ev1), kern, int(_lib.lib().vbmp_launch_count() - n0)


def secondary_cfg1(dev, threads, iters=20, n_per=5000, Kc=20):
    """BASELINE.json configs[0] — the one configuration the reference runs today at its full size: GaussianMixtureModel on
    two-moons data (examples/two_moons.py:4-21), N = 10 000, d = 2, K = 20.  Same rows, same initial means, `iters` EM
    iterations in ONE public call on both sides: this package on the GPU (launch-bound at this size) and, when it is
    importable, the unmodified reference on the host cores — a like-for-like pair with the ELBO of both beside it."""
    import math
    import pyvbmp_b200 as V
    g = torch.Generator().manual_seed(5)
    x = torch.linspace(-math.pi / 2, math.pi / 2, n_per)
    X = torch.cat([torch.stack([torch.sin(x), torch.cos(x) - 0.25], -1),
                   torch.stack([torch.sin(x) + 1.0, -torch.cos(x) + 0.25], -1)], 0)
    X = X + 0.05 * torch.randn(X.shape, generator=g)
    X = X / X.std()
    N = X.shape[0]
    idx = torch.randint(N, (Kc,), generator=g)
    Xd = X.to(dev)

    def ours():
        torch.manual_seed(0)
        m = V.GaussianMixtureModel(Kc, 2)
        m.dist.mu = X[idx].clone()
        m.to(dev)
        return m
    ours().update(Xd, iters)                                  # warm-up call (workspaces, module load)
    torch.cuda.synchronize(dev)
    best = None
    for _ in range(3):
        m = ours()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        m.update(Xd, iters)
        elbo = float(m.ELBO_last)                             # device -> host read of the result closes the call
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out = {"workload": f"GaussianMixtureModel.update(X, {iters}), two-moons N={N}, d=2, K={Kc} (BASELINE.json configs[0]), "
                       "one public call, wall clock incl. the final device->host read; launch-bound",
           "ms_per_iteration": best / iters * 1e3, "value": iters * N * Kc / best, "unit": UNIT, "elbo_last": elbo}
    ref = _reference_tree()
    if ref is not None:
        if ref not in sys.path:
            sys.path.insert(0, ref)
        import models as ref_models
        torch.set_num_threads(threads)
        rbest = None
        for _ in range(3):
            torch.manual_seed(0)
            r = ref_models.GaussianMixtureModel(Kc, 2)
            r.dist.mu = X[idx].clone()
            t0 = time.perf_counter()
            r.update(X, iters=iters, lr=1.0, verbose=False)
            dt = time.perf_counter() - t0
            rbest = dt if rbest is None else min(rbest, dt)
        relbo = float(r.ELBO_last)
        out["reference_cpu"] = {"ms_per_iteration": rbest / iters * 1e3, "value": iters * N * Kc / rbest, "cores": threads,
                                "kind": "reference", "same_config": True, "elbo_last": relbo}
        out["elbo_rel_diff"] = abs(elbo - relbo) / abs(relbo)
    return out


def secondary_cfg3(dev, peak, steps=5, warmup=3, N=8_388_608, p=32, n=32, Kc=64):
    """BASELINE.json configs[2]: MixtureofLinearTransforms (MatrixNormalWishart regression), SURVEY.md §8d recipe."""
    import pyvbmp_b200 as V
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(N, p, generator=g, device=dev)
    W = torch.randn(Kc, n, p, generator=g, device=dev) / p ** 0.5
    b = torch.randn(Kc, n, generator=g, device=dev)
    z = torch.randint(Kc, (N,), generator=g, device=dev)
    Y = torch.empty(N, n, device=dev)
    for a in range(0, N, 1 << 20):
        e = min(a + (1 << 20), N)
        Y[a:e] = torch.einsum("nij,nj->ni", W[z[a:e]], X[a:e]) + b[z[a:e]] + 0.1 * torch.randn(e - a, n, generator=g, device=dev)
    torch.manual_seed(0)
    m = V.MixtureofLinearTransforms(n, p, Kc, pad_X=True).to(dev)
    Xc, Yc = X.unsqueeze(-1), Y.unsqueeze(-1)
    ms, kern, launches = time_steps(lambda: m.raw_update(Xc, Yc, iters=1, lr=1), steps, warmup, dev, lambda: torch.cuda.synchronize(dev))
    fl = 2 * (n * (p + 1) + n * n + (p + 1) ** 2) + 2 * (n + p + 1) ** 2       # 14 788 at n = p = 32
    val = steps * N * Kc / (ms / 1e3)
    return {"workload": f"MixtureofLinearTransforms raw_update, N={N}, p={p} (+1 pad), n={n}, K={Kc} (BASELINE.json configs[2])",
            "ms_per_step": ms / steps, "value": val, "unit": UNIT, "steps": steps, "warmup": warmup,
            "flops_per_update": fl, "roofline_frac": val * fl / 1e12 / peak, "gpu_launches": launches,
            "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kern.items()}, "elbo_last": float(m.ELBO_last)}


def secondary_cfg4(dev, peak, steps=5, warmup=3, S=4096, T=1024, d=16, Kc=32):
    """BASELINE.json configs[3]: ARHMM with MatrixNormalWishart emissions, 4096 sequences x T = 1024 (SURVEY.md §8d recipe)."""
    import pyvbmp_b200 as V
    g = torch.Generator(device=dev).manual_seed(2)
    A = 0.95 * torch.linalg.qr(torch.randn(Kc, d, d, generator=g, device=dev))[0]
    P = 4 * torch.eye(Kc, device=dev) + torch.rand(Kc, Kc, generator=g, device=dev)
    P = P / P.sum(-1, keepdim=True)
    y = torch.zeros(T + 1, S, d, device=dev)
    zt = torch.randint(Kc, (S,), generator=g, device=dev)
    y[0] = torch.randn(S, d, generator=g, device=dev)
    for t in range(T):
        y[t + 1] = torch.einsum("sij,sj->si", A[zt], y[t]) + 0.3 * torch.randn(S, d, generator=g, device=dev)
        zt = torch.multinomial(P[zt], 1, generator=g).squeeze(-1)
    X = y[:-1].reshape(T, S, 1, d, 1).contiguous()
    Y = y[1:].reshape(T, S, 1, d, 1).contiguous()
    torch.manual_seed(0)
    m = V.ARHMM(Kc, d, d).to(dev)
    ms, kern, launches = time_steps(lambda: m.update((X, Y), iters=1, lr=1), steps, warmup, dev, lambda: torch.cuda.synchronize(dev))
    fl = 2 * (d * (d + 1) + d * d + (d + 1) ** 2) + 2 * (2 * d + 1) ** 2       # 3 812 at d = 16
    val = steps * S * T * Kc / (ms / 1e3)
    return {"workload": f"ARHMM update, {S} sequences x T={T}, d={d}, K={Kc} (BASELINE.json configs[3])",
            "ms_per_step": ms / steps, "value": val, "unit": UNIT, "steps": steps, "warmup": warmup,
            "flops_per_update": fl, "roofline_frac": val * fl / 1e12 / peak, "gpu_launches": launches,
            "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kern.items()}, "elbo_last": float(m.ELBO_last)}


def secondary_molt_beliefs(dev, peak, steps=3, warmup=2, N=1 << 20, p=32, n=32, Kc=64, hbm=6546.2):
    """SURVEY.md §8f #2 / #3 (the rows next to the hot path): MixtureofLinearTransforms.update(pX, pY) on Gaussian beliefs with
    per-sample covariances, and MixtureofLinearTransforms.predict.  Both are bound by bytes (the flattened covariances in,
    the predictive covariances out): reported against the measured HBM peak."""
    import pyvbmp_b200 as V
    g = torch.Generator(device=dev).manual_seed(3)
    torch.manual_seed(0)
    m = V.MixtureofLinearTransforms(n, p, Kc).to(dev)
    X = torch.randn(N, p, 1, generator=g, device=dev)
    W = torch.randn(Kc, n, p, generator=g, device=dev) / p ** 0.5
    z = torch.randint(Kc, (N,), generator=g, device=dev)
    Y = (torch.einsum("nij,nj->ni", W[z], X[..., 0]) + 0.1 * torch.randn(N, n, generator=g, device=dev)).unsqueeze(-1)
    m.raw_update(X, Y, iters=2)
    Sx = (0.01 * torch.eye(p, device=dev)).expand(N, p, p).contiguous()
    Sy = (0.01 * torch.eye(n, device=dev)).expand(N, n, n).contiguous()
    pX, pY = V.MultivariateNormal_vector_format(mu=X, Sigma=Sx), V.MultivariateNormal_vector_format(mu=Y, Sigma=Sy)
    sync = lambda: torch.cuda.synchronize(dev)                                  # noqa: E731
    ms_u, kern_u, l_u = time_steps(lambda: m.update(pX, pY, iters=1), steps, warmup, dev, sync)
    ms_p, kern_p, l_p = time_steps(lambda: m.predict(X), steps, warmup, dev, sync)
    # algorithmic bytes: update reads both covariance sets twice (E and M) + means + responsibilities out and in;
    # predict reads the inputs and writes mu, Sigma and the gate probabilities
    b_u = N * (2 * 4 * (p * p + n * n) + 2 * 4 * (p + n) + 2 * 4 * Kc)
    b_p = N * (4 * p + 4 * n * n + 4 * n + 4 * Kc)
    out = {}
    for name, ms, kern, ln, by, what in (("molt_update_beliefs", ms_u, kern_u, l_u, b_u, "update(pX, pY)"),
                                         ("molt_predict", ms_p, kern_p, l_p, b_p, "predict(X)")):
        out[name] = {"workload": f"MixtureofLinearTransforms.{what}, N={N}, p={p}, n={n}, K={Kc}, per-sample covariances (SURVEY.md 8f)",
                     "ms_per_step": ms / steps, "value": steps * N * Kc / (ms / 1e3), "unit": UNIT, "steps": steps, "warmup": warmup,
                     "algorithmic_gb_per_step": by / 1e9, "hbm_frac": by / 1e9 / (ms / steps / 1e3) / hbm, "gpu_launches": ln,
                     "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kern.items()}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows-per-gpu", type=int, default=ROWS_PER_GPU)
    ap.add_argument("--ref-rows", type=int, default=512)
    ap.add_argument("--cpu-baseline-iters", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--cfg5", choices=["auto", "on", "off"], default="auto",
                    help="also time BASELINE.json configs[4] at 8 388 608 rows per GPU (auto: when 8 ranks run)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    numa = numa_pin(local_rank)                              # before the first pinned allocation
    import pyvbmp_b200 as V
    from pyvbmp_b200 import _lib, sharding
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
        sharding.enable()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def new_model(X, n_rows):
        # replicated init: same seed on every rank; the initial means are rank 0's rows, broadcast once
        torch.manual_seed(0)
        m = V.GaussianMixtureModel(K, D)
        idx = torch.randint(min(n_rows, 1 << 20), (K,))
        m.to(dev)
        mu0 = X[idx.to(dev)].clone()
        if world > 1:
            sharding.broadcast_(mu0, 0)
        m.dist.mu = mu0
        return m

    n_rows = args.rows_per_gpu
    X = synth_rows(n_rows, dev, 1234 + rank)
    m = new_model(X, n_rows)

    # ---- warm-up, then K timed steps (device-resident inputs) ----------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                          # streaming by the time the timed region begins
    for _ in range(args.warmup):
        m.update(X, 1)
    barrier()
    sampler.mark()
    ms, kern, launches = time_steps(lambda: m.update(X, 1), args.steps, 0, dev, barrier)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    elbo = float(m.ELBO_last)
    total_rows = n_rows * world
    value = args.steps * total_rows * K / (ms / 1e3)

    replicas_equal = None
    if world > 1:
        # the posterior must be the same bits on every rank (one all-reduce, then the identical replicated update)
        sig = torch.cat([m.ELBO_last.reshape(1).double(), m.dist.mu.double().sum().reshape(1),
                         m.dist.invU.invU.double().sum().reshape(1), m.pi.alpha.double().sum().reshape(1)])
        buf = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(buf, sig)
        replicas_equal = all(torch.equal(buf[0], b) for b in buf[1:])

    # ---- end-to-end through the public API with HOST buffers -----------------------------------------
    e2e = e2e20 = None
    if not args.no_e2e:
        # the public call with HOST rows: Mixture.update(X_host) streams them through the device in row chunks
        # (H2D of chunk i+1 under the kernels of chunk i) and the caller reads ELBO / NA back every step
        Xh = torch.empty(X.shape, dtype=X.dtype, pin_memory=True)
        Xh.copy_(X)
        res_h = torch.empty(1 + K, dtype=torch.float32, pin_memory=True)
        for _ in range(2):
            m.update(Xh, 1)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_steps = args.steps
        s0.record()
        for _ in range(e2e_steps):
            m.update(Xh, 1)                                        # H2D of this step's inputs happens inside
            res_h.copy_(torch.cat([m.ELBO_last.reshape(1), m.NA.reshape(-1)]), non_blocking=True)   # D2H of the result
            torch.cuda.current_stream().synchronize()             # the caller reads the ELBO every step
        s1.record()
        barrier()
        tms = max_over_ranks(s0.elapsed_time(s1))
        e2e = {"value": e2e_steps * total_rows * K / (tms / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": X.numel() * 4, "d2h_bytes_per_step": res_h.numel() * 4,
               "ms_per_step": tms / e2e_steps,
               "api": "GaussianMixtureModel.update(X_pinned_host, 1): chunked H2D overlapped with E-step + Gram, every step"}
        # ONE call of 20 iterations on host rows (the reference's own usage, dists/Mixture.py:54-62: same X every iteration):
        # the rows cross the host link once, iterations 2..20 run on the resident copy
        # warm-up of this path with THREE iterations: two resident ones, so that both alternating responsibilities buffers exist
        # (with two iterations the first timed call still had one 4 GiB cudaMalloc to do, tools/time_e2e20.py)
        m.update(Xh, 3)
        barrier()
        calls, allocs = [], []
        for _ in range(2):                                         # two calls, both reported; the figure is the faster one
            n_alloc = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            m.update(Xh, 20)
            res_h.copy_(torch.cat([m.ELBO_last.reshape(1), m.NA.reshape(-1)]), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            s1.record()
            barrier()
            calls.append(max_over_ranks(s0.elapsed_time(s1)))
            allocs.append(int(torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - n_alloc))
        tms = min(calls)
        e2e20 = {"value": 20 * total_rows * K / (tms / 1e3), "unit": UNIT, "iters_per_call": 20,
                 "h2d_bytes_per_call": X.numel() * 4, "d2h_bytes_per_call": res_h.numel() * 4, "ms_per_iteration": tms / 20,
                 "ms_per_call": [round(c, 2) for c in calls], "cuda_mallocs_in_call": allocs,
                 "api": "GaussianMixtureModel.update(X_pinned_host, 20): rows streamed once, then device-resident"}
        del Xh

    # ---- BASELINE.json configs[4] at its stated size: 8 388 608 rows per GPU -------------------------------------
    cfg5 = None
    if args.cfg5 == "on" or (args.cfg5 == "auto" and world == 8):
        del m, X
        _lib.release_workspaces()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
        n5 = 8_388_608
        X5 = synth_rows(n5, dev, 4321 + rank)
        m5 = new_model(X5, n5)
        steps5 = max(5, args.steps // 2)
        for _ in range(3):
            m5.update(X5, 1)
        barrier()
        ms5, kern5, _l5 = time_steps(lambda: m5.update(X5, 1), steps5, 0, dev, barrier)
        ms5 = max_over_ranks(ms5)
        hw = torch.tensor([float(torch.cuda.max_memory_allocated(dev))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(hw, op=dist.ReduceOp.MAX)
        v5 = steps5 * n5 * world * K / (ms5 / 1e3)
        cfg5 = {"workload": f"GaussianMixtureModel NIW VB-EM, N={n5 * world} ({n5} rows/GPU x {world} GPU), d={D}, K={K} "
                            "(BASELINE.json configs[4]; sample-sharded, one all-reduce of the statistics per iteration)",
                "rows_per_gpu": n5, "n_total": n5 * world, "steps": steps5, "warmup": 3, "ms_per_step": ms5 / steps5,
                "value": v5, "unit": UNIT, "elbo_last": float(m5.ELBO_last),
                "per_gpu_vs_4Mi_rows_rate": (v5 / world) / (value / world),
                "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kern5.items()},
                "step_tensor_frac": None, "mem_high_water_gib": float(hw) / 2 ** 30}
        del m5, X5
        m = X = None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf32_burst, tf32_sus = measure_tf32_peak(dev)
        bf16_sus = peaks.get("bf16_tflops_sustained")
        peak, peak_src = (bf16_sus, "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)") \
            if bf16_sus else (1409.1, "fallback: B200_PROFILING.md sustained bf16 figure (MEASURED_PEAKS.json absent)")
        if cfg5 is not None:
            cfg5["step_tensor_frac"] = cfg5["value"] * FLOPS_PER_UPDATE / world / 1e12 / peak
        dom = max(kern, key=lambda k: kern[k]["ms_total"]) if kern else None
        roof = None
        if dom is not None:
            per_launch_flops = 2.0 * n_rows * K * D * D          # E-step GEMM or M-step Gram: 2 d^2 per update
            ach = per_launch_flops / (kern[dom]["ms_per_step"] / 1e3) / 1e12
            traffic, traffic_src = None, None
            for fn in ("r02_traffic.json", "r01_traffic.json"):
                try:
                    tr = json.load(open(os.path.join(ROOT, "profiles", fn)))
                    if n_rows == ROWS_PER_GPU and dom in tr:
                        traffic = tr[dom]["bytes"] / 1e9           # GB per launch, from the committed ncu capture
                        traffic_src = f"static:profiles/{fn} (ncu --set full capture of this workload; not measured in this run)"
                        break
                except Exception:
                    pass
            # The kernels issue 16-bit-operand MMAs (kind::f16, the bf16 rate), three split terms per algorithmic product:
            # the roofline denominator is the measured dense bf16 figure; `issued` counts the MMAs actually executed
            # per algorithmic flop (2 d^2 per sample*component): E-step 3 terms on the triangular 62.5 % of the columns +
            # the TF32 "-m" step = 2.12x; Gram 3 terms on the 2160 symmetric pair columns (2145 pairs padded to 16) of 4096.
            issued = {"vbmp_estep": 2.12, "vbmp_gram": 3.0 * (((D + 1) * (D + 2) // 2 + 15) // 16 * 16) / (D * D)}
            roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_unit": "GB per launch (ncu dram__bytes_read+write)",
                    "traffic_source": traffic_src,
                    "issued_per_algorithmic_flop": issued,
                    "issued_frac": ach * issued.get(dom, 1.0) / peak,
                    "per_kernel_algorithmic_tflops": {k: round(per_launch_flops / (v["ms_per_step"] / 1e3) / 1e12, 1)
                                                      for k, v in kern.items() if k in ("vbmp_estep", "vbmp_gram")},
                    "peak_source": peak_src,
                    "peak_bf16_burst": peaks.get("bf16_tflops"),
                    "cublas_tf32_this_run": {"burst": tf32_burst, "sustained": tf32_sus},
                    "algorithmic_flops_per_launch": per_launch_flops,
                    "step_tensor_frac": value * FLOPS_PER_UPDATE / world / 1e12 / peak,
                    "step_hbm_frac": (value / K / world) * BYTES_PER_SAMPLE / 1e9 / peaks.get("hbm_gbs", 6546.2),
                    "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kern.items()}}
        secondary = None
        if world == 1 and not args.no_secondary:
            del m, X
            _lib.release_workspaces()
            torch.cuda.empty_cache()
            secondary = {}
            for name, fn in (("cfg3", secondary_cfg3), ("cfg4", secondary_cfg4)):
                try:
                    secondary[name] = fn(dev, peak)
                except Exception as e:                    # a secondary configuration must not take the headline down
                    secondary[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                _lib.release_workspaces()
                torch.cuda.empty_cache()
            try:
                secondary.update(secondary_molt_beliefs(dev, peak, hbm=peaks.get("hbm_gbs", 6546.2)))
            except Exception as e:
                secondary["molt_beliefs"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            _lib.release_workspaces()
            torch.cuda.empty_cache()
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            if _ALL_CPUS:
                os.sched_setaffinity(0, _ALL_CPUS)       # the CPU arm gets every host core, not just the GPU's NUMA node
            threads = os.cpu_count() or 1
            v, s_per, kind = cpu_iteration_rate(args.ref_rows, args.cpu_baseline_iters, 1, threads)
            what = ("the unmodified reference (baseline/_ref: models.GaussianMixtureModel.update, CPU torch)" if kind == "reference"
                    else "oracle/ restatement in the reference's op order")
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"{args.cpu_baseline_iters} EM iterations on {args.ref_rows} rows of the same workload "
                             f"({s_per:.2f} s/iteration; extrapolates to {s_per * n_rows / args.ref_rows:.0f} s per "
                             f"full-N iteration); {what}: (N,K,d,d) broadcast-multiply-sum"}
            if secondary is not None:                     # after the affinity reset: the reference half uses every host core
                try:
                    secondary["cfg1"] = secondary_cfg1(dev, threads)
                except Exception as e:
                    secondary["cfg1"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"GaussianMixtureModel NIW VB-EM, N={n_rows} rows/GPU x {world} GPU, d={D}, K={K}, "
                                   "fp32 (BASELINE.json configs[1]; sample-sharded weak scaling for N>1)",
                       "rows_per_gpu": n_rows, "l2": "inputs_exceed_l2 (X 1 GiB + responsibilities 4 GiB per step)",
                       "arithmetic": "fp32 inputs, outputs and accumulators; products as 3-term fp16 split (22 significant "
                                     "bits after exact power-of-two scaling) on tcgen05 kind::f16",
                       "data_layout": "K3 reads the rows through a transposed, pre-scaled image made once per data set "
                                      "(vbmp_gram_zpack, reused while X is the same unedited tensor); every step still reads "
                                      "X (E-step) and that image (Gram) from HBM",
                       "parallelism": f"sample-shard x{world}, 1 all-reduce/iter" if world > 1 else "single GPU",
                       "host_affinity": numa},
            "clocks": clocks, "e2e": e2e, "e2e_iters20": e2e20, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "elbo_last": elbo,
        }
        if replicas_equal is not None:
            line["replicas_bitwise_equal"] = replicas_equal
        if secondary is not None:
            line["secondary"] = secondary
        if cfg5 is not None:
            line["cfg5"] = cfg5
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
