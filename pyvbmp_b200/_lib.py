"""ctypes binding of libvbmp_b200.so (the C ABI in include/vbmp_b200.h).

There is no CPU fallback: if the shared library is missing, or a tensor is not a CUDA tensor, the
call raises.  torch is used only for device memory and the current stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VBMP_LIB") or os.path.join(_HERE, "libvbmp_b200.so")   # VBMP_LIB: developer variant builds
_lib = None

FORCE_SIMT = int(os.environ.get("VBMP_FORCE_SIMT", "0"))   # tests: 1 = CUDA-core kernels for both, 2 = E-step only, 3 = Gram only


PROFILE = None     # bench.py sets this to {} to collect (start, end) CUDA events per C-ABI call
LAUNCHES = 0       # kernels launched by the library so far (vbmp_launch_count(): every launch site counts itself)
_PROFILE_AS = {"vbmp_estep_rpack": "vbmp_estep", "vbmp_diag_estep_rpack": "vbmp_diag_estep", "vbmp_gram_rpack": "vbmp_gram", "vbmp_gram_ex": "vbmp_gram",
               "vbmp_gram_zpack": "vbmp_gram"}


_EVENT_POOL = []   # timing events made ahead of a profiled region (profile_begin): creating them inside it costs CPU time per call


def profile_begin(n_calls=256):
    """Start collecting (start, end) CUDA events per C-ABI call into PROFILE, with the events of ``n_calls`` calls made up front."""
    global PROFILE
    while len(_EVENT_POOL) < 2 * n_calls:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()                      # forces the lazy cudaEventCreate now
        _EVENT_POOL.append(ev)
    torch.cuda.synchronize()
    PROFILE = {}


def profile_end():
    global PROFILE
    prof, PROFILE = PROFILE, None
    return prof or {}


class VbmpError(RuntimeError):
    pass


def lib():
    """Load libvbmp_b200.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VbmpError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C pyvbmp_b200/csrc` (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.vbmp_last_error.restype = c_char_p
        L.vbmp_version.restype = c_int
        L.vbmp_launch_count.restype = ctypes.c_ulonglong
        L.vbmp_estep_workspace_bytes.restype = c_size_t
        L.vbmp_estep_workspace_bytes.argtypes = [c_longlong, c_int, c_int, c_int, c_int]
        L.vbmp_rpack_bytes.restype = c_size_t
        L.vbmp_rpack_bytes.argtypes = [c_longlong, c_int]
        L.vbmp_gram_workspace_bytes.restype = c_size_t
        L.vbmp_gram_workspace_bytes.argtypes = [c_longlong, c_int, c_int, c_int, c_int, c_int]
        L.vbmp_gram_ex_workspace_bytes.restype = c_size_t
        L.vbmp_gram_ex_workspace_bytes.argtypes = [c_longlong, c_int, c_int, c_int, c_int, c_int, c_int, c_int]
        L.vbmp_diag_estep_workspace_bytes.restype = c_size_t
        L.vbmp_diag_estep_workspace_bytes.argtypes = [c_longlong, c_int, c_int, c_int, c_int]
        L.vbmp_zpack_bytes.restype = c_size_t
        L.vbmp_zpack_bytes.argtypes = [c_longlong, c_int, c_int]
        L.vbmp_softmax_rows_workspace_bytes.restype = c_size_t
        L.vbmp_softmax_rows_workspace_bytes.argtypes = [c_longlong, c_int]
        L.vbmp_rowgemm_workspace_bytes.restype = c_size_t
        L.vbmp_rowgemm_workspace_bytes.argtypes = [c_int, c_int, c_int]
        L.vbmp_rowterm_workspace_bytes.restype = c_size_t
        L.vbmp_rowterm_workspace_bytes.argtypes = [c_int, c_int]
        L.vbmp_wsum_workspace_bytes.restype = c_size_t
        L.vbmp_wsum_workspace_bytes.argtypes = [c_longlong, c_int, c_int]
        _lib = L
    return _lib


EXPORTS = (
    "vbmp_version", "vbmp_last_error", "vbmp_launch_count", "vbmp_niw_prep", "vbmp_mnw_prep", "vbmp_estep_workspace_bytes",
    "vbmp_estep", "vbmp_gram_workspace_bytes", "vbmp_gram", "vbmp_wishart_update", "vbmp_niw_update",
    "vbmp_mnw_update", "vbmp_wishart_elogdet", "vbmp_wishart_kl", "vbmp_niw_kl", "vbmp_mnw_kl",
    "vbmp_hmm_forward_backward", "vbmp_rpack_bytes", "vbmp_estep_rpack", "vbmp_gram_rpack",
    "vbmp_zpack_bytes", "vbmp_gram_zpack", "vbmp_gram_ex_workspace_bytes", "vbmp_gram_ex",
    "vbmp_diag_estep_workspace_bytes", "vbmp_diag_estep", "vbmp_diag_estep_rpack", "vbmp_mnw_prep_ex", "vbmp_moe_moments", "vbmp_rowgemm",
    "vbmp_wsum_workspace_bytes", "vbmp_wsum", "vbmp_rowterm_workspace_bytes", "vbmp_rowterm",
    "vbmp_rowgemm_workspace_bytes", "vbmp_rowgemm_ex", "vbmp_softmax_rows_workspace_bytes", "vbmp_softmax_rows",
)


def _check(rc, what):
    if rc != 0:
        raise VbmpError(f"{what} failed (code {rc}): {lib().vbmp_last_error().decode()}")


def _call(name, dev, *args):
    """Invoke one C-ABI entry point on ``dev``'s current stream (optionally bracketed by CUDA events).  The call runs under
    a device guard: kernel launches, attribute queries and the library's per-device caches all follow the CUDA *current*
    device, which need not be the tensors' device in a single-process multi-GPU program."""
    global LAUNCHES
    fn = getattr(lib(), name)
    with torch.cuda.device(dev):
        if PROFILE is not None:
            a = _EVENT_POOL.pop() if _EVENT_POOL else torch.cuda.Event(enable_timing=True)
            b = _EVENT_POOL.pop() if _EVENT_POOL else torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            PROFILE.setdefault(_PROFILE_AS.get(name, name), []).append((a, b))     # the hand-over variants time as K2 / K3
        else:
            rc = fn(*args)
    LAUNCHES = int(lib().vbmp_launch_count())
    _check(rc, name)


def _ptr(t):
    if t is None:
        return c_void_p(0)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise VbmpError("pyvbmp_b200 kernels need CUDA tensors (no CPU fallback); got "
                        f"{type(t).__name__} on {getattr(t, 'device', '?')}")
    if not t.is_contiguous():
        raise VbmpError("internal error: non-contiguous tensor reached the C ABI")
    return c_void_p(t.data_ptr())


def _stream(dev):
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def f32(t, device=None):
    """Contiguous fp32 copy/view of ``t`` (on ``device`` when given)."""
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    t = t.contiguous()
    if t.is_cuda and t.data_ptr() % 16:
        t = t.clone()        # the kernels use 16-byte vector / TMA accesses: a view at an odd element offset is re-based
    return t


def pad_dim(D):
    """Padded feature dimension the kernels tile by."""
    for Dp in (8, 16, 32, 64, 128):
        if D <= Dp:
            return Dp
    raise VbmpError(f"feature dimension {D} > 128 is outside the hot path this library implements")


_ws_cache = {}


def _workspace(nbytes, dev):
    """Per-device grow-only scratch buffer (kernels on one stream serialise, so reuse is safe)."""
    key = (dev.type, dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
        _ws_cache[key] = buf
    return buf


def release_workspaces():
    """Drop the cached scratch buffers (kernel workspaces and the K2 -> K3 hand-over images; ~10 GB at cfg2).  They are
    grow-only and per (device, stream); call this between workloads of very different size."""
    _ws_cache.clear()
    _rpack_cache.clear()
    _rpack_rec.clear()
    _zpack_cache.clear()


# ------------------------------------------------------------------------------------------------
# thin typed wrappers (all tensors contiguous fp32 CUDA unless noted)
# ------------------------------------------------------------------------------------------------

def niw_prep(invU, mu, nu, lam, logprior, C, d, Dp):
    dev = invU.device
    W = torch.empty((C, Dp, Dp), dtype=torch.float32, device=dev)
    m = torch.empty((C, Dp), dtype=torch.float32, device=dev)
    cst = torch.empty((C,), dtype=torch.float32, device=dev)
    info = torch.empty((C,), dtype=torch.int32, device=dev)
    _call("vbmp_niw_prep", dev, _ptr(invU), _ptr(mu), _ptr(nu), _ptr(lam), _ptr(logprior), c_int(C), c_int(d),
                               c_int(Dp), _ptr(W), _ptr(m), _ptr(cst), _ptr(info), _stream(dev))
    return W, m, cst, info


def mnw_prep(invU, nu, mu, invV, logprior, C, n, pp, pad, Dp, tau=None, elogdet=None):
    """tau (C, n) / elogdet (C): the diagonal-precision form (MatrixNormalGamma) in place of (invU, nu)."""
    dev = mu.device
    W = torch.empty((C, Dp, Dp), dtype=torch.float32, device=dev)
    m = torch.empty((C, Dp), dtype=torch.float32, device=dev)
    cst = torch.empty((C,), dtype=torch.float32, device=dev)
    info = torch.empty((C,), dtype=torch.int32, device=dev)
    if tau is not None:
        _call("vbmp_mnw_prep_ex", dev, _ptr(None), _ptr(None), _ptr(mu), _ptr(invV), _ptr(logprior), _ptr(tau), _ptr(elogdet),
              c_int(C), c_int(n), c_int(pp), c_int(int(pad)), c_int(Dp), _ptr(W), _ptr(m), _ptr(cst), _ptr(info), _stream(dev))
        return W, m, cst, info
    _call("vbmp_mnw_prep", dev, _ptr(invU), _ptr(nu), _ptr(mu), _ptr(invV), _ptr(logprior), c_int(C), c_int(n),
                               c_int(pp), c_int(int(pad)), c_int(Dp), _ptr(W), _ptr(m), _ptr(cst), _ptr(info),
                               _stream(dev))
    return W, m, cst, info


def diag_estep(x, N, GX, xg, mu, tau, cst, G, K, d, mode):
    """Diagonal-precision E-step (vbmp_diag_estep): x (N,GX,d), mu / tau (G*K,d), cst (G*K).  Returns logits (mode 0) or
    (p, logZn, NA, logZ) (mode 1)."""
    dev = x.device
    key = _rpack_key(dev)
    _rpack_rec.pop(key, None)
    out = torch.empty((N, G, K), dtype=torch.float32, device=dev)
    logZn = NA = logZ = None
    if mode == 1:
        logZn = torch.empty((N, G), dtype=torch.float32, device=dev)
        NA = torch.empty((G, K), dtype=torch.float32, device=dev)
        logZ = torch.empty((G,), dtype=torch.float32, device=dev)
    nbytes = lib().vbmp_diag_estep_workspace_bytes(c_longlong(N), c_int(G), c_int(K), c_int(d), c_int(mode))
    ws = _workspace(nbytes, dev)
    if mode == 1 and RPACK and G == 1 and GX == 1 and K <= 256 and not FORCE_SIMT:
        rb = int(lib().vbmp_rpack_bytes(c_longlong(N), c_int(K)))
        buf = _rpack_cache.get(key)
        if buf is None or buf.numel() < rb:
            buf = torch.empty(max(rb, 1), dtype=torch.uint8, device=dev)
            _rpack_cache[key] = buf
        packed = c_int(0)
        _call("vbmp_diag_estep_rpack", dev, _ptr(x), c_int(d), c_longlong(N), c_int(GX), _ptr(xg), _ptr(mu), _ptr(tau),
              _ptr(cst), c_int(G), c_int(K), c_int(mode), _ptr(out), _ptr(logZn), _ptr(NA), _ptr(logZ), _ptr(ws),
              c_size_t(ws.numel()), _stream(dev), _ptr(buf), c_size_t(buf.numel()), ctypes.byref(packed))
        if packed.value:
            _rpack_rec[key] = _tensor_rec(out) + (N, K)
        return out, logZn, NA, logZ
    _call("vbmp_diag_estep", dev, _ptr(x), c_int(d), c_longlong(N), c_int(GX), _ptr(xg), _ptr(mu), _ptr(tau), _ptr(cst),
          c_int(G), c_int(K), c_int(mode), _ptr(out), _ptr(logZn), _ptr(NA), _ptr(logZ), _ptr(ws), c_size_t(ws.numel()),
          _stream(dev))
    if mode == 0:
        return out
    return out, logZn, NA, logZ


# K2 -> K3 hand-over: the most recent mode-1 E-step on a stream may have left the responsibilities pre-split for the Gram
# kernel (vbmp_estep_rpack).  The record ties that buffer to the exact responsibilities it was made of: it HOLDS the
# storage object of that tensor (so the allocator cannot hand its address to another tensor while the record lives) and
# compares storage identity, offset, size and torch's version counter (any in-place ATen edit bumps it).  Writes that
# bypass ATen (another library's kernel through a raw pointer, DLPack consumers) are invisible to the counter: code that
# edits p that way must call invalidate_handover() (or set VBMP_RPACK=0).
RPACK = int(os.environ.get("VBMP_RPACK", "1"))
ZCACHE = int(os.environ.get("VBMP_ZCACHE", "1"))
_rpack_cache = {}     # stream key -> buffer
_rpack_rec = {}       # stream key -> (storage, storage offset, numel, version, N, K)
_zpack_cache = {}     # stream key -> (signature, storages, buffer): K3's sample image of the rows it was last given


def _rpack_key(dev):
    return (dev.type, dev.index, torch.cuda.current_stream(dev).cuda_stream)


def _same_tensor(rec, t):
    """rec = (storage, offset, numel, version) taken from a tensor earlier: is ``t`` that tensor, unedited?"""
    st = t.untyped_storage()
    return (st.data_ptr() == rec[0].data_ptr() and st.nbytes() == rec[0].nbytes() and t.storage_offset() == rec[1]
            and t.numel() == rec[2] and t._version == rec[3])


def _tensor_rec(t):
    return (t.untyped_storage(), t.storage_offset(), t.numel(), t._version)


def invalidate_handover():
    """Forget the K2 -> K3 weight images and the cached sample images (call after editing p / X behind torch's back)."""
    _rpack_rec.clear()
    _zpack_cache.clear()


def estep(z0, z1, N, GX, xg, W, m, cst, G, K, Dp, mode, out=None, logZn=None):
    """z0: (N,GX,d0), z1: (N,GX,d1) or None.  Returns logits (mode 0) or (p, logZn, NA, logZ) (mode 1)."""
    dev = z0.device
    d0 = z0.shape[-1]
    d1 = 0 if z1 is None else z1.shape[-1]
    key = _rpack_key(dev)
    _rpack_rec.pop(key, None)                    # whatever was packed before is about to be overwritten
    if out is None:
        out = torch.empty((N, G, K), dtype=torch.float32, device=dev)
    NA = logZ = None
    if mode == 1:
        if logZn is None:
            logZn = torch.empty((N, G), dtype=torch.float32, device=dev)
        NA = torch.empty((G, K), dtype=torch.float32, device=dev)
        logZ = torch.empty((G,), dtype=torch.float32, device=dev)
    nbytes = lib().vbmp_estep_workspace_bytes(c_longlong(N), c_int(G), c_int(K), c_int(Dp), c_int(mode))
    ws = _workspace(nbytes, dev)
    _note_path("estep", N, GX, G, K, Dp, d0, d1, True)
    if mode == 1 and RPACK and G == 1 and GX == 1 and K <= 256 and not FORCE_SIMT:
        rb = int(lib().vbmp_rpack_bytes(c_longlong(N), c_int(K)))
        buf = _rpack_cache.get(key)
        if buf is None or buf.numel() < rb:
            buf = torch.empty(max(rb, 1), dtype=torch.uint8, device=dev)
            _rpack_cache[key] = buf
        packed = c_int(0)
        _call("vbmp_estep_rpack", dev, _ptr(z0), c_int(d0), _ptr(z1), c_int(d1), c_longlong(N), c_int(GX), _ptr(xg),
              _ptr(W), _ptr(m), _ptr(cst), c_int(G), c_int(K), c_int(Dp), c_int(mode), c_int(0), _ptr(out), _ptr(logZn),
              _ptr(NA), _ptr(logZ), _ptr(ws), c_size_t(ws.numel()), _stream(dev), _ptr(buf), c_size_t(buf.numel()),
              ctypes.byref(packed))
        if packed.value:
            _rpack_rec[key] = _tensor_rec(out) + (N, K)
        return out, logZn, NA, logZ
    _call("vbmp_estep", dev, _ptr(z0), c_int(d0), _ptr(z1), c_int(d1), c_longlong(N), c_int(GX), _ptr(xg),
          _ptr(W), _ptr(m), _ptr(cst), c_int(G), c_int(K), c_int(Dp), c_int(mode),
          c_int(1 if FORCE_SIMT in (1, 2) else 0), _ptr(out), _ptr(logZn), _ptr(NA), _ptr(logZ),
          _ptr(ws), c_size_t(ws.numel()), _stream(dev))
    if mode == 0:
        return out
    return out, logZn, NA, logZ


def _zpack(z0, z1, N, K, Dp, key, dev):
    """K3's sample image of (z0, z1), made once per data set: reused while the rows are the same, unedited tensors."""
    if not ZCACHE:
        return None
    d0 = z0.shape[-1]
    d1 = 0 if z1 is None else z1.shape[-1]
    ent = _zpack_cache.get(key)
    if ent is not None:
        sig, recs, buf = ent
        if sig == (N, K, Dp, d0, d1) and _same_tensor(recs[0], z0) and (z1 is None or _same_tensor(recs[1], z1)):
            return buf
    zb = int(lib().vbmp_zpack_bytes(c_longlong(N), c_int(d0), c_int(d1)))
    buf = ent[2] if (ent is not None and ent[2].numel() >= zb) else None
    _zpack_cache.pop(key, None)
    if buf is None:
        buf = torch.empty(max(zb, 1), dtype=torch.uint8, device=dev)
    packed = c_int(0)
    _call("vbmp_gram_zpack", dev, _ptr(z0), c_int(d0), _ptr(z1), c_int(d1), c_longlong(N), c_int(K), c_int(Dp), _ptr(buf),
          c_size_t(buf.numel()), ctypes.byref(packed), _stream(dev))
    if not packed.value:
        return None
    _zpack_cache[key] = ((N, K, Dp, d0, d1), (_tensor_rec(z0), None if z1 is None else _tensor_rec(z1)), buf)
    return buf


def gram(z0, z1, N, GX, xg, p, GP, pg, G, K, Dp, diag=False):
    """Returns gram (G,K,D+1,D+1).  diag: only the diagonal, the last row / column and the corner are needed (the
    statistics of the diagonal-precision nodes); the tensor-core kernel then leaves zeros elsewhere."""
    dev = z0.device
    d0 = z0.shape[-1]
    d1 = 0 if z1 is None else z1.shape[-1]
    D1 = d0 + d1 + 1
    out = torch.empty((G, K, D1, D1), dtype=torch.float32, device=dev)
    key = _rpack_key(dev)
    _note_path("gram", N, GX, G, K, Dp, d0, d1, p is not None and GP == 1)
    if p is not None and not FORCE_SIMT and GP == 1 and GX == 1 and G == 1:
        rec = _rpack_rec.get(key)
        rbuf = None
        if rec is not None and rec[4:] == (N, K) and _same_tensor(rec, p):
            rbuf = _rpack_cache[key]
        zbuf = _zpack(z0, z1, N, K, Dp, key, dev)
        if rbuf is not None or zbuf is not None:
            nbytes = lib().vbmp_gram_ex_workspace_bytes(c_longlong(N), c_int(G), c_int(K), c_int(d0), c_int(d1), c_int(Dp),
                                                        c_int(rbuf is not None), c_int(zbuf is not None))
            ws = _workspace(nbytes, dev)
            _call("vbmp_gram_ex", dev, _ptr(z0), c_int(d0), _ptr(z1), c_int(d1), c_longlong(N), c_int(GX), _ptr(xg),
                  _ptr(p), c_int(GP), _ptr(pg), c_int(G), c_int(K), c_int(Dp), c_int(2 if diag else 0), _ptr(out), _ptr(ws),
                  c_size_t(ws.numel()), _stream(dev), _ptr(rbuf), _ptr(zbuf))
            return out
    nbytes = lib().vbmp_gram_workspace_bytes(c_longlong(N), c_int(G), c_int(K), c_int(d0), c_int(d1), c_int(Dp))
    ws = _workspace(nbytes, dev)
    _call("vbmp_gram", dev, _ptr(z0), c_int(d0), _ptr(z1), c_int(d1), c_longlong(N), c_int(GX), _ptr(xg),
          _ptr(p), c_int(GP), _ptr(pg), c_int(G), c_int(K), c_int(Dp),
          c_int((1 if FORCE_SIMT in (1, 3) else 0) | (2 if diag else 0)), _ptr(out), _ptr(ws), c_size_t(ws.numel()),
          _stream(dev))
    return out


_warned = set()


def _note_path(what, N, GX, G, K, Dp, d0, d1, weighted):
    """One warning per (kernel, reason) when a LARGE call leaves the tensor-core window and runs on the CUDA-core kernels
    (same results, ~20x slower at cfg2's shape: 559 vs 26 ms per iteration) — no silent performance cliff."""
    if FORCE_SIMT or N * K < (1 << 22):
        return
    why = []
    if G != 1 or GX != 1:
        why.append(f"replica / extra-event groups (G={G}, GX={GX})")
    if Dp not in _TC_DP:
        why.append(f"padded feature dimension {Dp}")
    if what == "gram":
        if not weighted:
            why.append("unit or per-group weights")
        if d0 % 4 or d1 % 4:
            why.append(f"feature blocks ({d0}, {d1}) not multiples of 4")
    if not why:
        return
    tag = (what, tuple(why))
    if tag in _warned:
        return
    _warned.add(tag)
    import warnings
    warnings.warn(f"pyvbmp_b200: {what} with N={N}, K={K} runs on the CUDA-core kernel, not tcgen05 ({'; '.join(why)}); "
                  "results are identical, throughput is ~20x lower", RuntimeWarning, stacklevel=3)


_TC_DP = (16, 32, 64, 128)


def wishart_update(SExx, N, invU0, nu0, invU_old, nu_old, C, d, lr):
    dev = SExx.device
    invU = torch.empty((C, d, d), dtype=torch.float32, device=dev)
    U = torch.empty((C, d, d), dtype=torch.float32, device=dev)
    nu = torch.empty((C,), dtype=torch.float32, device=dev)
    logdet = torch.empty((C,), dtype=torch.float32, device=dev)
    info = torch.empty((C,), dtype=torch.int32, device=dev)
    _call("vbmp_wishart_update", dev, _ptr(SExx), _ptr(N), _ptr(invU0), _ptr(nu0), _ptr(invU_old), _ptr(nu_old),
                                     c_int(C), c_int(d), c_float(lr), _ptr(invU), _ptr(nu), _ptr(U), _ptr(logdet),
                                     _ptr(info), _stream(dev))
    return invU, nu, U, logdet, info


def niw_update(SExx, SEx, N, lam0, mu0, invU0, nu0, lam_old, mu_old, invU_old, nu_old, C, d, lr, fixed_precision):
    dev = SExx.device
    lam = torch.empty((C,), dtype=torch.float32, device=dev)
    mu = torch.empty((C, d), dtype=torch.float32, device=dev)
    if fixed_precision:
        invU = nu = U = logdet = None
    else:
        invU = torch.empty((C, d, d), dtype=torch.float32, device=dev)
        U = torch.empty((C, d, d), dtype=torch.float32, device=dev)
        nu = torch.empty((C,), dtype=torch.float32, device=dev)
        logdet = torch.empty((C,), dtype=torch.float32, device=dev)
    info = torch.empty((C,), dtype=torch.int32, device=dev)
    _call("vbmp_niw_update", dev, _ptr(SExx), _ptr(SEx), _ptr(N), _ptr(lam0), _ptr(mu0), _ptr(invU0), _ptr(nu0),
                                 _ptr(lam_old), _ptr(mu_old), _ptr(invU_old), _ptr(nu_old), c_int(C), c_int(d),
                                 c_float(lr), c_int(int(bool(fixed_precision))), _ptr(lam), _ptr(mu), _ptr(invU),
                                 _ptr(nu), _ptr(U), _ptr(logdet), _ptr(info), _stream(dev))
    return lam, mu, invU, nu, U, logdet, info


def mnw_update(SExx, SEyx, SEyy, N, mu0, invV0, invU0, nu0, mu_old, invV_old, invU_old, nu_old, C, n, pp, lr,
               fixed_precision):
    dev = SExx.device
    e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)   # noqa: E731
    mu, invV, V, ldV = e(C, n, pp), e(C, pp, pp), e(C, pp, pp), e(C)
    if fixed_precision:
        invU = nu = U = ldU = None
    else:
        invU, nu, U, ldU = e(C, n, n), e(C), e(C, n, n), e(C)
    info = torch.empty((C,), dtype=torch.int32, device=dev)
    _call("vbmp_mnw_update", dev, _ptr(SExx), _ptr(SEyx), _ptr(SEyy), _ptr(N), _ptr(mu0), _ptr(invV0), _ptr(invU0),
                                 _ptr(nu0), _ptr(mu_old), _ptr(invV_old), _ptr(invU_old), _ptr(nu_old), c_int(C),
                                 c_int(n), c_int(pp), c_float(lr), c_int(int(bool(fixed_precision))), _ptr(mu),
                                 _ptr(invV), _ptr(V), _ptr(ldV), _ptr(invU), _ptr(nu), _ptr(U), _ptr(ldU),
                                 _ptr(info), _stream(dev))
    return mu, invV, V, ldV, invU, nu, U, ldU, info


def wishart_elogdet(nu, logdet, C, d):
    out = torch.empty((C,), dtype=torch.float32, device=nu.device)
    _call("vbmp_wishart_elogdet", nu.device, _ptr(nu), _ptr(logdet), c_int(C), c_int(d), _ptr(out), _stream(nu.device))
    return out


def wishart_kl(invU0, U, nu0, nu, logdet, logdet0, C, d):
    out = torch.empty((C,), dtype=torch.float32, device=U.device)
    _call("vbmp_wishart_kl", U.device, _ptr(invU0), _ptr(U), _ptr(nu0), _ptr(nu), _ptr(logdet), _ptr(logdet0), c_int(C),
                                 c_int(d), _ptr(out), _stream(U.device))
    return out


def niw_kl(lam0, lam, mu0, mu, invU0, U, nu0, nu, logdet, logdet0, C, d):
    out = torch.empty((C,), dtype=torch.float32, device=U.device)
    _call("vbmp_niw_kl", U.device, _ptr(lam0), _ptr(lam), _ptr(mu0), _ptr(mu), _ptr(invU0), _ptr(U), _ptr(nu0), _ptr(nu),
                             _ptr(logdet), _ptr(logdet0), c_int(C), c_int(d), _ptr(out), _stream(U.device))
    return out


def mnw_kl(mu0, mu, invV0, V, ldV, ldV0, invU0, U, nu0, nu, ldU, ldU0, C, n, pp):
    out = torch.empty((C,), dtype=torch.float32, device=U.device)
    _call("vbmp_mnw_kl", U.device, _ptr(mu0), _ptr(mu), _ptr(invV0), _ptr(V), _ptr(ldV), _ptr(ldV0), _ptr(invU0), _ptr(U),
                             _ptr(nu0), _ptr(nu), _ptr(ldU), _ptr(ldU0), c_int(C), c_int(n), c_int(pp), _ptr(out),
                             _stream(U.device))
    return out


def rowgemm(A, B, bias=None, out=None, accumulate=False):
    """out (N, M) (+)= A (N, Kd) @ B (Kd, M) (+ bias (M,)) with fp32-grade (3 x TF32) products (vbmp_rowgemm).  A, B, out: fp32,
    last dimension contiguous."""
    dev = A.device
    N, Kd = A.shape
    M = B.shape[1]
    assert B.shape[0] == Kd and A.stride(1) == 1 and B.stride(1) == 1
    if out is None:
        assert not accumulate
        out = torch.empty((N, M), dtype=torch.float32, device=dev)
    assert out.shape == (N, M) and out.stride(1) == 1
    nbytes = int(lib().vbmp_rowgemm_workspace_bytes(c_int(Kd), c_int(M), c_int(int(bias is not None)))) if not FORCE_SIMT else 0
    ws = _workspace(nbytes, dev) if nbytes else None
    _call("vbmp_rowgemm_ex", dev, c_void_p(A.data_ptr()), c_int(A.stride(0)), c_void_p(B.data_ptr()), c_int(B.stride(0)), _ptr(bias),
          c_void_p(out.data_ptr()), c_int(out.stride(0)), c_longlong(N), c_int(Kd), c_int(M), c_int(int(bool(accumulate))),
          _ptr(ws), c_size_t(ws.numel() if ws is not None else 0), _stream(dev))
    return out


def softmax_rows(logits, colbias=None, out=None):
    """logits (N, K) (+ colbias (K,)) -> (p (N, K), logZn (N,), NA (K,), logZ ()) (vbmp_softmax_rows); out may be logits itself."""
    dev = logits.device
    N, K = logits.shape
    assert logits.stride(1) == 1 and logits.dtype == torch.float32
    if out is None:
        out = torch.empty((N, K), dtype=torch.float32, device=dev)
    logZn = torch.empty((N,), dtype=torch.float32, device=dev)
    NA = torch.empty((K,), dtype=torch.float32, device=dev)
    logZ = torch.empty((), dtype=torch.float32, device=dev)
    nbytes = lib().vbmp_softmax_rows_workspace_bytes(c_longlong(N), c_int(K))
    ws = _workspace(nbytes, dev)
    _call("vbmp_softmax_rows", dev, c_void_p(logits.data_ptr()), c_int(logits.stride(0) if N > 0 else K), _ptr(colbias), c_longlong(N), c_int(K),
          c_void_p(out.data_ptr()), c_int(out.stride(0) if N > 0 else K), _ptr(logZn), _ptr(NA), _ptr(logZ), _ptr(ws), c_size_t(ws.numel()),
          _stream(dev))
    return out, logZn, NA, logZ


def rowterm_supported(N, F, K, lda):
    return N >= 128 and F >= 32 and F % 4 == 0 and lda % 4 == 0 and lda >= F and 1 <= K <= 256


def rowterm(A, B, C=None, alpha=1.0, accumulate=False):
    """C (N, K) = (accumulate ? C : 0) + alpha * A (N, F) @ B (F, K) for a long reduction F and few columns K (vbmp_rowterm:
    tcgen05, A through registers into tensor memory).  A may be row-strided; C is written in place when given."""
    dev = A.device
    N, F = A.shape
    K = B.shape[1]
    assert B.shape[0] == F and A.stride(1) == 1 and B.stride(1) == 1 and A.dtype == torch.float32 and B.dtype == torch.float32
    if C is None:
        assert not accumulate
        C = torch.empty((N, K), dtype=torch.float32, device=dev)
    assert C.shape == (N, K) and C.stride(1) == 1 and C.dtype == torch.float32
    nbytes = lib().vbmp_rowterm_workspace_bytes(c_int(F), c_int(K))
    ws = _workspace(nbytes, dev)
    _call("vbmp_rowterm", dev, c_void_p(A.data_ptr()), c_int(A.stride(0)), c_void_p(B.data_ptr()), c_int(B.stride(0)),
          c_void_p(C.data_ptr()), c_int(C.stride(0)), c_longlong(N), c_int(F), c_int(K), ctypes.c_float(alpha),
          c_int(int(bool(accumulate))), _ptr(ws), c_size_t(ws.numel()), _stream(dev))
    return C


def wsum_supported(N, K, F, lds):
    return N >= 2048 and K >= 4 and K % 4 == 0 and F >= 16 and lds % 4 == 0 and lds >= F


def wsum(p, S):
    """out (K, F) = p (N, K)^T @ S (N, F): responsibility-weighted column sums over the sample axis (vbmp_wsum)."""
    dev = S.device
    N, K = p.shape
    F = S.shape[1]
    assert S.shape[0] == N and p.is_contiguous() and S.stride(1) == 1
    out = torch.empty((K, F), dtype=torch.float32, device=dev)
    nbytes = lib().vbmp_wsum_workspace_bytes(c_longlong(N), c_int(K), c_int(F))
    ws = _workspace(nbytes, dev)
    assert S.is_cuda and S.dtype == torch.float32 and p.dtype == torch.float32
    _call("vbmp_wsum", dev, _ptr(p), c_void_p(S.data_ptr()), c_int(S.stride(0)), c_longlong(N), c_int(K), c_int(F), _ptr(out), _ptr(ws),
          c_size_t(ws.numel()), _stream(dev))
    return out


def moe_moments(mean, p, base, N, K, n, mu=None, Sigma=None):
    """mean (N,K,n), p (N,K), base (N,n,n) or None -> mu (N,n), Sigma (N,n,n)  (vbmp_moe_moments); mu / Sigma may be given
    (contiguous fp32 views to write into; Sigma may alias base)."""
    dev = mean.device
    if mu is None:
        mu = torch.empty((N, n), dtype=torch.float32, device=dev)
    if Sigma is None:
        Sigma = torch.empty((N, n, n), dtype=torch.float32, device=dev)
    _call("vbmp_moe_moments", dev, _ptr(mean), _ptr(p), _ptr(base), c_longlong(N), c_int(K), c_int(n), _ptr(mu), _ptr(Sigma),
          _stream(dev))
    return mu, Sigma


def hmm_forward_backward(logits, trans, init, T, S, G, K, ptemp):
    """logits (T,S,K), trans (G,K,K), init (G,K) contiguous fp32 -> p (T,S,K), SEzz (S,K,K), SEz0 (S,K), logZ (S)."""
    dev = logits.device
    p = torch.empty((T, S, K), dtype=torch.float32, device=dev)
    SEzz = torch.empty((S, K, K), dtype=torch.float32, device=dev)
    SEz0 = torch.empty((S, K), dtype=torch.float32, device=dev)
    logZ = torch.empty((S,), dtype=torch.float32, device=dev)
    _call("vbmp_hmm_forward_backward", dev, _ptr(logits), _ptr(trans), _ptr(init), c_int(T), c_longlong(S), c_int(G), c_int(K),
          c_float(float(ptemp)), _ptr(p), _ptr(SEzz), _ptr(SEz0), _ptr(logZ), _stream(dev))
    return p, SEzz, SEz0, logZ
