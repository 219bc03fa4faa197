"""Tensor-level wrappers over the C ABI (include/dinomc.h).

PyTorch is used here only as plumbing: device memory (tensors), the current CUDA stream and dtype
bookkeeping.  Every arithmetic operation is a libdinomc kernel; nothing falls back to torch ops or
to the CPU -- non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

# kernel launches issued through this module (bench.py reports it as `gpu_launches`)
launch_count = 0

# Cap on the persistent GEMM grid (0 = all 148 SMs).  Data-parallel runs set it below 148 so that a concurrent
# NCCL all-reduce kernel finds free SMs instead of queueing behind (or stalling) the one-CTA-per-SM GEMMs.
gemm_max_ctas = 0
backward_max_ctas = 0        # same cap, applied only inside the backward Functions (set by GradAllReduce)

# data_ptr -> CUDA event recorded right after the kernel that produced a gradient buffer; lets the gradient
# all-reduce start as soon as that buffer is final instead of when its autograd node returns.
ready_events = {}


class backward_cap:
    """Context manager: GEMMs launched inside use `backward_max_ctas` as their persistent-grid cap."""

    def __enter__(self):
        global gemm_max_ctas
        self.saved = gemm_max_ctas
        if backward_max_ctas:
            gemm_max_ctas = backward_max_ctas

    def __exit__(self, *exc):
        global gemm_max_ctas
        gemm_max_ctas = self.saved
        return False


track_ready = False      # set by an active GradAllReduce: without a consumer no events are recorded (nothing would pop them)


def mark_ready(t: torch.Tensor):
    """Record "this buffer is final" on the current stream (see ready_events).  No-op unless a gradient exchange is active;
    the exchange pops an entry when it takes the gradient and clears the table at the end of its step (`GradAllReduce.wait`),
    so an entry never outlives the step that recorded it."""
    if not track_ready:
        return
    ev = torch.cuda.Event()
    ev.record()
    if len(ready_events) > 64:
        ready_events.clear()
    ready_events[t.data_ptr()] = ev


def _count(n=1):
    global launch_count
    launch_count += n


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------------
# optional live profiling: CUDA events on the launching stream around every libdinomc call
# ---------------------------------------------------------------------------------------------
_prof_events = None


def profile_begin():
    global _prof_events
    _prof_events = []


def profile_end():
    """Returns {op name: (total ms, calls)}; call after torch.cuda.synchronize()."""
    global _prof_events
    ev, _prof_events = _prof_events, None
    out = {}
    for name, e0, e1 in ev or []:
        ms, n = out.get(name, (0.0, 0))
        out[name] = (ms + e0.elapsed_time(e1), n + 1)
    return out


_annotate = False       # profile_kernels(): wrap every op in a torch.profiler range so CUPTI kernels are attributed to it


class _timed:
    __slots__ = ("name", "e0", "rf")

    def __init__(self, name):
        self.name = name
        self.rf = None

    def __enter__(self):
        if _prof_events is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        if _annotate:
            self.rf = torch.profiler.record_function("dmc:" + self.name)
            self.rf.__enter__()

    def __exit__(self, *exc):
        if self.rf is not None:
            self.rf.__exit__(*exc)
            self.rf = None
        if _prof_events is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof_events.append((self.name, self.e0, e1))
        return False


def profile_kernels(fn, steps: int):
    """Run `fn` `steps` times under torch.profiler (CUPTI activity records) with every libdinomc op wrapped in a
    profiler range, and return {op tag: (kernel microseconds per step, kernels per step)} -- pure kernel durations
    (no launch gaps, no event overhead), each kernel attributed to the op whose C-ABI call launched it.  Returns {}
    when the profiler delivered no attributed kernels (the caller then falls back to event timing)."""
    global _annotate
    from torch.profiler import ProfilerActivity, profile
    _annotate = True
    try:
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(steps):
                fn()
            torch.cuda.synchronize()
    finally:
        _annotate = False
    # Attribute every kernel to the op range whose C-ABI call launched it: kernel -> (correlation id) -> the CUDA runtime
    # launch call on the host -> the "dmc:<tag>" range that contains that call's start time.
    import bisect
    evs = list(prof.events())
    cuda_t = torch.autograd.DeviceType.CUDA
    ranges = sorted((e.time_range.start, e.time_range.end, e.name[4:]) for e in evs
                    if e.device_type != cuda_t and e.name.startswith("dmc:"))
    starts = [r[0] for r in ranges]
    launches = {}
    for e in evs:
        if e.device_type != cuda_t and e.name.startswith("cudaLaunch"):
            launches[e.id] = e.time_range.start
    out = {}
    for e in evs:
        if e.device_type != cuda_t or "memcpy" in e.name.lower() or "memset" in e.name.lower():
            continue
        t = launches.get(e.id)
        if t is None:
            continue
        i = bisect.bisect_right(starts, t) - 1
        if i < 0 or t > ranges[i][1]:
            continue
        us, n = out.get(ranges[i][2], (0.0, 0))
        out[ranges[i][2]] = (us + float(e.time_range.end - e.time_range.start), n + 1)
    return {k: (us / steps, n / steps) for k, (us, n) in out.items()}


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.DMC_F32
    if t.dtype == torch.bfloat16:
        return L.DMC_BF16
    raise TypeError(f"dinomc_b200 kernels take float32 or bfloat16 tensors, got {t.dtype}")


def _dtype_code(dtype: torch.dtype) -> int:
    return L.DMC_BF16 if dtype == torch.bfloat16 else L.DMC_F32


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dinomc_b200 has no CPU path: all tensors must be CUDA tensors "
                               f"(got a tensor on {t.device})")


def _p(t):
    return None if t is None else t.data_ptr()


def _rows2d(t: torch.Tensor) -> torch.Tensor:
    """2-D tensor whose rows are contiguous (stride(1) == 1); copies only if it must."""
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


# ---------------------------------------------------------------------------------------------
# workspace: one growing byte buffer per (device, stream); kernels on a stream use it in order
# ---------------------------------------------------------------------------------------------
_workspaces = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    key = (torch.device(device).index, _stream())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
def tma_ok(t: torch.Tensor) -> bool:
    """Can the tensor-core kernel's TMA descriptor express this stored matrix?"""
    return t.data_ptr() % 16 == 0 and (t.stride(0) * t.element_size()) % 16 == 0


def gemm(A, B, M, N, K, *, a_mn=False, b_mn=False, A_lo=None, B_lo=None, out=None, out_dtype=torch.float32,
         col_scale=None, bias=None, alpha=1.0, alpha_dev=None, act=L.ACT_NONE, aux=None, simt=False, split_k=0,
         tag="gemm", stats=None, row_scale=None, row_eps=0.0):
    """D[M,N] = epilogue(sum_k A(m,k) B(n,k)).  A is stored [M,K] (a_mn=False) or [K,M] (a_mn=True);
    B is stored [N,K] (b_mn=False) or [K,N] (b_mn=True).  See dmc_gemm in include/dinomc.h."""
    lib = L.load()
    A, B = _rows2d(A), _rows2d(B)
    _need_cuda(A, B, out, aux, col_scale, bias, alpha_dev, A_lo, B_lo)
    exp_a = (K, M) if a_mn else (M, K)
    exp_b = (K, N) if b_mn else (N, K)
    if tuple(A.shape) != exp_a or tuple(B.shape) != exp_b:
        raise ValueError(f"gemm: operand shapes {tuple(A.shape)}, {tuple(B.shape)} do not match "
                         f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}")
    if A.dtype != B.dtype:
        raise TypeError("gemm: A and B must share a dtype")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=A.device)
    g = L.GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn_major = A.data_ptr(), A.stride(0), int(a_mn)
    g.B, g.ldb, g.b_mn_major = B.data_ptr(), B.stride(0), int(b_mn)
    if A_lo is not None:
        A_lo, B_lo = _rows2d(A_lo), _rows2d(B_lo)
        assert A_lo.stride(0) == A.stride(0) and B_lo.stride(0) == B.stride(0)
    g.A_lo, g.B_lo = _p(A_lo), _p(B_lo)
    g.in_dtype = _dt(A)
    g.D, g.ldd, g.out_dtype = out.data_ptr(), out.stride(0), _dt(out)
    g.col_scale, g.bias, g.alpha_dev, g.alpha = _p(col_scale), _p(bias), _p(alpha_dev), float(alpha)
    g.act = act
    if aux is not None:
        g.aux, g.ldaux, g.aux_dtype = aux.data_ptr(), aux.stride(0), _dt(aux)
    g.split_k = split_k
    g.max_ctas = gemm_max_ctas
    if row_scale is not None:
        _need_cuda(row_scale)
        g.row_scale, g.row_eps = row_scale.data_ptr(), float(row_eps)
    if stats is not None:        # fused softmax / column-sum statistics of the stored output (see dmc_gemm_args)
        g.stat_scale = float(stats["scale"])
        g.stat_center = _p(stats.get("center"))
        g.stat_row_partials = stats["row_partials"].data_ptr()
        g.stat_colsum_partials = _p(stats.get("colsum_partials"))
        g.stat_bound = _p(stats.get("bound"))
        g.stat_bound2 = _p(stats.get("bound2"))
    if simt:
        with _timed(tag):
            L.check(lib.dmc_gemm_simt(C.byref(g), _stream()), "dmc_gemm_simt")
        _count()
        return out
    nbytes = lib.dmc_gemm_workspace_bytes(M, N, K, g.in_dtype) if split_k == 0 else split_k * M * N * 4
    if act == L.ACT_NORMALIZE_BWD:
        nbytes = max(nbytes, 2 * M * N * 4)         # this epilogue lives in the split-K reducer: the contraction is always split
    if nbytes:
        ws = workspace(nbytes, A.device)
        g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
    with _timed(tag):
        L.check(lib.dmc_gemm(C.byref(g), _stream()), "dmc_gemm")
    _count(2 if nbytes else 1)
    return out


def gemm_stats_parts(N: int) -> int:
    return int(L.load().dmc_gemm_stats_parts(N))


def teacher_finalize(row_partials, colsum_partials, Nt, K):
    """(row_stats [Nt,2], colsum [K] or None) from the partials the last-layer GEMM epilogue wrote.  With
    `colsum_partials` None only the row statistics are merged (the column sums then come from `rowdot`)."""
    lib = L.load()
    row_stats = torch.empty((Nt, 2), dtype=torch.float32, device=row_partials.device)
    colsum_ = torch.empty(K, dtype=torch.float32, device=row_partials.device) if colsum_partials is not None else None
    with _timed("teacher_finalize"):
        L.check(lib.dmc_teacher_finalize(row_partials.data_ptr(), _p(colsum_partials), Nt, K, row_partials.shape[1],
                                         colsum_partials.shape[0] if colsum_partials is not None else 0,
                                         row_stats.data_ptr(), _p(colsum_), _stream()),
                "dmc_teacher_finalize")
    _count()
    return row_stats, colsum_


def rowdot(W: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """out[k] = sum_j W[k, j] * x[j]  (W [K, dim] fp32 / bf16 contiguous, x fp32 [dim])."""
    lib = L.load()
    _need_cuda(W, x)
    if not W.is_contiguous() or W.dim() != 2 or x.dtype != torch.float32 or x.numel() != W.shape[1] or not x.is_contiguous():
        raise ValueError("rowdot: W must be a contiguous [K, dim] matrix and x a contiguous fp32 vector of dim entries")
    out = torch.empty(W.shape[0], dtype=torch.float32, device=W.device)
    with _timed("rowdot"):
        L.check(lib.dmc_rowdot(W.data_ptr(), _dt(W), W.shape[0], W.shape[1], x.data_ptr(), out.data_ptr(), _stream()), "dmc_rowdot")
    _count()
    return out


class no_pdl:
    """Context manager: kernels launched inside are NOT programmatic dependent launches.  For a kernel that follows a
    long-running kernel on its stream (e.g. the weight-norm backward behind the cross-rank exchange): launched early, its
    thousands of CTAs would sit resident in griddepcontrol.wait for the whole exchange and keep every SM full -- measured:
    the last layer's dgrad on the other stream could not start for 150 us."""

    def __enter__(self):
        self.prev = L.load().dmc_set_pdl(0)

    def __exit__(self, *exc):
        L.load().dmc_set_pdl(self.prev)
        return False


class polite:
    """Context manager: row-streaming kernels launched inside (weight-norm forward / backward, rowdot) use at most `n`
    CTAs (default one per SM) and so leave room for a one-CTA-per-SM GEMM on every SM (dmc_set_streaming_ctas)."""

    def __init__(self, n: int = 148):
        self.n = n

    def __enter__(self):
        self.prev = L.load().dmc_set_streaming_ctas(int(self.n))

    def __exit__(self, *exc):
        L.load().dmc_set_streaming_ctas(self.prev)
        return False


def lse_finalize(row_partials):
    lib = L.load()
    M, parts = row_partials.shape[0], row_partials.shape[1]
    lse = torch.empty(M, dtype=torch.float32, device=row_partials.device)
    with _timed("lse_finalize"):
        L.check(lib.dmc_lse_finalize(row_partials.data_ptr(), M, parts, lse.data_ptr(), _stream()), "dmc_lse_finalize")
    _count()
    return lse


def ce_fused(s, t, center, t_stats, s_lse, B, C, G, inv_ts, inv_tt):
    """One pass: (loss, ds for an upstream gradient of 1)."""
    lib = L.load()
    s, t = _rows2d(s), _rows2d(t)
    _need_cuda(s, t, center, t_stats, s_lse)
    K = s.shape[1]
    ds = torch.empty((s.shape[0], K), dtype=s.dtype, device=s.device)
    loss = torch.empty((), dtype=torch.float32, device=s.device)
    nbytes = lib.dmc_ce_workspace_bytes(B, C, G, K)
    ws = workspace(nbytes, s.device)
    with _timed("ce_fused"):
        L.check(lib.dmc_ce_fused(s.data_ptr(), _dt(s), s.stride(0), t.data_ptr(), _dt(t), t.stride(0), center.data_ptr(),
                                 t_stats.data_ptr(), s_lse.data_ptr(), B, C, G, K, inv_ts, inv_tt, ds.data_ptr(), _dt(ds),
                                 ds.stride(0), loss.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "dmc_ce_fused")
    _count(2)
    return loss, ds


def scale_inplace_if(x, scale_dev, expected=1.0):
    lib = L.load()
    _need_cuda(x, scale_dev)
    assert x.is_contiguous()
    scale_dev = scale_dev.to(torch.float32).contiguous()
    with _timed("scale_if"):
        L.check(lib.dmc_scale_inplace_if(x.data_ptr(), _dt(x), x.numel(), scale_dev.data_ptr(), float(expected), _stream()),
                "dmc_scale_inplace_if")
    _count()
    return x


def split_tf32(x: torch.Tensor):
    lib = L.load()
    _need_cuda(x)
    x = x.contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    with _timed("split_tf32"):
        L.check(lib.dmc_split_tf32(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream()), "dmc_split_tf32")
    _count()
    return hi, lo


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    _need_cuda(x)
    if x.dtype == torch.bfloat16:
        return x
    x = x.float().contiguous() if x.dtype != torch.float32 else x.contiguous()
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with _timed("cast_bf16"):
        L.check(lib.dmc_cast_f32_to_bf16(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "dmc_cast_f32_to_bf16")
    _count()
    return y


def cast_bf16_batch(tensors):
    """bf16 copies of several tensors with one launch (fp32 inputs; bf16 inputs are passed through)."""
    lib = L.load()
    outs = [None] * len(tensors)
    todo = []
    for i, t in enumerate(tensors):
        _need_cuda(t)
        if t.dtype == torch.bfloat16:
            outs[i] = t
        else:
            t = t.float().contiguous() if t.dtype != torch.float32 else t.contiguous()
            outs[i] = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
            todo.append((t, outs[i]))
    for k in range(0, len(todo), 8):
        grp = todo[k:k + 8]
        n = len(grp)
        srcs = (L.vp * n)(*[a.data_ptr() for a, _ in grp])
        dsts = (L.vp * n)(*[b.data_ptr() for _, b in grp])
        ns = (L.i64 * n)(*[a.numel() for a, _ in grp])
        with _timed("cast_bf16"):
            L.check(lib.dmc_cast_f32_to_bf16_batch(srcs, dsts, ns, n, _stream()), "dmc_cast_f32_to_bf16_batch")
        _count()
    return outs


def widen_bf16_batch(srcs, dsts):
    """dsts[i] (fp32, contiguous) <- srcs[i] (bf16, contiguous), eight tensors per launch (exact)."""
    lib = L.load()
    pairs = list(zip(srcs, dsts))
    for a, b in pairs:
        _need_cuda(a, b)
        if a.dtype != torch.bfloat16 or b.dtype != torch.float32 or a.numel() != b.numel() or not (a.is_contiguous() and b.is_contiguous()):
            raise ValueError("widen_bf16_batch: expects contiguous bf16 sources and fp32 destinations of equal size")
    for k in range(0, len(pairs), 8):
        grp = pairs[k:k + 8]
        n = len(grp)
        sp = (L.vp * n)(*[a.data_ptr() for a, _ in grp])
        dp = (L.vp * n)(*[b.data_ptr() for _, b in grp])
        ns = (L.i64 * n)(*[a.numel() for a, _ in grp])
        with _timed("widen_bf16"):
            L.check(lib.dmc_cast_bf16_to_f32_batch(sp, dp, ns, n, _stream()), "dmc_cast_bf16_to_f32_batch")
        _count()


def narrow_bf16_into(srcs, dsts):
    """dsts[i] (bf16 views, e.g. slices of one flat exchange buffer) <- srcs[i] (fp32), eight tensors per launch."""
    lib = L.load()
    pairs = list(zip(srcs, dsts))
    for a, b in pairs:
        _need_cuda(a, b)
        if a.dtype != torch.float32 or b.dtype != torch.bfloat16 or a.numel() != b.numel() or not (a.is_contiguous() and b.is_contiguous()):
            raise ValueError("narrow_bf16_into: expects contiguous fp32 sources and bf16 destinations of equal size")
    for k in range(0, len(pairs), 8):
        grp = pairs[k:k + 8]
        n = len(grp)
        sp = (L.vp * n)(*[a.data_ptr() for a, _ in grp])
        dp = (L.vp * n)(*[b.data_ptr() for _, b in grp])
        ns = (L.i64 * n)(*[a.numel() for a, _ in grp])
        with _timed("cast_bf16"):
            L.check(lib.dmc_cast_f32_to_bf16_batch(sp, dp, ns, n, _stream()), "dmc_cast_f32_to_bf16_batch")
        _count()


def colsum(X: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    X = _rows2d(X)
    _need_cuda(X)
    M, N = X.shape
    out = torch.empty(N, dtype=torch.float32, device=X.device)
    nbytes = lib.dmc_colsum_workspace_bytes(M, N)
    ws = workspace(nbytes, X.device)
    with _timed("colsum"):
        L.check(lib.dmc_colsum(X.data_ptr(), _dt(X), M, N, X.stride(0), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                "dmc_colsum")
    _count(2)
    return out


# ---------------------------------------------------------------------------------------------
# head row kernels
# ---------------------------------------------------------------------------------------------
def normalize_rows_fwd(z: torch.Tensor, eps=1e-12, want_bf16=False):
    """Returns (zhat_f32 [N,dim], zhat_bf16 or None, inv_den [N])."""
    lib = L.load()
    z = _rows2d(z)
    _need_cuda(z)
    n, dim = z.shape
    zhat = torch.empty((n, dim), dtype=torch.float32, device=z.device)
    zb = torch.empty((n, dim), dtype=torch.bfloat16, device=z.device) if want_bf16 else None
    inv_den = torch.empty(n, dtype=torch.float32, device=z.device)
    with _timed("normalize_fwd"):
        L.check(lib.dmc_normalize_rows_fwd(z.data_ptr(), _dt(z), n, dim, z.stride(0), eps, zhat.data_ptr(), _p(zb), None,
                                           inv_den.data_ptr(), _stream()), "dmc_normalize_rows_fwd")
    _count()
    return zhat, zb, inv_den


def normalize_rows_bwd(dzhat, zhat, inv_den, eps=1e-12, out_dtype=torch.float32):
    lib = L.load()
    _need_cuda(dzhat, zhat, inv_den)
    dzhat, zhat = dzhat.contiguous(), zhat.contiguous()
    assert dzhat.dtype == torch.float32 and zhat.dtype == torch.float32
    n, dim = zhat.shape
    dz = torch.empty((n, dim), dtype=out_dtype, device=zhat.device)
    with _timed("normalize_bwd"):
        L.check(lib.dmc_normalize_rows_bwd(dzhat.data_ptr(), zhat.data_ptr(), inv_den.data_ptr(), n, dim, eps, dz.data_ptr(),
                                           _dt(dz), _stream()), "dmc_normalize_rows_bwd")
    _count()
    return dz


last_gmax = [None]      # device scalar max|g| of the most recent weightnorm_fwd (read by NormLastLayerFn for the GEMM)


def weightnorm_fwd(v: torch.Tensor, g: torch.Tensor, mode: str):
    """mode 'bf16' -> (w_bf16, None); 'tf32x3' -> (w_hi, w_lo); 'f32' -> (w_f32, None).
    Also returns scale[K] = g/||v|| and inv_vnorm[K]."""
    lib = L.load()
    _need_cuda(v, g)
    v = v.contiguous()
    g = g.contiguous()
    if v.dtype != torch.float32 or g.dtype != torch.float32:
        raise TypeError("weight_norm parameters must be float32")
    K, dim = v.shape
    dev = v.device
    scale = torch.empty(K, dtype=torch.float32, device=dev)
    inv_vnorm = torch.empty(K, dtype=torch.float32, device=dev)
    gmax = torch.empty((), dtype=torch.float32, device=dev)      # max |g|: bound of the logits of unit rows
    w_f32 = w_lo = w_bf16 = None
    if mode == "bf16":
        w_bf16 = torch.empty((K, dim), dtype=torch.bfloat16, device=dev)
    else:
        w_f32 = torch.empty((K, dim), dtype=torch.float32, device=dev)
        if mode == "tf32x3":
            w_lo = torch.empty((K, dim), dtype=torch.float32, device=dev)
    with _timed("weightnorm_fwd"):
        L.check(lib.dmc_weightnorm_fwd(v.data_ptr(), g.data_ptr(), K, dim, _p(w_f32), _p(w_lo), _p(w_bf16), scale.data_ptr(),
                                       inv_vnorm.data_ptr(), gmax.data_ptr(), _stream()), "dmc_weightnorm_fwd")
    _count()
    last_gmax[0] = gmax
    return (w_bf16, None, scale, inv_vnorm) if mode == "bf16" else (w_f32, w_lo, scale, inv_vnorm)


def weightnorm_bwd(dw, v, scale, inv_vnorm, want_dg: bool):
    lib = L.load()
    _need_cuda(dw, v)
    dw, v = dw.contiguous(), v.contiguous()
    if dw.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"weightnorm_bwd: dW must be float32 or bfloat16, got {dw.dtype}")
    K, dim = v.shape
    dv = torch.empty_like(v)
    dg = torch.empty((K, 1), dtype=torch.float32, device=v.device) if want_dg else None
    fn, name = ((lib.dmc_weightnorm_bwd, "dmc_weightnorm_bwd") if dw.dtype == torch.float32
                else (lib.dmc_weightnorm_bwd_bf16, "dmc_weightnorm_bwd_bf16"))
    with _timed("weightnorm_bwd"):
        L.check(fn(dw.data_ptr(), v.data_ptr(), scale.data_ptr(), inv_vnorm.data_ptr(), K, dim, dv.data_ptr(), _p(dg), _stream()), name)
    _count()
    return dv, dg


# ---------------------------------------------------------------------------------------------
# loss kernels
# ---------------------------------------------------------------------------------------------
def absmax_into(x: torch.Tensor, out: torch.Tensor):
    """out[0] <- max |x| (fp32 device scalar slot; no synchronisation)."""
    lib = L.load()
    _need_cuda(x, out)
    x = x.reshape(-1)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise TypeError("absmax_into: contiguous float32 input")
    with _timed("absmax"):
        L.check(lib.dmc_absmax(x.data_ptr(), x.numel(), out.data_ptr(), _stream()), "dmc_absmax")
    _count()


def teacher_stats_colsum(t: torch.Tensor, center: torch.Tensor, inv_temp: float, bounds=None):
    """One pass over teacher logits: (row_stats [Nt,2] = (max, 1/sum), colsum [K]).  `bounds`: optional device tensor
    {>= max|t|, >= max|center|} enabling the fixed-shift fast path (see dmc_teacher_stats_colsum_bounded)."""
    lib = L.load()
    t = _rows2d(t)
    _need_cuda(t, center)
    Nt, K = t.shape
    center = center.reshape(-1)
    assert center.dtype == torch.float32 and center.numel() == K and center.is_contiguous()
    row_stats = torch.empty((Nt, 2), dtype=torch.float32, device=t.device)
    colsum_ = torch.empty(K, dtype=torch.float32, device=t.device)
    nbytes = lib.dmc_teacher_workspace_bytes(Nt, K)
    ws = workspace(nbytes, t.device)
    with _timed("teacher_stats_colsum"):
        if bounds is not None:
            _need_cuda(bounds)
            L.check(lib.dmc_teacher_stats_colsum_bounded(t.data_ptr(), _dt(t), Nt, K, t.stride(0), center.data_ptr(), inv_temp,
                                                         bounds.data_ptr(), row_stats.data_ptr(), colsum_.data_ptr(), ws.data_ptr(),
                                                         ws.numel(), _stream()), "dmc_teacher_stats_colsum_bounded")
        else:
            L.check(lib.dmc_teacher_stats_colsum(t.data_ptr(), _dt(t), Nt, K, t.stride(0), center.data_ptr(), inv_temp,
                                                 row_stats.data_ptr(), colsum_.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                    "dmc_teacher_stats_colsum")
    _count(2)
    return row_stats, colsum_


def center_update(center: torch.Tensor, colsum_: torch.Tensor, count: float, momentum: float) -> torch.Tensor:
    """Returns the NEW center tensor (same shape as `center`); `center` itself is left untouched."""
    lib = L.load()
    _need_cuda(center, colsum_)
    out = torch.empty_like(center)
    K = center.numel()
    with _timed("center_update"):
        L.check(lib.dmc_center_update(center.data_ptr(), out.data_ptr(), colsum_.data_ptr(), K, float(count), float(momentum),
                                      float(1 - momentum), _stream()), "dmc_center_update")
    _count()
    return out


def ce_fwd(s, t, center, t_stats, B, C, G, inv_ts, inv_tt):
    lib = L.load()
    s, t = _rows2d(s), _rows2d(t)
    _need_cuda(s, t, center, t_stats)
    K = s.shape[1]
    s_lse = torch.empty(C * B, dtype=torch.float32, device=s.device)
    loss = torch.empty((), dtype=torch.float32, device=s.device)
    nbytes = lib.dmc_ce_workspace_bytes(B, C, G, K)
    ws = workspace(nbytes, s.device)
    with _timed("ce_fwd"):
        L.check(lib.dmc_ce_fwd(s.data_ptr(), _dt(s), s.stride(0), t.data_ptr(), _dt(t), t.stride(0), center.data_ptr(),
                               t_stats.data_ptr(), B, C, G, K, inv_ts, inv_tt, s_lse.data_ptr(), loss.data_ptr(),
                               ws.data_ptr(), ws.numel(), _stream()), "dmc_ce_fwd")
    _count(2)
    return loss, s_lse


def ce_bwd(s, t, center, t_stats, s_lse, grad_out, B, C, G, inv_ts, inv_tt):
    lib = L.load()
    s, t = _rows2d(s), _rows2d(t)
    _need_cuda(s, t, center, t_stats, s_lse, grad_out)
    K = s.shape[1]
    ds = torch.empty((s.shape[0], K), dtype=s.dtype, device=s.device)
    grad_out = grad_out.to(torch.float32).contiguous()
    with _timed("ce_bwd"):
        L.check(lib.dmc_ce_bwd(s.data_ptr(), _dt(s), s.stride(0), t.data_ptr(), _dt(t), t.stride(0), center.data_ptr(),
                               t_stats.data_ptr(), s_lse.data_ptr(), grad_out.data_ptr(), B, C, G, K, inv_ts, inv_tt,
                               ds.data_ptr(), _dt(ds), ds.stride(0), _stream()), "dmc_ce_bwd")
    _count()
    return ds


# ---------------------------------------------------------------------------------------------
# EMA
# ---------------------------------------------------------------------------------------------
# StepGraph runs its warm-up steps with this set: every kernel of the step runs (allocator pools, plans and workspaces get
# their final shape) but persistent training state is left exactly as it was -- the EMA runs with m = 1 (an exact identity),
# DINOLoss keeps its center, optimizers refuse to step.
preserve_state = False
# filled while a StepGraph captures: ("ema", plan, m) / ("loss", module, epoch, temperature)
capture_notes = None


class EmaPlan:
    """Device-resident chunk table for one (teacher, student) parameter-list pair (plan v2, see dmc_ema_build_plan2).

    `shadows`: {index in the zipped lists: bf16 tensor} -- bf16 copies of the new teacher values written by the same pass;
    `wn`: (v_index, g_index, dim, w_bf16, scale, inv_norm) -- the weight-normed last layer handled row-wise.
    The momentum lives in a 2-float device buffer refilled from a ring of pinned host slots, so a captured graph reads
    the value of the CURRENT iteration (StepGraph.replay(momentum=...))."""

    _RING = 64

    def __init__(self, teacher_params, student_params, shadows=None, wn=None):
        lib = L.load()
        teacher_params, student_params = list(teacher_params), list(student_params)
        n = min(len(teacher_params), len(student_params))      # zip() semantics of main_dino_mc.py:405
        if n == 0:
            raise ValueError("ema: empty parameter list")
        tp, sp = teacher_params[:n], student_params[:n]
        for pk, pq in zip(tp, sp):
            _need_cuda(pk, pq)
            if pk.dtype != torch.float32 or pq.dtype != torch.float32:
                raise TypeError("ema: parameters must be float32")
            if pk.shape != pq.shape:
                raise ValueError(f"ema: shape mismatch {tuple(pk.shape)} vs {tuple(pq.shape)}")
            if not pk.is_contiguous() or not pq.is_contiguous():
                raise ValueError("ema: parameters must be contiguous")
        shadows = dict(shadows or {})
        for i, sh in shadows.items():
            _need_cuda(sh)
            if sh.dtype != torch.bfloat16 or not sh.is_contiguous() or sh.numel() != tp[i].numel():
                raise ValueError("ema: a shadow must be a contiguous bfloat16 tensor of the parameter's size")
        numels = (L.i64 * n)(*[pk.numel() for pk in tp])
        tptr = (L.vp * n)(*[pk.data_ptr() for pk in tp])
        sptr = (L.vp * n)(*[pq.data_ptr() for pq in sp])
        shp = (L.vp * n)(*[(shadows[i].data_ptr() if i in shadows else None) for i in range(n)])
        wv, wg, wdim, ww, wscale, winv = (-1, -1, 0, None, None, None) if wn is None else wn
        if wn is not None:
            _need_cuda(ww, wscale, winv)
        nbytes = lib.dmc_ema_plan2_bytes(numels, n, wv, wdim)
        host = torch.empty(max(nbytes, 8), dtype=torch.uint8).pin_memory()
        n_chunks = L.i64(0)
        L.check(lib.dmc_ema_build_plan2(tptr, sptr, numels, shp, n, wv, wg, wdim, _p(ww), _p(wscale), _p(winv), host.data_ptr(),
                                        host.numel(), C.byref(n_chunks)), "dmc_ema_build_plan2")
        self.n_chunks = n_chunks.value
        self.n_params = sum(pk.numel() for pk in tp)
        self.device = tp[0].device
        self.plan = host.to(self.device, non_blocking=False)
        self.keep = (list(shadows.values()), ww, wscale, winv)          # outputs the plan's raw pointers refer to
        self.scalars = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._ring = torch.zeros((self._RING, 2), dtype=torch.float32).pin_memory()
        self._ring_np = self._ring.numpy()
        self._ring_ev = [None] * self._RING
        self._ring_i = 0
        self.m_on_device = None

    def set_momentum(self, m: float):
        """Refill the device scalars (m, 1-m) on the current stream.  (1 - m) is formed in float64 like the reference
        (`(1 - m) * param_q`, m = np.float64 from cosine_scheduler), then both are rounded to fp32."""
        i = self._ring_i
        self._ring_i = (i + 1) % self._RING
        if self._ring_ev[i] is not None:
            self._ring_ev[i].synchronize()                      # the copy that last read this slot has completed (normally long ago)
        self._ring_np[i, 0] = float(m)
        self._ring_np[i, 1] = 1.0 - float(m)
        self.scalars.copy_(self._ring[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ring_ev[i] = ev
        self.m_on_device = float(m)

    def run(self, m: float):
        lib = L.load()
        m = 1.0 if preserve_state else float(m)
        if torch.cuda.is_current_stream_capturing():
            # the refill is not captured: StepGraph issues it before each replay (the graph must not freeze m)
            if capture_notes is not None:
                capture_notes.append(("ema", self, m))
        elif m != self.m_on_device:
            self.set_momentum(m)
        with _timed("ema"):
            L.check(lib.dmc_ema_multi_tensor2(self.plan.data_ptr(), self.n_chunks, self.scalars.data_ptr(), _stream()),
                    "dmc_ema_multi_tensor2")
        _count()


class ClipPlan:
    """Device-resident chunk table over a list of gradient tensors (see dmc_clip_grads)."""

    def __init__(self, grads):
        lib = L.load()
        grads = list(grads)
        if not grads:
            raise ValueError("clip: empty gradient list")
        for g in grads:
            _need_cuda(g)
            if g.dtype != torch.float32:
                raise TypeError("clip: gradients must be float32")
            if not g.is_contiguous():
                raise ValueError("clip: gradients must be contiguous")
        n = len(grads)
        self.key = tuple((g.data_ptr(), g.numel()) for g in grads)
        numels = (L.i64 * n)(*[g.numel() for g in grads])
        ptrs = (L.vp * n)(*[g.data_ptr() for g in grads])
        nbytes = lib.dmc_clip_plan_bytes(numels, n)
        host = torch.empty(max(nbytes, 8), dtype=torch.uint8).pin_memory()
        n_chunks = L.i64(0)
        L.check(lib.dmc_clip_build_plan(ptrs, numels, n, host.data_ptr(), host.numel(), C.byref(n_chunks)), "dmc_clip_build_plan")
        self.n_chunks, self.n_tensors = n_chunks.value, n
        self.device = grads[0].device
        self.plan = host.to(self.device, non_blocking=False)
        self.partials = torch.empty(max(self.n_chunks, 1), dtype=torch.float32, device=self.device)

    def run(self, clip: float) -> torch.Tensor:
        """Clips in place; returns the pre-clip norms [n_tensors] (device tensor, no synchronisation)."""
        lib = L.load()
        norms = torch.zeros(self.n_tensors, dtype=torch.float32, device=self.device)
        if self.n_chunks == 0:
            return norms
        with _timed("clip_grads"):
            L.check(lib.dmc_clip_grads(self.plan.data_ptr(), self.n_chunks, float(clip), norms.data_ptr(), self.partials.data_ptr(),
                                       self.partials.numel() * 4, _stream()), "dmc_clip_grads")
        _count(2)
        return norms


class AdamWPlan:
    """Device-resident chunk table over (param, grad, exp_avg, exp_avg_sq) quadruples of one parameter group."""

    def __init__(self, params, grads, exp_avgs, exp_avg_sqs):
        lib = L.load()
        n = len(params)
        if n == 0:
            raise ValueError("adamw: empty parameter group")
        for t in list(params) + list(grads) + list(exp_avgs) + list(exp_avg_sqs):
            _need_cuda(t)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError("adamw: parameters, gradients and moments must be contiguous float32 tensors")
        arr = lambda ts: (L.vp * n)(*[t.data_ptr() for t in ts])
        numels = (L.i64 * n)(*[p.numel() for p in params])
        nbytes = lib.dmc_adamw_plan_bytes(numels, n)
        host = torch.empty(max(nbytes, 8), dtype=torch.uint8).pin_memory()
        n_chunks = L.i64(0)
        L.check(lib.dmc_adamw_build_plan(arr(params), arr(grads), arr(exp_avgs), arr(exp_avg_sqs), numels, n, host.data_ptr(),
                                         host.numel(), C.byref(n_chunks)), "dmc_adamw_build_plan")
        self.n_chunks = n_chunks.value
        self.plan = host.to(params[0].device, non_blocking=False)

    def run(self, lr, beta1, beta2, eps, weight_decay, step):
        if self.n_chunks == 0:
            return
        lib = L.load()
        with _timed("adamw"):
            L.check(lib.dmc_adamw_multi_tensor(self.plan.data_ptr(), self.n_chunks, float(lr), float(beta1), float(beta2), float(eps),
                                               float(weight_decay), int(step), _stream()), "dmc_adamw_multi_tensor")
        _count()


class LarsPlan:
    """Device-resident chunk table over (param, grad, mu) triples of one parameter group (see dmc_lars_multi_tensor)."""

    def __init__(self, params, grads, mus):
        lib = L.load()
        n = len(params)
        if n == 0:
            raise ValueError("lars: empty parameter group")
        for t in list(params) + list(grads) + list(mus):
            _need_cuda(t)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError("lars: parameters, gradients and momentum buffers must be contiguous float32 tensors")
        arr = lambda ts: (L.vp * n)(*[t.data_ptr() for t in ts])
        numels = (L.i64 * n)(*[p.numel() for p in params])
        adapt = (L.i32 * n)(*[int(p.dim() != 1) for p in params])
        nbytes = lib.dmc_lars_plan_bytes(numels, n)
        host = torch.empty(max(nbytes, 8), dtype=torch.uint8).pin_memory()
        n_chunks = L.i64(0)
        L.check(lib.dmc_lars_build_plan(arr(params), arr(grads), arr(mus), numels, adapt, n, host.data_ptr(), host.numel(),
                                        C.byref(n_chunks)), "dmc_lars_build_plan")
        self.n_chunks = n_chunks.value
        self.plan = host.to(params[0].device, non_blocking=False)
        self.partials = torch.empty(max(2 * self.n_chunks, 2), dtype=torch.float32, device=params[0].device)

    def run(self, lr, weight_decay, momentum, eta):
        if self.n_chunks == 0:
            return
        lib = L.load()
        with _timed("lars"):
            L.check(lib.dmc_lars_multi_tensor(self.plan.data_ptr(), self.n_chunks, float(lr), float(weight_decay), float(momentum),
                                              float(eta), self.partials.data_ptr(), self.partials.numel() * 4, _stream()),
                    "dmc_lars_multi_tensor")
        _count(2)
