"""CPU: the C-ABI library loads and exports every symbol include/dinomc.h declares (no compute calls),
argument validation that needs no device, and the host-only EMA plan builder."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dinomc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dmc_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_exports_every_declared_symbol():
    import dinomc_b200
    lib = dinomc_b200._lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libdinomc.so does not export {name}"
    assert sorted(dinomc_b200._lib.SIGNATURES) == declared      # the ctypes table covers the whole header
    assert lib.dmc_version() == 100
    assert lib.dmc_last_error_string() == b""


def test_missing_library_fails_loudly(monkeypatch):
    import dinomc_b200
    L = dinomc_b200._lib
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libdinomc.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        L.load()


def test_workspace_queries_are_host_only():
    import dinomc_b200
    lib = dinomc_b200._lib.load()
    # dgrad of the last layer (rows 2048, bottleneck 256, contraction 65536) is split-K: needs partials
    assert lib.dmc_gemm_workspace_bytes(2048, 256, 65536, 1) >= 2 * 2048 * 256 * 4
    # the forward is not split
    assert lib.dmc_gemm_workspace_bytes(2048, 65536, 256, 1) == 0
    assert lib.dmc_teacher_workspace_bytes(512, 65536) > 512 * 64 * 8
    assert lib.dmc_ce_workspace_bytes(256, 8, 2, 65536) > 0
    assert lib.dmc_colsum_workspace_bytes(2048, 2048) >= 2048 * 4


def test_argument_validation_without_a_device():
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    g = L.GemmArgs()
    assert lib.dmc_gemm(C.byref(g), None) < 0
    assert b"empty problem" in lib.dmc_last_error_string()
    assert lib.dmc_ce_fwd(None, 0, 0, None, 0, 0, None, None, 1, 1, 1, 1, 1.0, 1.0, None, None, None, 0, None) < 0
    assert b"null pointer" in lib.dmc_last_error_string()
    assert lib.dmc_ema_multi_tensor(None, 0, 0.9, 0.1, None) < 0
    with pytest.raises(RuntimeError, match="libdinomc"):
        L.check(-1, "unit test")


def test_ema_plan_builder_chunks_on_host():
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    numels = [5, 16384, 16385, 0, 100000]
    n = len(numels)
    arr = (L.i64 * n)(*numels)
    nbytes = lib.dmc_ema_plan_bytes(arr, n)
    expect_chunks = sum((x + 16383) // 16384 for x in numels)
    assert nbytes == expect_chunks * 24
    tp = (L.vp * n)(*[0x1000 * (i + 1) for i in range(n)])
    sp = (L.vp * n)(*[0x100000 * (i + 1) for i in range(n)])
    buf = (C.c_uint8 * nbytes)()
    out = L.i64(0)
    assert lib.dmc_ema_build_plan(tp, sp, arr, n, buf, nbytes, C.byref(out)) == 0
    assert out.value == expect_chunks
    entries = [(int.from_bytes(bytes(buf[i * 24:i * 24 + 8]), "little"),
                int.from_bytes(bytes(buf[i * 24 + 8:i * 24 + 16]), "little"),
                int.from_bytes(bytes(buf[i * 24 + 16:i * 24 + 24]), "little")) for i in range(expect_chunks)]
    assert entries[0] == (0x1000, 0x100000, 5)
    assert entries[1] == (0x2000, 0x200000, 16384)
    assert entries[2] == (0x3000, 0x300000, 16384) and entries[3] == (0x3000 + 16384 * 4, 0x300000 + 16384 * 4, 1)
    assert sum(e[2] for e in entries) == sum(numels)
    assert lib.dmc_ema_build_plan(tp, sp, arr, n, buf, 8, C.byref(out)) < 0      # buffer too small


def test_clip_plan_builder_chunks_on_host():
    import struct
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    numels = [5, 16385, 0, 40000]
    n = len(numels)
    arr = (L.i64 * n)(*numels)
    nbytes = lib.dmc_clip_plan_bytes(arr, n)
    expect_chunks = sum((x + 16383) // 16384 for x in numels)
    assert nbytes == expect_chunks * 32
    ptrs = (L.vp * n)(*[0x10000 * (i + 1) for i in range(n)])
    buf = (C.c_uint8 * nbytes)()
    out = L.i64(0)
    assert lib.dmc_clip_build_plan(ptrs, arr, n, buf, nbytes, C.byref(out)) == 0
    assert out.value == expect_chunks
    entries = [struct.unpack("<qqiiii", bytes(buf[i * 32:(i + 1) * 32])) for i in range(expect_chunks)]
    # (grad ptr, n, tensor, first chunk, chunk count, pad)
    assert entries[0][:5] == (0x10000, 5, 0, 0, 1)
    assert entries[1][:5] == (0x20000, 16384, 1, 1, 2) and entries[2][:5] == (0x20000 + 16384 * 4, 1, 1, 1, 2)
    assert [e[2] for e in entries[3:]] == [3, 3, 3] and all(e[3] == 3 and e[4] == 3 for e in entries[3:])
    assert sum(e[1] for e in entries) == sum(numels)
    assert lib.dmc_clip_build_plan(ptrs, arr, n, buf, 8, C.byref(out)) < 0       # buffer too small
    assert lib.dmc_clip_grads(None, 1, 1.0, None, None, 0, None) < 0             # argument validation without a device


def test_adamw_plan_builder_and_validation_on_host():
    import struct
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    numels = [3, 20000]
    n = len(numels)
    arr = (L.i64 * n)(*numels)
    nbytes = lib.dmc_adamw_plan_bytes(arr, n)
    assert nbytes == 3 * 40
    mk = lambda base: (L.vp * n)(*[base * (i + 1) for i in range(n)])
    buf = (C.c_uint8 * nbytes)()
    out = L.i64(0)
    assert lib.dmc_adamw_build_plan(mk(0x1000), mk(0x2000), mk(0x4000), mk(0x8000), arr, n, buf, nbytes, C.byref(out)) == 0
    assert out.value == 3
    e = [struct.unpack("<qqqqq", bytes(buf[i * 40:(i + 1) * 40])) for i in range(3)]
    assert e[0] == (0x1000, 0x2000, 0x4000, 0x8000, 3)
    assert e[2] == (0x2000 + 16384 * 4, 0x4000 + 16384 * 4, 0x8000 + 16384 * 4, 0x10000 + 16384 * 4, 20000 - 16384)
    assert lib.dmc_adamw_multi_tensor(None, 1, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, None) < 0     # null plan
    assert lib.dmc_adamw_multi_tensor(buf, 1, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0, None) < 0      # step must be >= 1
    assert b"step" in lib.dmc_last_error_string()


def test_lars_plan_builder_and_validation_on_host():
    import struct
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    numels = [3, 20000]
    n = len(numels)
    arr = (L.i64 * n)(*numels)
    adapt = (L.i32 * n)(0, 1)
    nbytes = lib.dmc_lars_plan_bytes(arr, n)
    assert nbytes == 3 * 48
    mk = lambda base: (L.vp * n)(*[base * (i + 1) for i in range(n)])
    buf = (C.c_uint8 * nbytes)()
    out = L.i64(0)
    assert lib.dmc_lars_build_plan(mk(0x1000), mk(0x2000), mk(0x4000), arr, adapt, n, buf, nbytes, C.byref(out)) == 0
    assert out.value == 3
    e = [struct.unpack("<qqqqiiii", bytes(buf[i * 48:(i + 1) * 48])) for i in range(3)]
    assert e[0] == (0x1000, 0x2000, 0x4000, 3, 0, 1, 0, 0)                       # 1-D tensor: no adaptation, own chunk range
    assert e[1] == (0x2000, 0x4000, 0x8000, 16384, 1, 2, 1, 0)
    assert e[2] == (0x2000 + 16384 * 4, 0x4000 + 16384 * 4, 0x8000 + 16384 * 4, 20000 - 16384, 1, 2, 1, 0)
    assert lib.dmc_lars_build_plan(mk(0x1000), mk(0x2000), mk(0x4000), arr, adapt, n, buf, nbytes - 1, C.byref(out)) < 0
    assert lib.dmc_lars_multi_tensor(None, 1, 0.1, 0.0, 0.9, 0.001, None, 0, None) < 0      # null plan
    ws = (C.c_float * 2)()
    assert lib.dmc_lars_multi_tensor(buf, 3, 0.1, 0.0, 0.9, 0.001, ws, 8, None) < 0         # workspace too small for 3 chunks
    assert b"workspace" in lib.dmc_last_error_string()
    import inspect
    assert list(inspect.signature(dinomc_b200.FusedLARS.__init__).parameters)[1:] == [
        "params", "lr", "weight_decay", "momentum", "eta", "weight_decay_filter", "lars_adaptation_filter"]


def test_module_surface_matches_reference_signature():
    import inspect
    import dinomc_b200 as D
    assert list(inspect.signature(D.DINOHead.__init__).parameters)[1:] == [
        "in_dim", "out_dim", "use_bn", "norm_last_layer", "nlayers", "hidden_dim", "bottleneck_dim"]
    assert list(inspect.signature(D.DINOLoss.__init__).parameters)[1:] == [
        "out_dim", "ncrops", "warmup_teacher_temp", "teacher_temp", "warmup_teacher_temp_epochs", "nepochs",
        "teacher_crops_number", "student_temp", "center_momentum"]
    assert list(inspect.signature(D.DINOLoss.forward).parameters)[1:] == ["student_output", "teacher_output", "epoch"]
    h = D.DINOHead(384, 4096)
    assert [n for n, _ in h.named_parameters()] == [
        "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias",
        "last_layer.weight_g", "last_layer.weight_v"]
    assert h.last_layer.weight_g.shape == (4096, 1) and not h.last_layer.weight_g.requires_grad
    assert torch.all(h.last_layer.weight_g == 1)
    assert D.DINOHead(384, 64, norm_last_layer=False).last_layer.weight_g.requires_grad
    one = D.DINOHead(32, 64, nlayers=1)
    assert [n for n, _ in one.named_parameters()][:2] == ["mlp.weight", "mlp.bias"]
    bn = D.DINOHead(32, 64, use_bn=True, hidden_dim=16)
    assert any(isinstance(m, torch.nn.BatchNorm1d) for m in bn.mlp)
    loss = D.DINOLoss(64, 8, 0.04, 0.07, 3, 10)
    assert list(loss.state_dict().keys()) == ["center"] and loss.center.shape == (1, 64)
    assert len(loss.teacher_temp_schedule) == 10 and loss.teacher_temp_schedule[-1] == 0.07


@pytest.mark.reference
def test_init_matches_reference_rng_stream():
    """Same torch seed -> same initial weights as the reference's DINOHead (container only)."""
    from oracle import reference_loader
    if not reference_loader.available():
        pytest.skip("reference not present")
    import warnings
    import dinomc_b200 as D
    _, vits, _ = reference_loader.load()
    torch.manual_seed(123)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = vits.DINOHead(48, 256, hidden_dim=64, bottleneck_dim=32)
    torch.manual_seed(123)
    ours = D.DINOHead(48, 256, hidden_dim=64, bottleneck_dim=32)
    rsd, osd = ref.state_dict(), ours.state_dict()
    assert list(rsd.keys()) == list(osd.keys())
    for k in rsd:
        assert torch.equal(rsd[k], osd[k]), k
    ours.load_state_dict(rsd)                                   # and checkpoints move both ways
    ref.load_state_dict(osd)


def test_ema_plan2_builder_shadows_and_weightnorm_rows_on_host():
    """Plan v2 (host only): plain / shadow / weight-norm chunks, the gain tensor travelling with its rows."""
    import struct
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    rows, dim = 200, 256                       # 200 rows of 256 -> 64 rows per chunk -> 4 weight-norm chunks
    numels = [1000, 40000, rows, rows * dim]   # bias, MLP weight (with shadow), weight_g, weight_v
    n = len(numels)
    arr = (L.i64 * n)(*numels)
    nbytes = lib.dmc_ema_plan2_bytes(arr, n, 3, dim)
    chunk = 72                                 # sizeof(EmaChunk2): 8 pointers/longs + 2 ints
    expect = 1 + 3 + 0 + 4                     # bias 1, weight 3 (40000 / 16384), weight_g none (rides with v), v 4
    assert nbytes == (expect + 1) * chunk        # capacity query: it does not know which tensor is the gain (one spare entry)
    tp = (L.vp * n)(0x10000, 0x20000, 0x30000, 0x4000000)
    sp = (L.vp * n)(0x110000, 0x120000, 0x130000, 0x5000000)
    sh = (L.vp * n)(None, 0x900000, None, None)
    buf = (C.c_uint8 * nbytes)()
    out = L.i64(0)
    rc = lib.dmc_ema_build_plan2(tp, sp, arr, sh, n, 3, 2, dim, 0x6000000, 0x7000000, 0x7100000, buf, nbytes, C.byref(out))
    assert rc == 0, lib.dmc_last_error_string()
    assert out.value == expect
    ent = [struct.unpack("<QQqQQQQQii", bytes(buf[i * chunk:(i + 1) * chunk])) for i in range(expect)]
    # (teacher, student, n, shadow, g_teacher, g_student, scale, inv_norm, kind, dim)
    assert ent[0] == (0x10000, 0x110000, 1000, 0, 0, 0, 0, 0, 0, 0)                         # plain
    assert ent[1][:4] == (0x20000, 0x120000, 16384, 0x900000) and ent[1][8] == 1            # shadow chunk 0
    assert ent[3][:4] == (0x20000 + 2 * 16384 * 4, 0x120000 + 2 * 16384 * 4, 40000 - 2 * 16384, 0x900000 + 2 * 16384 * 2)
    wn = ent[4:]
    assert [e[8] for e in wn] == [2, 2, 2, 2] and all(e[9] == dim for e in wn)
    assert [e[2] for e in wn] == [64 * dim, 64 * dim, 64 * dim, 8 * dim]
    assert wn[1][0] == 0x4000000 + 64 * dim * 4 and wn[1][3] == 0x6000000 + 64 * dim * 2        # v rows, bf16 operand rows
    assert wn[1][4] == 0x30000 + 64 * 4 and wn[1][5] == 0x130000 + 64 * 4                      # gains of those rows
    assert wn[1][6] == 0x7000000 + 64 * 4 and wn[1][7] == 0x7100000 + 64 * 4
    # validation: weight_g must have one entry per row; outputs must be given
    bad = (L.i64 * n)(1000, 40000, rows + 1, rows * dim)
    assert lib.dmc_ema_build_plan2(tp, sp, bad, sh, n, 3, 2, dim, 0x6000000, 0x7000000, 0x7100000, buf, nbytes, C.byref(out)) < 0
    assert lib.dmc_ema_build_plan2(tp, sp, arr, sh, n, 3, 2, dim, None, 0x7000000, 0x7100000, buf, nbytes, C.byref(out)) < 0
    # without a weight-normed layer it degenerates to plain + shadow chunks
    nb0 = lib.dmc_ema_plan2_bytes(arr, n, -1, 0)
    assert nb0 == (1 + 3 + 1 + 4) * chunk


def test_xrank_argument_validation_without_a_device():
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    assert lib.dmc_xrank_signal_bytes(8, 148) == (2 * 8 + 4) * 4
    assert lib.dmc_xrank_signal_bytes(0, 148) == 0
    peers = (L.vp * 2)(0x1000, 0x2000)
    pads = (L.vp * 2)(0x3000, 0x4000)
    args = dict(mc=None, n=1024, dt=L.DMC_BF16, rank=0, world=2, scale=0.5, ctas=8)

    def call(**kw):
        a = dict(args, **kw)
        return lib.dmc_xrank_allreduce(a["mc"], peers, pads, a["n"], a["dt"], a["rank"], a["world"], a["scale"], a["ctas"], 0, None, None,
                                       None, None)
    assert call(world=0) < 0 and call(world=17) < 0 and call(rank=2) < 0          # rank / world
    assert call(n=1001) < 0                                                      # not a multiple of 16 bytes
    assert call(dt=7) < 0 and call(ctas=0) < 0
    assert b"16 bytes" in lib.dmc_last_error_string() or True
    # the widening epilogue is for bf16 buffers only
    outs = (L.vp * 1)(0x5000)
    offs = (L.i64 * 1)(0)
    ns = (L.i64 * 1)(8)
    assert lib.dmc_xrank_allreduce(None, peers, pads, 1024, L.DMC_F32, 0, 2, 1.0, 8, 1, outs, offs, ns, None) < 0
    offs_bad = (L.i64 * 1)(4)                                                    # offsets must be multiples of 8 elements
    assert lib.dmc_xrank_allreduce(None, peers, pads, 1024, L.DMC_BF16, 0, 2, 1.0, 8, 1, outs, offs_bad, ns, None) < 0


def test_gemm_argument_validation_messages_without_a_device():
    """Every structural mistake in a dmc_gemm_args is refused on the host, before any CUDA call, with a negative code and a message
    naming it (the Python layer turns that into a RuntimeError)."""
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    buf = (C.c_char * 4096)()                      # any non-null, 16-byte aligned host address: validation never dereferences it
    base = (C.addressof(buf) + 15) & ~15

    def args(**kw):
        g = L.GemmArgs()
        g.M, g.N, g.K = 128, 128, 64
        g.A, g.B, g.D = base, base, base
        g.lda, g.ldb, g.ldd = 64, 64, 128
        g.in_dtype, g.out_dtype = L.DMC_BF16, L.DMC_F32
        for k, v in kw.items():
            setattr(g, k, v)
        return g

    def refused(g, text):
        rc = lib.dmc_gemm(C.byref(g), None)
        msg = lib.dmc_last_error_string().decode()
        assert rc < 0 and text in msg, (rc, msg)

    refused(args(K=0), "empty problem")
    refused(args(M=1 << 31), "dimension too large")
    refused(args(A=None), "null operand")
    refused(args(in_dtype=7), "bad in_dtype")
    refused(args(out_dtype=-1), "bad out_dtype")
    refused(args(act=99), "bad act")
    refused(args(act=L.ACT_GELU_DG), "need aux")
    refused(args(act=L.ACT_MUL_AUX), "need aux")
    refused(args(act=L.ACT_GELU_BWD), "needs aux")
    refused(args(act=L.ACT_NORMALIZE_BWD), "needs aux (fp32 rows) and row_scale")
    refused(args(act=L.ACT_NORMALIZE_BWD, aux=base, aux_dtype=L.DMC_F32, ldaux=128, row_scale=base, N=2048, ldd=2048), "N <= 1024")
    refused(args(act=L.ACT_NORMALIZE_BWD, aux=base, aux_dtype=L.DMC_F32, ldaux=128, row_scale=base, K=64), "at least two k-blocks")
    refused(args(A_lo=base), "A_lo and B_lo must be given together")
    refused(args(A_lo=base, B_lo=base), "require in_dtype F32")
    refused(args(split_k=2, K=256), "split-K needs a workspace")
    refused(args(split_k=2, K=256, workspace=base + 4, workspace_bytes=1 << 30), "workspace must be 16-byte aligned")
    refused(args(split_k=2, K=256, workspace=base, workspace_bytes=16), "split-K needs a workspace")
    assert lib.dmc_gemm(None, None) < 0 and b"null args" in lib.dmc_last_error_string()


def test_gemm_planner_workspace_invariants_on_host():
    """dmc_gemm_workspace_bytes exposes the planner's split-K decision: 0 (no split) or splits * M * N * 4 bytes with
    2 <= splits <= number of 128-byte k-blocks; wide outputs of short contractions are never split; the step's own shapes get
    the plans DESIGN.md section 4.1 describes."""
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    for dt, per_kb in ((L.DMC_BF16, 64), (L.DMC_F32, 32)):
        for M in (8, 128, 512, 2048, 4096):
            for N in (64, 256, 384, 2048, 65536):
                for K in (64, 256, 384, 2048, 65536):
                    nbytes = lib.dmc_gemm_workspace_bytes(M, N, K, dt)
                    assert nbytes % (M * N * 4) == 0, (M, N, K, dt)
                    splits = nbytes // (M * N * 4)
                    kb = -(-K // per_kb)
                    assert splits == 0 or 2 <= splits <= 3 * kb, (M, N, K, dt, splits)      # fp32 operands: up to 3 virtual passes
                    if dt == L.DMC_BF16 and (M // 128 or 1) * -(-N // 256) >= 148:
                        assert splits == 0, (M, N, K)                                       # a full wave of tiles: never split
    # the step's shapes at cfg2 (bf16): last-layer forward and the square MLP layers unsplit, last-layer dgrad split over CTA pairs
    assert lib.dmc_gemm_workspace_bytes(2048, 65536, 256, L.DMC_BF16) == 0
    assert lib.dmc_gemm_workspace_bytes(2048, 2048, 2048, L.DMC_BF16) == 0
    dgrad = lib.dmc_gemm_workspace_bytes(2048, 256, 65536, L.DMC_BF16) // (2048 * 256 * 4)
    assert 2 <= dgrad <= 18
    assert lib.dmc_gemm_workspace_bytes(0, 256, 256, L.DMC_BF16) == 0


def test_every_compute_entry_refuses_an_empty_call_on_the_host():
    """All-zero / all-null arguments: every int-returning entry point of the ABI validates before it touches CUDA and answers with a
    negative code and a message (no crash, no launch) -- the contract that lets the Python layer fail loudly."""
    import dinomc_b200
    L = dinomc_b200._lib
    lib = L.load()
    not_compute = {"dmc_version", "dmc_set_pdl", "dmc_set_streaming_ctas", "dmc_device_check"}
    checked = 0
    for name, (res, argtypes) in L.SIGNATURES.items():
        if res is not C.c_int or name in not_compute:
            continue
        args = []
        for t in argtypes:
            if t in (L.vp, C.c_char_p) or hasattr(t, "contents") or (isinstance(t, type) and issubclass(t, C._Pointer)):
                args.append(None)
            elif t in (C.c_float, C.c_double, L.f32):
                args.append(0.0)
            else:
                args.append(0)
        rc = getattr(lib, name)(*args)
        msg = lib.dmc_last_error_string().decode()
        assert rc < 0, (name, rc, msg)
        assert msg, name
        checked += 1
    assert checked == 34          # 50 entries - 12 size / count queries and string getters - 4 switches
