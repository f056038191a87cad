"""CPU-side checks of the C-ABI library (no compute calls without a GPU): it loads, exports every
symbol include/fa_b200.h declares, evaluates the attended pattern bit-exactly (same inline code
the kernels use), validates shapes like the reference OpKernels, and its tile schedule is
conservative."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import pattern
from tests.helpers import case_id, load_pattern_golden
from tf_flash_attention_b200 import _capi
from tf_flash_attention_b200 import flash_attention as fa

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CASES = load_pattern_golden()


def _problem(c, dtype=0, batch=(2, 3), d=8, v_d=5):
    qs, ks = tuple(c["q_shape"]), tuple(c["k_shape"])
    return _capi.make_problem(dtype, c["dims"], c["rule"], c["sync_mode"], batch + (d,) + qs,
                              batch + (d,) + ks, batch + (v_d,) + ks, c["window_size"],
                              c["log2_stride_size"], c["is_causal"])


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fa_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fa_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = C.CDLL(_capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/fa_b200.h but not exported"
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    assert b"sm_100a" in _capi.lib.fa_version()


@pytest.mark.parametrize("c", CASES, ids=case_id)
def test_pattern_bit_exact_vs_reference(c):
    p = _problem(c)
    assert np.array_equal(_capi.pattern_mask(p), c["mask"])
    ref, qo, ko = _capi.orders(p)
    assert ref == c["ref_shape"]
    assert qo.tolist() == c["q_order"] and ko.tolist() == c["k_order"]
    assert _capi.count_attended(p) == c["nnz"]


@pytest.mark.parametrize("tile", [(4, 4), (8, 16), (3, 5), (64, 64), (128, 128)])
def test_tile_schedule_is_conservative(tile):
    tq, tk = tile
    for c in CASES:
        cls = _capi.classify_tiles(_problem(c), tq, tk)
        m = c["mask"]
        for a in range(cls.shape[0]):
            for b in range(cls.shape[1]):
                blk = m[a * tq:(a + 1) * tq, b * tk:(b + 1) * tk]
                if cls[a, b] == 0:
                    assert not blk.any(), (case_id(c), a, b)
                elif cls[a, b] == 2:
                    assert blk.all(), (case_id(c), a, b)
                else:
                    assert cls[a, b] == 1


def test_tile_schedule_is_conservative_on_random_strided_windows():
    """Strided local windows (|delta| a multiple of 2^s): tiles whose delta interval holds no such multiple are skipped;
    SKIP must still imply that no pair attends, FULL that all do, and the skipping must actually happen."""
    rng = np.random.default_rng(7)
    skipped = live_partial = 0
    for _ in range(120):
        dims = int(rng.integers(1, 3))
        if dims == 1:
            qs, ks = (int(rng.integers(1, 400)),), (int(rng.integers(1, 400)),)
        else:
            qs = (int(rng.integers(1, 14)), int(rng.integers(1, 40)))
            ks = (int(rng.integers(1, 14)), int(rng.integers(1, 40)))
        mode = pattern.SYNC_MODES[int(rng.integers(0, 3))]
        w, st, cz = int(rng.integers(1, 40)), int(rng.integers(1, 9)), bool(rng.integers(0, 2))
        tq, tk = [(8, 16), (32, 32), (64, 64), (128, 64)][int(rng.integers(0, 4))]
        p = _capi.make_problem(0, dims, "local", mode, (1, 4) + qs, (1, 4) + ks, (1, 4) + ks, w, st, cz)
        cls = _capi.classify_tiles(p, tq, tk)
        m = pattern.tests_mask(qs, ks, mode, "local", w, st, cz)
        for a in range(cls.shape[0]):
            for b in range(cls.shape[1]):
                blk = m[a * tq:(a + 1) * tq, b * tk:(b + 1) * tk]
                if cls[a, b] == 0:
                    assert not blk.any(), (dims, qs, ks, mode, w, st, cz, tq, tk, a, b)
                    skipped += 1
                elif cls[a, b] == 2:
                    assert blk.all()
                else:
                    live_partial += 1
    assert skipped > live_partial // 4


def test_tile_schedule_prunes_baseline_configs():
    # C3: 150 of the 1024 128x128 tiles are live
    p = _capi.make_problem(0, 2, "local", "none_front", (1, 64, 64, 64), (1, 64, 64, 64), (1, 64, 64, 64), 8, 0, True)
    cls = _capi.classify_tiles(p, 128, 128)
    m = pattern.tests_mask((64, 64), (64, 64), "none_front", "local", 8, 0, True)
    live = sum(bool(m[a * 128:(a + 1) * 128, b * 128:(b + 1) * 128].any()) for a in range(32) for b in range(32))
    assert live == 150
    assert int((cls != 0).sum()) <= 160  # the conservative schedule keeps at most a few empty tiles
    # C2: causal 8192 -> lower triangle of 64x64 tiles, diagonal partial, rest full
    p = _capi.make_problem(0, 1, "causal", "none_front", (1, 128, 8192), (1, 128, 8192), (1, 128, 8192))
    cls = _capi.classify_tiles(p, 128, 128)
    assert np.array_equal(cls, np.tril(np.full((64, 64), 2), -1) + np.eye(64, dtype=int))
    assert _capi.count_attended(p) == 8192 * 8193 // 2


def test_count_attended_full_size():
    # C1 README example
    p = _capi.make_problem(1, 1, "local", "scale_front", (8, 32, 1024), (8, 32, 2048), (8, 16, 2048), 32, 0, False)
    assert _capi.count_attended(p) == 64016
    # C4 cross attention, scale_end
    p = _capi.make_problem(0, 1, "full", "scale_end", (4, 64, 1024), (4, 64, 8192), (4, 64, 8192))
    assert _capi.count_attended(p) == 1024 * 8192


def test_shape_validation_mirrors_reference():
    E = _capi.InvalidArgumentError
    ok = _capi.make_problem(0, 1, "full", "none_front", (2, 3, 8, 10), (2, 3, 8, 12), (2, 3, 5, 12))
    assert (ok.batch, ok.d, ok.v_d, ok.q_shape[0], ok.k_shape[0]) == (6, 8, 5, 10, 12)
    with pytest.raises(E) as e:  # ranks differ (forward.cc:100-101)
        _capi.make_problem(0, 1, "full", "none_front", (2, 8, 10), (2, 3, 8, 12), (2, 3, 5, 12))
    assert e.value.status == _capi.FA_EINVAL_RANK
    with pytest.raises(E) as e:  # rank < seq_dims + 2 (forward.cc:103-104)
        _capi.make_problem(0, 2, "full", "none_front", (8, 4, 4), (8, 4, 4), (8, 4, 4))
    assert e.value.status == _capi.FA_EINVAL_RANK
    with pytest.raises(E) as e:  # Q/K channels (forward.cc:126-127)
        _capi.make_problem(0, 1, "full", "none_front", (2, 8, 10), (2, 9, 12), (2, 5, 12))
    assert e.value.status == _capi.FA_EINVAL_CHANNEL
    with pytest.raises(E) as e:  # batch shapes (forward.cc:129-130)
        _capi.make_problem(0, 1, "full", "none_front", (2, 8, 10), (3, 8, 12), (2, 5, 12))
    assert e.value.status == _capi.FA_EINVAL_BATCH
    with pytest.raises(E) as e:  # K/V sequence shapes (forward.cc:132-133)
        _capi.make_problem(0, 1, "full", "none_front", (2, 8, 10), (2, 8, 12), (2, 5, 13))
    assert e.value.status == _capi.FA_EINVAL_SEQ_SHAPE
    with pytest.raises(E) as e:  # forward.cc:275-276
        _capi.make_problem(0, 1, "full", "scale_middle", (2, 8, 10), (2, 8, 12), (2, 5, 12))
    assert e.value.status == _capi.FA_EINVAL_SYNC_MODE and "Unsupported sync_mode: scale_middle" in str(e.value)
    # rule attributes are validated when the rule is compiled
    p = _capi.make_problem(0, 1, "local", "none_front", (2, 8, 10), (2, 8, 12), (2, 5, 12), 0, 0, False)
    n = C.c_int64()
    assert _capi.lib.fa_count_attended(C.byref(p), C.byref(n)) == _capi.FA_EINVAL_WINDOW
    p = _capi.make_problem(0, 1, "local", "none_front", (2, 8, 10), (2, 8, 12), (2, 5, 12), 4, 31, False)
    assert _capi.lib.fa_count_attended(C.byref(p), C.byref(n)) == _capi.FA_EINVAL_STRIDE
    p = _capi.make_problem(0, 1, "local", "none_front", (2, 8, 10), (2, 8, 12), (2, 5, 12), 1 << 20, 12, False)
    assert _capi.lib.fa_count_attended(C.byref(p), C.byref(n)) == _capi.FA_EINVAL_STRIDE


def test_backward_shape_validation():
    E = _capi.InvalidArgumentError
    p = _capi.make_problem(0, 1, "full", "none_front", (2, 8, 10), (2, 8, 12), (2, 5, 12))
    good = [(2, 8, 10), (2, 8, 12), (2, 5, 12), (2, 5, 10), (2, 10), (2, 10), (2, 5, 10)]
    _capi.check_backward_shapes(p, good)
    bad = list(good); bad[4] = (2, 1, 10)       # l rank (backward.cc:201-203)
    with pytest.raises(E) as e:
        _capi.check_backward_shapes(p, bad)
    assert e.value.status == _capi.FA_EINVAL_RANK
    bad = list(good); bad[3] = (2, 6, 10)       # V/O channels (backward.cc:245-246)
    with pytest.raises(E) as e:
        _capi.check_backward_shapes(p, bad)
    assert e.value.status == _capi.FA_EINVAL_CHANNEL
    bad = list(good); bad[6] = (2, 5, 11)       # sequence shape of dO (backward.cc:257-258)
    with pytest.raises(E) as e:
        _capi.check_backward_shapes(p, bad)
    assert e.value.status == _capi.FA_EINVAL_SEQ_SHAPE
    bad = list(good); bad[5] = (3, 10)          # batch of m (backward.cc:249-252)
    with pytest.raises(E) as e:
        _capi.check_backward_shapes(p, bad)
    assert e.value.status == _capi.FA_EINVAL_BATCH


def test_python_api_surface_matches_reference():
    import inspect
    names = ("full_1d", "causal_1d", "local_1d", "full_2d", "causal_2d", "local_2d")
    # the reference's positional surface; `layout` is this package's one addition and is keyword-only
    sig = {n: [k for k, v in inspect.signature(getattr(fa, n)).parameters.items() if v.kind != v.KEYWORD_ONLY]
           for n in names}
    for n in names:
        extra = [k for k, v in inspect.signature(getattr(fa, n)).parameters.items() if v.kind == v.KEYWORD_ONLY]
        assert extra == ["layout"] and inspect.signature(getattr(fa, n)).parameters["layout"].default == "channel_first"
    assert sig["full_1d"] == sig["full_2d"] == ["Q", "K", "V", "sync_mode", "returning_l_m"]
    assert sig["causal_1d"] == sig["causal_2d"] == ["Q", "K", "V", "sync_mode", "returning_l_m"]
    assert sig["local_1d"] == sig["local_2d"] == ["Q", "K", "V", "window_size", "log2_stride_size", "is_causal",
                                                  "sync_mode", "returning_l_m"]
    assert inspect.signature(fa.full_1d).parameters["sync_mode"].default == "none_front"
    assert inspect.signature(fa.causal_1d).parameters["sync_mode"].default is inspect.Parameter.empty


def test_ring_driver_argument_checks():
    """fa_ring_causal_forward / _backward / arena sizes (include/fa_b200.h): validated before anything is queued."""
    lib = _capi.lib
    one = 256
    assert lib.fa_ring_causal_arena_bytes(9, 4, 64, 64, 128, 1, 0) == 0
    assert lib.fa_ring_causal_arena_bytes(_capi.FA_F16, 0, 64, 64, 128, 1, 0) == 0
    fwd1 = lib.fa_ring_causal_arena_bytes(_capi.FA_F16, 4, 64, 64, 128, 1, 0)
    fwd2 = lib.fa_ring_causal_arena_bytes(_capi.FA_F16, 4, 64, 64, 128, 2, 0)
    acc = 2 * 4 * 128 * (64 + 2) * 4                 # fp32 O, l, m accumulators of both local chunks
    part = 2 * 4 * 128 * (64 * 2 + 4 + 2)            # the step's own O, l, m
    assert fwd1 >= acc + part
    assert fwd2 >= fwd1 + 3 * 2 * 4 * 64 * 128 * 2   # + [Q_hi; Q_hi], [K_lo; K_lo], [V_lo; V_lo]
    bwd2 = lib.fa_ring_causal_arena_bytes(_capi.FA_F16, 4, 64, 64, 128, 2, 1)
    assert bwd2 >= 3 * 2 * 4 * 64 * 128 * 4 + 3 * 2 * 4 * 64 * 128 * 2
    assert lib.fa_ring_causal_slot_bytes(_capi.FA_F16, 4, 64, 32, 128, 0) == 2 * 4 * 128 * (64 + 32) * 2
    assert lib.fa_ring_causal_slot_bytes(_capi.FA_F16, 4, 64, 32, 128, 1) == 2 * 4 * 128 * (64 + 32) * 4
    f = lambda dtype=_capi.FA_F16, q=one, arena=one, nbytes=fwd1: lib.fa_ring_causal_forward(  # noqa: E731
        None, dtype, 4, 64, 64, 128, q, one, one, one, one, one, arena, nbytes, None)
    assert f(dtype=5) == _capi.FA_EINVAL_DTYPE
    assert f(q=None) == _capi.FA_EINVAL_NULL
    assert f(arena=None) == _capi.FA_EINVAL_NULL
    assert f(nbytes=fwd1 - 1) == _capi.FA_EINVAL_WORKSPACE
    assert f(arena=one + 16) == _capi.FA_EINVAL_WORKSPACE          # arena must be 256-byte aligned
    b = lib.fa_ring_causal_backward(None, None, _capi.FA_F16, 4, 64, 64, 128, one, one, one, one, one, one, one, one, one,
                                    None, one, 1 << 30, None)
    assert b == _capi.FA_EINVAL_NULL


def test_dispatch_path_names_the_family_and_the_reason():
    """fa_dispatch_path (host only): no silent cliffs - a caller can ask which family a problem takes and why the
    tensor-core kernels decline it (sequences beyond the 2048-tile schedule dropped to SIMT with
    fa_last_path() as the only trace)."""
    import ctypes as C
    dp = fa.dispatch_path
    assert dp(1, "causal", (4, 128, 8192), (4, 128, 8192), (4, 128, 8192), np.float16) == ("tcgen05_f16", "")
    assert dp(1, "causal", (4, 24, 342), (4, 24, 342), (4, 9, 342), np.float16, backward=True) == ("tcgen05_f16", "")
    assert dp(2, "local", (2, 17, 13, 9), (2, 17, 13, 9), (2, 31, 13, 9), np.float32, window_size=3) == ("tcgen05_f32_split", "")
    assert dp(1, "full", (2, 64, 100), (2, 64, 100), (2, 64, 100), np.float64, backward=True) == ("dmma_f64", "")
    # the schedule holds 2048 streamed tiles: 64-key tiles for head_dim <= 64, 128-key tiles for the head_dim-128 forward
    assert dp(1, "full", (1, 64, 64), (1, 64, 131072), (1, 64, 131072), np.float16)[0] == "tcgen05_f16"
    name, why = dp(1, "full", (1, 64, 64), (1, 64, 131072 + 64), (1, 64, 131072 + 64), np.float16)
    assert name == "generic_simt" and "2048" in why and "ring" in why
    assert dp(1, "full", (1, 128, 64), (1, 128, 262144), (1, 128, 262144), np.float16)[0] == "tcgen05_f16"
    assert dp(1, "full", (1, 128, 64), (1, 128, 262144 + 128), (1, 128, 262144 + 128), np.float16)[0] == "generic_simt"
    name, why = dp(1, "full", (1, 128, 64), (1, 128, 131072 + 64), (1, 128, 131072 + 64), np.float16, backward=True)
    assert name == "generic_simt" and "2048" in why
    assert dp(1, "full", (1, 160, 64), (1, 160, 64), (1, 160, 64), np.float16) == ("generic_simt", "more than 128 channels: generic kernels")
    assert dp(1, "full", (1, 96, 64), (1, 96, 64), (1, 96, 64), np.float32)[1].startswith("more than 64 channels")
    assert dp(1, "full", (1, 96, 64), (1, 96, 64), (1, 96, 64), np.float64, backward=True)[0] == "generic_simt"
    p = _capi.make_problem(_capi.FA_F16, 1, "full", "none_front", (1, 64, 64), (1, 64, 64), (1, 64, 64))
    p.accumulate = 1
    assert _capi.dispatch_path(p) == ("generic_simt", "accumulate = 1 is served by the generic forward kernels")
    assert _capi.dispatch_path(p, backward=True)[0] == "tcgen05_f16"
    p.accumulate = 0
    p.layout, p.heads = _capi.FA_LAYOUT_CHANNEL_LAST, 1
    assert _capi.dispatch_path(p)[0] == "tcgen05_f16"
    p32 = _capi.make_problem(_capi.FA_F32, 1, "full", "none_front", (1, 64, 64), (1, 64, 64), (1, 64, 64))
    p32.layout, p32.heads = _capi.FA_LAYOUT_CHANNEL_LAST, 1
    with pytest.raises(_capi.InvalidArgumentError) as e:
        _capi.dispatch_path(p32)
    assert e.value.status == _capi.FA_EINVAL_LAYOUT
    try:
        _capi.lib.fa_set_path_override(1)
        assert dp(1, "causal", (4, 128, 512), (4, 128, 512), (4, 128, 512), np.float16)[0] == "generic_simt"
    finally:
        _capi.lib.fa_set_path_override(0)
    assert _capi.lib.fa_dispatch_path(None, 0, None, 0) == _capi.FA_EINVAL_NULL
    assert _capi.lib.fa_dispatch_path(C.byref(p), 0, None, 0) == 2            # the reason buffer is optional


def test_backward_accumulate_probe_and_argument_checks():
    """fa_backward_accumulate_supported is a host-only probe: the fused head_dim-128 fp16 kernel on unpacked channel-first
    tensors is the only path that adds dQ into an accumulator; everything else answers no (and the entry point
    FA_EINVAL_SHAPE), so callers keep fa_backward + fa_grad_accumulate as the fallback."""
    import ctypes as C
    sup = _capi.lib.fa_backward_accumulate_supported
    mk = lambda dt, d, vd, nq, nk, rule="causal": _capi.make_problem(dt, 1, rule, "none_front", (4, d, nq), (4, d, nk), (4, vd, nk))  # noqa: E731
    assert sup(C.byref(mk(_capi.FA_F16, 128, 128, 512, 512)), 0) == 1
    assert sup(C.byref(mk(_capi.FA_F16, 128, 64, 512, 384, "full")), 2) == 1
    assert sup(C.byref(mk(_capi.FA_F16, 128, 128, 512, 512)), 5) == 0       # fold larger than the batch
    assert sup(C.byref(mk(_capi.FA_F16, 128, 128, 500, 512)), 0) == 0       # packed lengths
    assert sup(C.byref(mk(_capi.FA_F16, 64, 64, 512, 512)), 0) == 0         # two-kernel backward
    assert sup(C.byref(mk(_capi.FA_F16, 128, 96, 512, 512)), 0) == 0        # value channels padded by the descriptors
    assert sup(C.byref(mk(_capi.FA_F32, 64, 64, 512, 512)), 0) == 0
    p = mk(_capi.FA_F16, 128, 128, 512, 512)
    p.layout, p.heads = _capi.FA_LAYOUT_CHANNEL_LAST, 2
    assert sup(C.byref(p), 0) == 0
    try:
        _capi.lib.fa_set_grad_precision(1)                                   # precise everywhere: two-kernel backward
        assert sup(C.byref(mk(_capi.FA_F16, 128, 128, 512, 512)), 0) == 0
        _capi.lib.fa_set_grad_precision(0)
        _capi.lib.fa_set_path_override(4)
        assert sup(C.byref(mk(_capi.FA_F16, 128, 128, 512, 512)), 0) == 0
    finally:
        _capi.lib.fa_set_grad_precision(0)
        _capi.lib.fa_set_path_override(0)
    one = 256
    q = mk(_capi.FA_F16, 64, 64, 512, 512)
    ws = _capi.lib.fa_workspace_bytes(C.byref(q), 1)
    call = lambda pr, acc=one, nbytes=1 << 30: _capi.lib.fa_backward_accumulate(  # noqa: E731
        C.byref(pr), one, one, one, one, one, one, one, acc, one, one, 0, one, nbytes, None)
    assert call(q) == _capi.FA_EINVAL_SHAPE
    assert call(q, acc=None) == _capi.FA_EINVAL_NULL
    assert call(q, nbytes=ws - 1) == _capi.FA_EINVAL_WORKSPACE


def test_channel_last_problem_validation():
    """fa_problem_t.layout / heads (include/fa_b200.h): validated on the host before any launch; a dtype no kernel
    reads channel-last answers FA_EINVAL_LAYOUT so the caller can run the adapter (fa_layout_transpose) instead."""
    import ctypes as C
    p = _capi.make_problem(_capi.FA_F16, 1, "causal", "none_front", (2, 3, 64, 128), (2, 3, 64, 128), (2, 3, 64, 128))
    assert p.layout == _capi.FA_LAYOUT_CHANNEL_FIRST and p.heads == 0
    one = C.c_void_p(256)
    call = lambda: _capi.lib.fa_forward(C.byref(p), one, one, one, one, one, one, None, 0, None)  # noqa: E731
    p.layout = 2
    assert call() == _capi.FA_EINVAL_LAYOUT
    p.layout, p.heads = _capi.FA_LAYOUT_CHANNEL_LAST, 0
    assert call() == _capi.FA_EINVAL_LAYOUT
    p.heads = 4                                      # 6 batch elements are not a multiple of 4 heads
    assert call() == _capi.FA_EINVAL_LAYOUT
    assert "channel-last" in _capi.lib.fa_strerror(_capi.FA_EINVAL_LAYOUT).decode()
    p32 = _capi.make_problem(_capi.FA_F32, 1, "causal", "none_front", (2, 3, 64, 128), (2, 3, 64, 128), (2, 3, 64, 128))
    p32.layout, p32.heads = _capi.FA_LAYOUT_CHANNEL_LAST, 3
    ws = _capi.lib.fa_workspace_bytes(C.byref(p32), 0)
    assert _capi.lib.fa_forward(C.byref(p32), one, one, one, one, one, one, one, ws, None) == _capi.FA_EINVAL_LAYOUT
    # the host-buffer entry points chunk the batch: channel-first only
    p.heads = 3
    assert _capi.lib.fa_host_arena_bytes(C.byref(p), 0) == 0
    assert _capi.lib.fa_forward_host(C.byref(p), one, one, one, one, one, one, one, 1 << 30, None) == _capi.FA_EINVAL_LAYOUT
    # python side: shapes are outer + sequence + (heads, channels)
    from tf_flash_attention_b200.flash_attention import _cf_shape
    assert _cf_shape((2, 100, 3, 64), 1) == (2, 3, 64, 100)
    assert _cf_shape((5, 2, 10, 12, 3, 64), 2) == (5, 2, 3, 64, 10, 12)
    with pytest.raises(_capi.InvalidArgumentError):
        fa.full_1d(np.zeros((1, 8, 2, 8), np.float16), np.zeros((1, 8, 2, 8), np.float16),
                   np.zeros((1, 8, 2, 8), np.float16), layout="channel_last")     # host arrays: channel-first only
    with pytest.raises(_capi.InvalidArgumentError):
        fa.full_1d(np.zeros((1, 8, 8), np.float16), np.zeros((1, 8, 8), np.float16), np.zeros((1, 8, 8), np.float16),
                   layout="rows")


def test_layout_adapter_argument_checks():
    """fa_layout_transpose validates before touching the device: dtype code, negative sizes, null pointers; empty
    tensors are a no-op."""
    f = _capi.lib.fa_layout_transpose
    assert f(7, None, None, 1, 1, 1, 1, 1, None) == _capi.FA_EINVAL_DTYPE
    assert f(0, None, None, -1, 1, 1, 1, 1, None) == _capi.FA_EINVAL_SHAPE
    assert f(0, None, None, 1, 1, 1, 1, 1, None) == _capi.FA_EINVAL_NULL
    assert f(1, None, None, 4, 0, 2, 8, 0, None) == _capi.FA_OK
    with pytest.raises(_capi.InvalidArgumentError):
        fa.from_channel_last(np.zeros((1, 2, 3, 4), dtype=np.float32))   # host arrays: no CPU fallback


def test_layout_fast_path_mapping():
    """Host emulation of the thread mapping of layout_transpose_f16_kernel (fa_layout.cu): every 8-byte unit of the
    64 x 64 tile is written once, read once by the thread that stores it to the transposed position, and neither the unit
    stores nor the unit loads of any half-warp collide on a shared-memory bank (16 banks of 8 bytes)."""
    smem = {}
    for tid in range(256):
        kq, rq = tid % 16, tid // 16
        for c in range(4):
            addr = (4 * kq + c) * 16 + (rq ^ kq)
            assert addr not in smem
            smem[addr] = (4 * kq + c, rq)            # destination row, group of 4 source rows
    assert len(smem) == 1024
    for w in range(8):
        for c in range(4):
            for hw in range(2):
                lanes = [w * 32 + hw * 16 + i for i in range(16)]
                assert len({((4 * (t % 16) + c) * 16 + ((t // 16) ^ (t % 16))) % 16 for t in lanes}) == 16
    seen = set()
    for w in range(8):
        for i in range(2):
            for hw in range(2):
                for which in range(2):
                    banks = set()
                    for lane in range(hw * 16, hw * 16 + 16):
                        u, rsel = lane % 8, lane // 8
                        rho = 16 * (w >> 1) + 4 * rsel + 2 * (w & 1) + i
                        addr = rho * 16 + ((2 * u + which) ^ (rho >> 2))
                        assert smem[addr] == (rho, 2 * u + which)
                        banks.add(addr % 16)
                        seen.add((rho, 2 * u + which))
                    assert len(banks) == 16
    assert len(seen) == 1024


def test_no_cpu_fallback():
    """numpy (host) inputs must fail loudly when no CUDA device is present."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    x = np.zeros((1, 8, 16), dtype=np.float32)
    with pytest.raises(_capi.FlashAttentionError):
        fa.full_1d(x, x, x)


def test_estimate_forward_flops_matches_reference():
    """fa_estimate_forward_flops against the reference's own EstimateForwardFlops (golden values
    produced by tests/golden/make_flops_golden.py from oracle/_ref/libref_fa.so): bit-identical."""
    import json
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "flops_golden.json")))
    assert len(cases) >= 60
    for c in cases:
        qs, ks = tuple(c["q_shape"]), tuple(c["k_shape"])
        p = _capi.make_problem(c["dtype"], c["dims"], c["rule"], c["sync_mode"], (c["batch"], c["d"]) + qs,
                               (c["batch"], c["d"]) + ks, (c["batch"], c["v_d"]) + ks, c["window_size"],
                               c["log2_stride_size"], c["is_causal"])
        assert _capi.estimate_forward_flops(p, c["smem"]) == c["flops"], c
    dt = {0: np.float16, 1: np.float32, 2: np.float64}
    c = cases[0]
    got = fa.estimate_forward_flops(c["dims"], c["rule"], (c["batch"], c["d"]) + tuple(c["q_shape"]),
                                    (c["batch"], c["d"]) + tuple(c["k_shape"]),
                                    (c["batch"], c["v_d"]) + tuple(c["k_shape"]), dt[c["dtype"]], c["sync_mode"],
                                    c["window_size"], c["log2_stride_size"], c["is_causal"])
    assert got == c["flops"]


@pytest.mark.parametrize("tile,resident_is_q", [(64, 1), (128, 1), (64, 0)])
def test_closed_form_tile_masks_match_reference(tile, resident_is_q):
    """The 32-column masks the tcgen05 kernels build per PARTIAL tile (fa_fast_mask32: intervals per grid row, and for
    strided windows the multiples of 2^s inside them, instead of 32 evaluations of the element rule) reproduce the
    reference pattern bit for bit, with queries resident (forward / dQ kernels) and with keys resident (dK/dV kernel)."""
    n = 0
    for c in CASES:
        assert np.array_equal(_capi.pattern_mask_fast(_problem(c), tile, resident_is_q), c["mask"]), case_id(c)
        n += 1
    assert n >= 70
    rng = np.random.default_rng(tile + resident_is_q)
    for _ in range(150):
        dims = int(rng.integers(1, 3))
        if dims == 1:
            qs, ks = (int(rng.integers(1, 300)),), (int(rng.integers(1, 300)),)
        else:
            qs = (int(rng.integers(1, 12)), int(rng.integers(1, 70)))
            ks = (int(rng.integers(1, 12)), int(rng.integers(1, 70)))
        rule = ["full", "causal", "local"][int(rng.integers(0, 3))]
        mode = pattern.SYNC_MODES[int(rng.integers(0, 3))]
        w, cz = int(rng.integers(1, 10)), bool(rng.integers(0, 2))
        st = int(rng.integers(0, 5)) if rule == "local" else 0      # strided windows: |delta| a multiple of 2^st
        p = _capi.make_problem(0, dims, rule, mode, (1, 4) + qs, (1, 4) + ks, (1, 4) + ks, w, st, cz)
        assert np.array_equal(_capi.pattern_mask_fast(p, tile, resident_is_q),
                              pattern.tests_mask(qs, ks, mode, rule, w, st, cz)), (dims, qs, ks, rule, mode, w, st, cz)


def test_workspace_sizes_follow_the_dispatch_rules():
    """fa_workspace_bytes is host-only. What the op author must allocate:
    fp16 head_dim 128 backward = statistics + fp32 dQ scratch (fused kernel); with override 4 (two-kernel backward) and for
    head_dim 64 = statistics only; fp32 forward = hi/lo TF32 copies when the tcgen05 forward takes the shape; fp32 backward
    = three bf16 pieces of Q, K, V, dO for the supported head dims; everything else = the generic kernels' statistics."""
    import ctypes as C

    def ws(dtype, d, vd, nq, nk, bwd, batch=(2, 3)):
        p = _capi.make_problem(dtype, 1, "causal", "none_front", batch + (d, nq), batch + (d, nk), batch + (vd, nk))
        return int(_capi.lib.fa_workspace_bytes(C.byref(p), bwd))
    B = 6
    stats16 = lambda nq: (2 * (B * nq + 64) * 4 + 255) // 256 * 256  # noqa: E731
    try:
        assert ws(_capi.FA_F16, 128, 128, 512, 512, 0) == 0
        fused = ws(_capi.FA_F16, 128, 128, 512, 384, 1)
        assert fused >= stats16(512) + B * 128 * 512 * 4                      # + fp32 dQ scratch
        assert ws(_capi.FA_F16, 64, 64, 512, 384, 1) < B * 64 * 512 * 4       # head_dim 64: statistics only
        _capi.lib.fa_set_path_override(4)
        assert ws(_capi.FA_F16, 128, 128, 512, 384, 1) < B * 128 * 512 * 4    # two-kernel backward: no scratch
        _capi.lib.fa_set_path_override(0)
        assert ws(_capi.FA_F32, 64, 64, 512, 384, 0) >= 2 * 4 * B * 64 * (512 + 2 * 384)   # hi + lo of Q, K, V
        # 48 channels ride the 64-channel 3xTF32 kernel: padded hi/lo copies of Q, K, V and a padded O
        assert ws(_capi.FA_F32, 48, 48, 512, 384, 0) >= 2 * 4 * B * 64 * (512 + 2 * 384) + 4 * B * 64 * 512
        assert ws(_capi.FA_F32, 72, 72, 512, 384, 0) == 0                     # above 64 channels: generic kernels
        pieces = 3 * 2 * B * (64 * 512 * 2 + 64 * 384 * 2)                    # 3 bf16 pieces of Q, dO, K, V
        assert ws(_capi.FA_F32, 64, 64, 512, 384, 1) >= pieces
        assert ws(_capi.FA_F32, 32, 16, 512, 384, 1) >= 3 * 2 * B * (32 * 512 + 16 * 512 + 32 * 384 + 16 * 384)
        assert ws(_capi.FA_F32, 48, 48, 512, 384, 1) >= pieces                # padded to the 64-channel kernel
        assert ws(_capi.FA_F32, 72, 72, 512, 384, 1) == 2 * B * 512 * 4       # generic: LSE + D
        assert ws(_capi.FA_F64, 64, 64, 512, 384, 1) == 2 * B * 512 * 8
        _capi.lib.fa_set_path_override(1)                                     # generic family only
        assert ws(_capi.FA_F32, 64, 64, 512, 384, 1) == 2 * B * 512 * 4
        assert ws(_capi.FA_F32, 64, 64, 512, 384, 0) == 0
    finally:
        _capi.lib.fa_set_path_override(0)


def test_host_step_entry_points_validate_before_touching_the_device():
    """fa_forward_backward_host / fa_backward_host_resident / fa_step_host_arena_bytes: sizes and argument checks are
    host-only."""
    import ctypes as C
    lib = _capi.lib
    for dtype, d in ((_capi.FA_F16, 128), (_capi.FA_F32, 64), (_capi.FA_F64, 24)):
        p = _capi.make_problem(dtype, 1, "causal", "none_front", (2, 3, d, 512), (2, 3, d, 384), (2, 3, d, 384))
        step = int(lib.fa_step_host_arena_bytes(C.byref(p)))
        fwd, bwd = int(lib.fa_host_arena_bytes(C.byref(p), 0)), int(lib.fa_host_arena_bytes(C.byref(p), 1))
        assert step >= bwd >= fwd > 0
        # the step arena holds the backward layout plus the larger of the two workspaces
        assert step - bwd == max(0, (int(lib.fa_workspace_bytes(C.byref(p), 0)) + 255) // 256 * 256
                                 - (int(lib.fa_workspace_bytes(C.byref(p), 1)) + 255) // 256 * 256)
    one = C.c_void_p(0x1000)    # never dereferenced: every call below fails validation first
    args = [one] * 10
    assert lib.fa_forward_backward_host(C.byref(p), *args, None, step, None) == _capi.FA_EINVAL_WORKSPACE
    assert lib.fa_forward_backward_host(C.byref(p), *args, one, step - 1, None) == _capi.FA_EINVAL_WORKSPACE
    assert lib.fa_forward_backward_host(C.byref(p), None, *args[1:], one, step, None) == _capi.FA_EINVAL_NULL
    assert lib.fa_backward_host_resident(C.byref(p), None, one, one, one, one, bwd, None) == _capi.FA_EINVAL_NULL
    assert lib.fa_backward_host_resident(C.byref(p), one, one, one, one, one, bwd - 1, None) == _capi.FA_EINVAL_WORKSPACE
    p.accumulate = 1
    assert lib.fa_forward_backward_host(C.byref(p), *args, one, step, None) == _capi.FA_EINVAL_SHAPE
    p.accumulate = 0
    p.batch = 0                 # empty batch: nothing to do
    assert lib.fa_forward_backward_host(C.byref(p), *args, one, 1 << 30, None) == _capi.FA_OK
    assert lib.fa_backward_host_resident(C.byref(p), one, one, one, one, one, 1 << 30, None) == _capi.FA_OK


def test_round2_ab_script_child_compiles():
    """tools/ab_fwd_variants.py runs its per-variant body in a child interpreter: the template must at least compile."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "ab_fwd_variants.py")
    spec = importlib.util.spec_from_file_location("ab_fwd_variants", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    compile(mod.CHILD % {"root": "/tmp"}, "child", "exec")


def test_tensor_core_paths_decline_sequences_beyond_the_tile_schedule():
    """The per-CTA tile schedule covers 2048 streamed 64-wide tiles (131072 positions); longer sequences must fall back to
    the generic kernels instead of overrunning it. Observable on the host through the fp32 workspace sizes."""
    import ctypes as C

    def ws(nq, nk, bwd):
        p = _capi.make_problem(_capi.FA_F32, 1, "causal", "none_front", (1, 64, nq), (1, 64, nk), (1, 64, nk))
        return int(_capi.lib.fa_workspace_bytes(C.byref(p), bwd))
    assert ws(1024, 131072, 0) > 0 and ws(1024, 131072 + 64, 0) == 0
    assert ws(1024, 131072, 1) > 2 * 1024 * 4 * 10
    assert ws(1024, 131072 + 64, 1) == 2 * 1024 * 4          # generic: LSE + D only
    assert ws(131072 + 64, 1024, 1) == 2 * (131072 + 64) * 4


def test_null_workspace_is_rejected_when_one_is_needed():
    """fa_forward / fa_backward with workspace == NULL while fa_workspace_bytes() > 0 must fail validation instead of
    launching kernels that write through address 0 (the fp32 forward keeps its TF32 pieces there)."""
    import ctypes as C
    lib = _capi.lib
    p = _capi.make_problem(_capi.FA_F32, 1, "full", "none_front", (2, 64, 256), (2, 64, 512), (2, 64, 512))
    need_f, need_b = int(lib.fa_workspace_bytes(C.byref(p), 0)), int(lib.fa_workspace_bytes(C.byref(p), 1))
    assert need_f > 0 and need_b > 0
    one = C.c_void_p(0x1000)    # never dereferenced: validation fails first
    assert lib.fa_forward(C.byref(p), one, one, one, one, one, one, None, need_f, None) == _capi.FA_EINVAL_WORKSPACE
    assert lib.fa_forward(C.byref(p), one, one, one, one, one, one, one, need_f - 1, None) == _capi.FA_EINVAL_WORKSPACE
    assert lib.fa_backward(C.byref(p), *([one] * 10), None, need_b, None) == _capi.FA_EINVAL_WORKSPACE


def test_ring_data_plane_argument_checks_and_no_device():
    """fa_ring_* (csrc/fa_ring.cu): argument validation is host-only; without a CUDA device fa_ring_create fails with a
    status instead of crashing (there is no CPU fallback for the data plane either)."""
    import ctypes as C
    lib = _capi.lib
    blob = C.create_string_buffer(_capi.FA_RING_HANDLE_BYTES)
    h = C.c_void_p()
    assert lib.fa_ring_create(0, 2, 1024, 2, None, blob) == _capi.FA_EINVAL_NULL
    assert lib.fa_ring_create(0, 2, 1024, 2, C.byref(h), None) == _capi.FA_EINVAL_NULL
    assert lib.fa_ring_create(2, 2, 1024, 2, C.byref(h), blob) == _capi.FA_EINVAL_SHAPE      # rank out of range
    assert lib.fa_ring_create(0, 2, 0, 2, C.byref(h), blob) == _capi.FA_EINVAL_SHAPE         # empty slots
    assert lib.fa_ring_create(0, 2, 1024, 9, C.byref(h), blob) == _capi.FA_EINVAL_SHAPE      # too many slots
    assert lib.fa_ring_send(None, 0, 1, None, None, -1, None) == _capi.FA_EINVAL_NULL
    assert lib.fa_ring_recv_wait(None, 0, None) == _capi.FA_EINVAL_NULL
    assert lib.fa_ring_destroy(None) == _capi.FA_OK
    import torch
    if not torch.cuda.is_available():
        rc = lib.fa_ring_create(0, 2, 1024, 2, C.byref(h), blob)
        assert rc in (_capi.FA_ECUDA, _capi.FA_ENODEVICE)
