"""channel-last operands ([outer, sequence..., heads, channels], what a projection produces)
read and written directly through the kernels' 4-D TMA descriptors. The channel-last kernels issue the same products
in the same order as the channel-first ones (only the operand descriptors flip between MN-major and K-major), so the
results must be bit-identical to `to_channel_last(op(from_channel_last(x)))`, which is itself pinned to the oracle by
tests/test_gpu_parity.py / test_gpu_sm100.py."""
import zlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tf_flash_attention_b200 import _capi, flash_attention as fa  # noqa: E402


def _cl(x, sd):   # torch-side permutation (test plumbing): outer + (h, c) + seq -> outer + seq + (h, c)
    n = x.dim()
    perm = list(range(n - sd - 2)) + list(range(n - sd, n)) + [n - sd - 2, n - sd - 1]
    return x.permute(perm).contiguous()


def _cf(x, sd):
    n = x.dim()
    perm = list(range(n - sd - 2)) + [n - 2, n - 1] + list(range(n - sd - 2, n - 2))
    return x.permute(perm).contiguous()


def _call(dims, rule, Q, K, V, mode, w, s, c, **kw):
    if rule == "full":
        f = fa.full_1d if dims == 1 else fa.full_2d
        return f(Q, K, V, mode, returning_l_m=True, **kw)
    if rule == "causal":
        f = fa.causal_1d if dims == 1 else fa.causal_2d
        return f(Q, K, V, mode, returning_l_m=True, **kw)
    f = fa.local_1d if dims == 1 else fa.local_2d
    return f(Q, K, V, w, s, c, mode, returning_l_m=True, **kw)


CASES = [
    # dims rule mode w s c outer heads d vd q k
    (1, "causal", "none_front", 1, 0, False, (2,), 3, 128, 128, (1000,), (1000,)),   # 128-key tiles, ragged tail
    (1, "full", "scale_end", 1, 0, False, (1,), 2, 64, 64, (300,), (515,)),          # 64-key tiles, two CTAs per SM
    (1, "causal", "scale_front", 1, 0, False, (2,), 1, 128, 64, (256,), (384,)),
    (1, "local", "none_front", 40, 0, True, (1,), 4, 64, 128, (200,), (333,)),
    (1, "full", "none_front", 1, 0, False, (2, 2), 3, 32, 16, (77,), (131,)),        # channels padded by the descriptors
    (1, "causal", "none_front", 1, 0, False, (1,), 5, 8, 8, (129,), (129,)),
    (2, "local", "none_front", 3, 0, True, (2,), 2, 24, 40, (13, 9), (13, 9)),
    (2, "causal", "scale_front", 1, 0, False, (1,), 2, 64, 64, (9, 14), (18, 14)),
    (1, "local", "none_front", 2, 0, False, (1,), 2, 16, 16, (200,), (40,)),         # rows without keys
    (1, "full", "none_front", 1, 0, False, (1,), 1, 96, 104, (1,), (1,)),
]


def _inputs(case, dtype=torch.float16):
    dims, rule, mode, w, s, c, outer, heads, d, vd, qs, ks = case
    g = torch.Generator(device="cuda").manual_seed(zlib.crc32(repr(case).encode()) % 10000)
    mk = lambda ch, seq: (torch.rand(outer + (heads, ch) + seq, generator=g, device="cuda") * 4 - 2).to(dtype)  # noqa: E731
    return mk(d, qs), mk(d, ks), mk(vd, ks), mk(vd, qs)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-{c[1]}-{c[2]}-h{c[7]}-d{c[8]}x{c[9]}-q{'x'.join(map(str, c[10]))}")
def test_forward_channel_last_is_bit_identical_and_direct(case):
    dims, rule, mode, w, s, c = case[:6]
    Q, K, V, _ = _inputs(case)
    O, l, m = _call(dims, rule, Q, K, V, mode, w, s, c)
    Ql, Kl, Vl = (_cl(x, dims) for x in (Q, K, V))
    torch.cuda.synchronize()
    _capi.lib.fa_launch_count(1)
    Ol, ll, ml = _call(dims, rule, Ql, Kl, Vl, mode, w, s, c, layout="channel_last")
    torch.cuda.synchronize()
    assert _capi.lib.fa_launch_count(1) == 1, "channel-last fp16 forward must be ONE kernel (no adapter / pack pass)"
    assert _capi.lib.fa_last_path() == 2
    assert Ol.shape == Ql.shape[:-1] + (V.shape[-dims - 1],)
    assert torch.equal(_cf(Ol, dims), O)
    assert torch.equal(ll, l) and torch.equal(ml.view(torch.int16), m.view(torch.int16))


BWD_CASES = CASES + [
    (1, "causal", "none_front", 1, 0, False, (1,), 2, 128, 128, (640,), (640,)),     # fused dQ/dK/dV kernel, 5 key blocks
    (1, "full", "scale_end", 1, 0, False, (2,), 2, 128, 128, (200,), (1000,)),       # fused, ragged, cross lengths
    (1, "causal", "none_front", 1, 0, False, (1,), 3, 64, 64, (96,), (96,)),         # precise (split-operand) kernels
    (1, "local", "none_front", 4, 0, True, (1,), 2, 128, 128, (300,), (300,)),       # precise, head_dim 128 two-kernel
    (1, "causal", "scale_end", 1, 0, False, (1,), 2, 32, 32, (1000,), (88,)),        # precise: many queries per key
]


@pytest.mark.parametrize("case", BWD_CASES, ids=lambda c: f"{c[0]}d-{c[1]}-{c[2]}-h{c[7]}-d{c[8]}x{c[9]}-q{'x'.join(map(str, c[10]))}-k{'x'.join(map(str, c[11]))}")
def test_backward_channel_last_matches_channel_first_and_is_direct(case):
    """Same kernels, same products: dK / dV and the two-kernel dQ differ from the channel-first results at most by the
    summation order of D = rowsum(dO o O) in the statistics pass (last-bit fp32 differences before one fp16 rounding),
    the fused kernel's dQ additionally by the order of its reduce-adds (run-dependent in both layouts)."""
    dims, rule, mode, w, s, c = case[:6]
    Q, K, V, dO = _inputs(case)
    Q.requires_grad_(True), K.requires_grad_(True), V.requires_grad_(True)
    O, l, m = _call(dims, rule, Q, K, V, mode, w, s, c)
    grads = torch.autograd.grad(O, (Q, K, V), dO)
    Ql, Kl, Vl = (_cl(x.detach(), dims).requires_grad_(True) for x in (Q, K, V))
    Ol, ll, ml = _call(dims, rule, Ql, Kl, Vl, mode, w, s, c, layout="channel_last")
    dOl = _cl(dO, dims)
    torch.cuda.synchronize()
    _capi.lib.fa_launch_count(1)
    gl = torch.autograd.grad(Ol, (Ql, Kl, Vl), dOl)
    torch.cuda.synchronize()
    launches = _capi.lib.fa_launch_count(1)
    assert launches <= 4, f"{launches} launches: the channel-last backward must not go through adapter / pack passes"
    assert _capi.lib.fa_last_path() == 2
    for name, a, b in zip(("dQ", "dK", "dV"), gl, grads):
        assert a.shape == _cl(b, dims).shape
        a, b = _cf(a, dims).float(), b.float()
        err = ((a - b).abs() / b.abs().clamp(min=1.0)).max().item()
        assert err <= 1.5e-3, f"{name}: {err}"
        if a.numel() >= 4096:   # most elements round to the same fp16 value
            assert (a == b).float().mean().item() >= 0.9, name


def test_backward_channel_last_against_the_oracle():
    """One case checked against the oracle itself (not only against the channel-first kernels)."""
    from oracle import dense_attention as da
    rng = np.random.default_rng(5)
    Q, K, V, dO = da.random_inputs(rng, np.float16, (2, 3), 64, 64, (200,), (264,))
    ref = da.attention(Q, K, V, 1, "causal", "none_front", dO=dO)
    tq, tk, tv, tdo = (_cl(torch.from_numpy(x).cuda(), 1) for x in (Q, K, V, dO))
    tq.requires_grad_(True), tk.requires_grad_(True), tv.requires_grad_(True)
    O = fa.causal_1d(tq, tk, tv, "none_front", layout="channel_last")
    grads = torch.autograd.grad(O, (tq, tk, tv), tdo)
    assert np.max(np.abs(_cf(O.detach(), 1).cpu().numpy().astype(np.float64) - ref["O"])) <= 2e-3
    for name, g in zip(("dQ", "dK", "dV"), grads):
        got = _cf(g, 1).cpu().numpy().astype(np.float64)
        err = np.max(np.abs(got - ref[name]) / np.maximum(1.0, np.abs(ref[name])))
        assert err <= 2e-3, f"{name}: {err}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_other_dtypes_take_the_adapter(dtype):
    case = (1, "causal", "none_front", 1, 0, False, (2,), 3, 32, 32, (130,), (130,))
    Q, K, V, dO = _inputs(case, dtype)
    Q.requires_grad_(True), K.requires_grad_(True), V.requires_grad_(True)
    O, l, m = fa.causal_1d(Q, K, V, "none_front", returning_l_m=True)
    grads = torch.autograd.grad(O, (Q, K, V), dO)
    Ql, Kl, Vl = (_cl(x.detach(), 1).requires_grad_(True) for x in (Q, K, V))
    Ol, ll, ml = fa.causal_1d(Ql, Kl, Vl, "none_front", returning_l_m=True, layout="channel_last")
    gl = torch.autograd.grad(Ol, (Ql, Kl, Vl), _cl(dO, 1))
    assert torch.equal(_cf(Ol, 1), O) and torch.equal(ll, l) and torch.equal(ml, m)
    for a, b in zip(gl, grads):
        assert torch.equal(_cf(a, 1), b)


def test_channel_last_needs_aligned_channel_counts():
    """channels that are not multiples of 8 halves cannot be strides of a tensor map: the adapter path takes them"""
    case = (1, "full", "none_front", 1, 0, False, (2,), 3, 12, 20, (64,), (96,))
    Q, K, V, _ = _inputs(case)
    O = fa.full_1d(Q, K, V, "none_front")
    torch.cuda.synchronize()
    _capi.lib.fa_launch_count(1)
    Ol = fa.full_1d(*(_cl(x, 1) for x in (Q, K, V)), "none_front", layout="channel_last")
    torch.cuda.synchronize()
    assert _capi.lib.fa_launch_count(1) > 1
    assert torch.equal(_cf(Ol, 1), O)
