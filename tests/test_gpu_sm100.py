"""GPU parity tests for the tcgen05 / TMEM / TMA fp16 kernels (forward, dQ, dK/dV), called through
the C ABI and compared with the dense oracle. Shapes satisfy the fast path's constraints
(d, v_d in {64,128}; sequence lengths multiples of 8) and the test asserts that the tcgen05 family
was the one dispatched (fa_last_path() == 2), so a silent fallback cannot pass."""
import zlib

import numpy as np
import pytest

from oracle import dense_attention as da
from tests.helpers import max_abs_err, scaled_err

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
from tf_flash_attention_b200 import _capi  # noqa: E402
from tf_flash_attention_b200 import flash_attention as fa  # noqa: E402

TOL = 2e-3  # BASELINE.json: fp16 max-abs on O and on the gradients (scaled by max(1,|ref|) for gradients)


# fp16 gradients: one bar, BASELINE.json's 2e-3 (scaled by max(1, |ref|)). dS enters the tensor cores as a hi + lo pair
# of fp16 values on every path but the fused head_dim-128 kernel (fa_set_grad_precision, include/fa_b200.h): with a
# single fp16 rounding of dS the case "1000 queries x 88 keys" measured 4.2e-3 on dK (rows that attend one or two keys
# have |dS| of several units) and 2135 heads x 64 x 64 measured 2.005e-3.


def _run(dims, rule, mode, w, s, c, batch, d, vd, qs, ks, seed=0, grads=True):
    rng = np.random.default_rng(seed)
    Q, K, V, dO = da.random_inputs(rng, np.float16, batch, d, vd, qs, ks)
    ref = da.attention(Q, K, V, dims, rule, mode, w, s, c, dO=dO)
    tq, tk, tv = (torch.from_numpy(x).cuda().requires_grad_(True) for x in (Q, K, V))
    if rule == "full":
        O, l, m = (fa.full_1d if dims == 1 else fa.full_2d)(tq, tk, tv, mode, True)
    elif rule == "causal":
        O, l, m = (fa.causal_1d if dims == 1 else fa.causal_2d)(tq, tk, tv, mode, True)
    else:
        O, l, m = (fa.local_1d if dims == 1 else fa.local_2d)(tq, tk, tv, w, s, c, mode, True)
    torch.cuda.synchronize()
    assert _capi.lib.fa_last_path() == 2, "forward did not take the tcgen05 path"
    tag = f"{dims}d {rule} {mode} w{w} s{s} c{c} b{batch} d{d} vd{vd} q{qs} k{ks}"
    On = O.detach().cpu().numpy()
    assert max_abs_err(On, ref["O"]) <= TOL, f"O {tag}"
    ln, mn = l.cpu().numpy().astype(np.float64), m.cpu().numpy()
    empty = ~np.isfinite(ref["m"])
    if empty.any():
        assert np.all(ln[empty] == 0) and np.all(mn[empty].view(np.uint8) == 0xFA)
        assert np.all(On[np.broadcast_to(np.expand_dims(empty, -dims - 1), On.shape)] == 0)
    live = ~empty
    lse = mn.astype(np.float64)[live] + np.log(ln[live])
    assert np.max(np.abs(lse - (ref["m"][live] + np.log(ref["l"][live])))) <= 2e-2, f"lse {tag}"
    assert np.max(np.abs(mn.astype(np.float64)[live] - ref["m"][live]) / np.maximum(1, np.abs(ref["m"][live]))) <= 2e-3
    if grads:
        dQ, dK, dV = torch.autograd.grad(O, (tq, tk, tv), torch.from_numpy(dO).cuda())
        torch.cuda.synchronize()
        assert _capi.lib.fa_last_path() == 2, "backward did not take the tcgen05 path"
        nq, nk = int(np.prod(qs)), int(np.prod(ks))
        for name, g in (("dQ", dQ), ("dK", dK), ("dV", dV)):
            err = scaled_err(g.cpu().numpy(), ref[name])
            print(f"GRADERR {name} {tag}: {err:.3e}")
            assert err <= TOL, f"{name} {tag}: {err}"


CASES = [
    # dims rule mode w s c batch d vd q k
    (1, "full", "none_front", 1, 0, 0, (1,), 128, 128, (256,), (128,)),
    (1, "full", "scale_front", 1, 0, 0, (2,), 128, 128, (512,), (640,)),
    (1, "full", "scale_end", 1, 0, 0, (2,), 64, 64, (128,), (1024,)),          # C4-like cross attention
    (1, "causal", "none_front", 1, 0, 0, (2, 2), 128, 128, (1024,), (1024,)),  # C2-like
    (1, "causal", "none_front", 1, 0, 0, (3,), 64, 64, (768,), (768,)),
    (1, "causal", "scale_front", 1, 0, 0, (2,), 64, 128, (264,), (1032,)),
    (1, "causal", "scale_end", 1, 0, 0, (2,), 128, 64, (128,), (1024,)),
    (1, "causal", "scale_end", 1, 0, 0, (2,), 64, 64, (1000,), (88,)),         # more queries than keys
    (1, "full", "none_front", 1, 0, 0, (2,), 64, 128, (200,), (328,)),         # ragged tiles
    (1, "local", "scale_front", 32, 0, 0, (2,), 64, 64, (512,), (1024,)),      # C1-like window
    (1, "local", "none_front", 5, 2, 1, (1,), 128, 128, (520,), (520,)),       # strided + causal
    (1, "local", "none_front", 2, 0, 0, (2,), 64, 64, (640,), (128,)),         # rows with no keys
    (1, "local", "scale_end", 40, 1, 0, (1,), 128, 128, (384,), (768,)),
    (2, "full", "none_front", 1, 0, 0, (1,), 64, 64, (16, 24), (24, 16)),
    (2, "causal", "scale_front", 1, 0, 0, (1,), 128, 128, (16, 24), (32, 24)),
    (2, "causal", "scale_end", 1, 0, 0, (2,), 64, 64, (8, 40), (24, 40)),
    (2, "local", "none_front", 4, 0, 1, (2,), 64, 64, (24, 32), (24, 32)),     # C3-like
    (2, "local", "scale_front", 3, 1, 0, (1,), 64, 128, (20, 24), (40, 24)),
    (2, "local", "none_front", 8, 0, 1, (1,), 64, 64, (64, 64), (64, 64)),     # C3 at full grid size
    (1, "full", "scale_end", 1, 0, 0, (2,), 64, 64, (1024,), (8192,)),         # C4 at full size (cross attention)
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-{c[1]}-{c[2]}-w{c[3]}s{c[4]}c{c[5]}-d{c[7]}x{c[8]}-q{'x'.join(map(str, c[9]))}-k{'x'.join(map(str, c[10]))}")
def test_tcgen05_paths_match_oracle(case):
    _run(*case, seed=zlib.crc32(repr(case).encode()) % 1000)


def test_full_size_c2_one_head_vs_chunked_oracle():
    """BASELINE.json configs[1] at full size for one head: S = 8192, d = 128, causal."""
    rng = np.random.default_rng(1234)
    Q, K, V, dO = da.random_inputs(rng, np.float16, (1,), 128, 128, (8192,), (8192,))
    ref = da.forward_backward_chunked(Q[0], K[0], V[0], dO[0], (8192,), (8192,), "none_front", "causal")
    tq, tk, tv = (torch.from_numpy(x).cuda().requires_grad_(True) for x in (Q, K, V))
    O = fa.causal_1d(tq, tk, tv, "none_front")
    assert _capi.lib.fa_last_path() == 2
    assert max_abs_err(O.detach().cpu().numpy()[0], ref["O"]) <= TOL
    dQ, dK, dV = torch.autograd.grad(O, (tq, tk, tv), torch.from_numpy(dO).cuda())
    assert _capi.lib.fa_last_path() == 2
    for name, g in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        gn = g.cpu().numpy()[0].astype(np.float64)
        err = np.abs(gn - ref[name]) / np.maximum(1.0, np.abs(ref[name]))
        print(name, "max scaled err", err.max(), "frac within 2e-3", np.mean(err <= TOL))
        assert err.max() <= TOL, name


def test_full_size_properties_c2():
    """Size-independent properties at the full C2 batch (16 x 16 heads would take the oracle hours):
    (1) V = ones  ->  O = 1 on every row (softmax rows sum to one);
    (2) linearity in V: O(V1 + V2) = O(V1) + O(V2);
    (3) causal prefix: the first 1024 rows do not depend on later keys;
    (4) batch elements are independent: permuting heads permutes outputs bit-exactly."""
    g = torch.Generator(device="cuda").manual_seed(7)
    B, d, S = 8, 128, 8192

    def u(*shape):
        return (torch.rand(shape, generator=g, device="cuda") * 4 - 2).half()
    Q, K, V1, V2 = u(B, d, S), u(B, d, S), u(B, d, S), u(B, d, S)
    ones = torch.ones_like(V1)
    O1 = fa.causal_1d(Q, K, ones, "none_front")
    assert _capi.lib.fa_last_path() == 2
    assert float((O1.float() - 1).abs().max()) <= 2e-3
    Oa, Ob = fa.causal_1d(Q, K, V1, "none_front"), fa.causal_1d(Q, K, V2, "none_front")
    Oab = fa.causal_1d(Q, K, (V1.float() * 0.5 + V2.float() * 0.5).half(), "none_front")
    assert float((Oab.float() - 0.5 * (Oa.float() + Ob.float())).abs().max()) <= 4e-3
    Op = fa.causal_1d(Q[:, :, :1024].contiguous(), K[:, :, :1024].contiguous(), V1[:, :, :1024].contiguous(), "none_front")
    assert float((Op.float() - Oa[:, :, :1024].float()).abs().max()) <= 2e-3
    perm = torch.randperm(B, device="cuda")
    Oq = fa.causal_1d(Q[perm].contiguous(), K[perm].contiguous(), V1[perm].contiguous(), "none_front")
    assert torch.equal(Oq, Oa[perm])


def test_full_size_backward_properties_c2():
    """Size-independent properties of the backward pass at the full C2 sequence length (8 heads):
    (1) linearity in dO: grad(dO1 + dO2) = grad(dO1) + grad(dO2) (the same P is recomputed in every call);
    (2) batch elements are independent: permuting heads permutes dK / dV bit-exactly (dQ up to the fp32 reduction order);
    (3) sum_k dK relation: with dO = 0 every gradient is exactly zero."""
    g = torch.Generator(device="cuda").manual_seed(11)
    B, d, S = 8, 128, 8192

    def u(*shape, scale=2.0):
        return ((torch.rand(shape, generator=g, device="cuda") * 2 - 1) * scale).half()
    Q, K, V = (u(B, d, S).requires_grad_(True) for _ in range(3))
    dO1, dO2 = u(B, d, S, scale=1.0), u(B, d, S, scale=1.0)
    O = fa.causal_1d(Q, K, V, "none_front")

    def grads(dO):
        out = torch.autograd.grad(O, (Q, K, V), dO, retain_graph=True)
        assert _capi.lib.fa_last_path() == 2
        return [x.float() for x in out]
    g1, g2, g12 = grads(dO1), grads(dO2), grads((dO1.float() + dO2.float()).half())
    for a, b, c, name in zip(g1, g2, g12, ("dQ", "dK", "dV")):
        err = (c - (a + b)).abs() / (a + b).abs().clamp(min=1.0)
        assert float(err.max()) <= 8e-3, f"{name}: {float(err.max())}"       # three fp16 roundings of sums of 8192 terms
        assert float((err <= 2e-3).float().mean()) >= 0.995, name
    perm = torch.randperm(B, device="cuda")
    Qp, Kp, Vp = (x.detach()[perm].contiguous().requires_grad_(True) for x in (Q, K, V))
    Op = fa.causal_1d(Qp, Kp, Vp, "none_front")
    gp = torch.autograd.grad(Op, (Qp, Kp, Vp), dO1[perm].contiguous())
    assert torch.equal(gp[1].float(), g1[1][perm]) and torch.equal(gp[2].float(), g1[2][perm])
    assert float((gp[0].float() - g1[0][perm]).abs().max()) <= 2e-3 * max(1.0, float(g1[0].abs().max()))
    gz = grads(torch.zeros_like(dO1))
    assert all(float(x.abs().max()) == 0.0 for x in gz)


D128_CASES = [c for c in CASES if c[7] == 128]


@pytest.mark.parametrize("case", D128_CASES, ids=lambda c: f"{c[0]}d-{c[1]}-{c[2]}-w{c[3]}s{c[4]}c{c[5]}-d{c[7]}x{c[8]}-q{'x'.join(map(str, c[9]))}-k{'x'.join(map(str, c[10]))}")
def test_split_backward_variant_matches_oracle(case):
    """head_dim 128 normally runs the fused dQ/dK/dV kernel; fa_set_path_override(4) selects the two-kernel
    variant (dQ kernel, then dK/dV kernel) that head_dim 64 always uses. Both must meet the same bar."""
    _capi.lib.fa_set_path_override(4)
    try:
        _run(*case, seed=zlib.crc32(repr(case).encode()) % 1000)
    finally:
        _capi.lib.fa_set_path_override(0)


def test_fused_backward_agrees_with_split_variant():
    """Same inputs through both backward variants: dK and dV are computed by identical arithmetic (bit-equal);
    dQ is accumulated in fp32 across key tiles in a different order (fused: atomic adds) -> within 1 fp16 ulp-ish."""
    rng = np.random.default_rng(5)
    Q, K, V, dO = da.random_inputs(rng, np.float16, (3,), 128, 128, (1024,), (1536,))
    tq, tk, tv = (torch.from_numpy(x).cuda().requires_grad_(True) for x in (Q, K, V))
    tdo = torch.from_numpy(dO).cuda()
    out = {}
    for variant in (0, 4):
        _capi.lib.fa_set_path_override(variant)
        try:
            O = fa.causal_1d(tq, tk, tv, "scale_end")
            out[variant] = torch.autograd.grad(O, (tq, tk, tv), tdo)
            assert _capi.lib.fa_last_path() == 2
        finally:
            _capi.lib.fa_set_path_override(0)
    assert torch.equal(out[0][1], out[4][1]) and torch.equal(out[0][2], out[4][2])
    dq_a, dq_b = out[0][0].float(), out[4][0].float()
    assert float(((dq_a - dq_b).abs() / dq_b.abs().clamp(min=1)).max()) <= 1e-3


D64_CASES = [c for c in CASES if c[7] == 64 and c[8] == 64][:6]


@pytest.mark.parametrize("override", [4, 5])
@pytest.mark.parametrize("case", D64_CASES, ids=lambda c: f"{c[0]}d-{c[1]}-{c[2]}-w{c[3]}s{c[4]}c{c[5]}-q{'x'.join(map(str, c[9]))}-k{'x'.join(map(str, c[10]))}")
def test_head_dim64_one_cta_per_sm_variants_match_oracle(case, override):
    """head_dim 64 normally runs the 256-TMEM-column configurations (two CTAs per SM: 64-key forward tiles, one-slot dQ
    and dK/dV kernels). Override 5 selects the 128-key forward, override 4 the two-slot backward kernels; both stay
    in the library for A/B runs and must meet the same bar."""
    _capi.lib.fa_set_path_override(override)
    try:
        _run(*case, seed=zlib.crc32(repr(case).encode()) % 1000)
    finally:
        _capi.lib.fa_set_path_override(0)


TINY_CASES = [
    (1, "causal", "none_front", 1, 0, 0, (2,), 64, 64, (8,), (8,)),      # a single, mostly empty tile per CTA
    (1, "full", "scale_end", 1, 0, 0, (3,), 128, 128, (8,), (24,)),
    (1, "local", "none_front", 2, 0, 1, (1,), 128, 64, (16,), (16,)),
    (2, "causal", "none_front", 1, 0, 0, (2,), 64, 64, (2, 4), (2, 4)),
]


@pytest.mark.parametrize("case", TINY_CASES, ids=lambda c: f"{c[0]}d-{c[1]}-d{c[7]}x{c[8]}-q{'x'.join(map(str, c[9]))}-k{'x'.join(map(str, c[10]))}")
def test_tiny_sequences_on_the_tcgen05_paths(case):
    """Sequences far shorter than a tile (TMA boxes reach past the tensor: zero fill on loads, clipping on stores)."""
    _run(*case, seed=1)


def test_tiny_sequences_fp32_split_paths():
    rng = np.random.default_rng(2)
    Q, K, V, dO = da.random_inputs(rng, np.float32, (2,), 64, 64, (8,), (16,))
    ref = da.attention(Q, K, V, 1, "causal", "scale_end", dO=dO)
    tq, tk, tv = (torch.from_numpy(x).cuda().requires_grad_(True) for x in (Q, K, V))
    O = fa.causal_1d(tq, tk, tv, "scale_end")
    assert _capi.lib.fa_last_path() == 3
    dQ, dK, dV = torch.autograd.grad(O, (tq, tk, tv), torch.from_numpy(dO).cuda())
    assert _capi.lib.fa_last_path() == 3
    assert max_abs_err(O.detach().cpu().numpy(), ref["O"]) <= 1e-5
    for name, g in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        assert scaled_err(g.cpu().numpy(), ref[name]) <= 1e-5, name


F32_CASES = [
    (1, "full", "none_front", 1, 0, 0, (2,), 64, 64, (256,), (320,)),
    (1, "causal", "none_front", 1, 0, 0, (2,), 64, 64, (512,), (512,)),
    (1, "full", "scale_end", 1, 0, 0, (2,), 64, 64, (1024,), (8192,)),          # C4 fp32 variant at full size
    (1, "local", "scale_front", 32, 0, 0, (8,), 32, 16, (1024,), (2048,)),      # C1 (README example)
    (1, "causal", "scale_front", 1, 0, 0, (2,), 32, 32, (100,), (204,)),        # ragged tiles
    (1, "local", "none_front", 2, 0, 0, (2,), 32, 32, (640,), (128,)),          # rows with no keys
    (2, "local", "none_front", 4, 0, 1, (2,), 64, 64, (24, 32), (24, 32)),
    (2, "causal", "scale_end", 1, 0, 0, (1,), 32, 32, (8, 40), (24, 40)),
    (1, "causal", "none_front", 1, 0, 0, (2, 2), 64, 64, (1024,), (1024,)),    # fp32 backward on the tensor cores
    (1, "local", "none_front", 5, 1, 1, (2,), 64, 64, (520,), (520,)),          # ragged, strided window
    (1, "causal", "scale_end", 1, 0, 0, (2,), 64, 64, (1000,), (88,)),          # more queries than keys
    (2, "causal", "scale_front", 1, 0, 0, (1,), 64, 64, (16, 24), (32, 24)),
]


@pytest.mark.parametrize("case", F32_CASES, ids=lambda c: f"{c[0]}d-{c[1]}-{c[2]}-d{c[7]}x{c[8]}-q{'x'.join(map(str, c[9]))}-k{'x'.join(map(str, c[10]))}")
def test_fp32_3xtf32_forward_matches_oracle(case):
    """fp32 forward on the tensor cores (tcgen05 kind::tf32, 3xTF32 split): BASELINE.json bar 1e-5 max-abs on O;
    the backward (tensor-core kernels for head_dim 64, FFMA kernels otherwise) then consumes its l, m."""
    dims, rule, mode, w, s, c, batch, d, vd, qs, ks = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) % 1000)
    Q, K, V, dO = da.random_inputs(rng, np.float32, batch, d, vd, qs, ks)
    ref = da.attention(Q, K, V, dims, rule, mode, w, s, c, dO=dO)
    tq, tk, tv = (torch.from_numpy(x).cuda().requires_grad_(True) for x in (Q, K, V))
    if rule == "full":
        O, l, m = (fa.full_1d if dims == 1 else fa.full_2d)(tq, tk, tv, mode, True)
    elif rule == "causal":
        O, l, m = (fa.causal_1d if dims == 1 else fa.causal_2d)(tq, tk, tv, mode, True)
    else:
        O, l, m = (fa.local_1d if dims == 1 else fa.local_2d)(tq, tk, tv, w, s, c, mode, True)
    torch.cuda.synchronize()
    assert _capi.lib.fa_last_path() == 3, "fp32 forward did not take the 3xTF32 tcgen05 path"
    On = O.detach().cpu().numpy()
    assert max_abs_err(On, ref["O"]) <= 1e-5
    ln, mn = l.cpu().numpy().astype(np.float64), m.cpu().numpy()
    empty = ~np.isfinite(ref["m"])
    if empty.any():
        assert np.all(ln[empty] == 0) and np.all(mn[empty].view(np.uint8) == 0xFA)
        assert np.all(On[np.broadcast_to(np.expand_dims(empty, -dims - 1), On.shape)] == 0)
    live = ~empty
    lse = mn.astype(np.float64)[live] + np.log(ln[live])
    assert np.max(np.abs(lse - (ref["m"][live] + np.log(ref["l"][live])))) <= 1e-5
    dQ, dK, dV = torch.autograd.grad(O, (tq, tk, tv), torch.from_numpy(dO).cuda())
    nq, nk = int(np.prod(qs)), int(np.prod(ks))
    if (d, vd) in ((64, 64), (32, 32), (32, 16)) and nq % 8 == 0 and nk % 8 == 0:
        # split-precision tcgen05 backward (three bf16 pieces per operand, fa_bwd_f32_sm100.cu)
        assert _capi.lib.fa_last_path() == 3, "fp32 backward did not take the tensor-core path"
    for name, g in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        err = scaled_err(g.cpu().numpy(), ref[name])
        print(name, err)
        assert err <= 1e-5, name


def test_full_batch_c2_two_random_heads_vs_chunked_oracle():
    """BASELINE.json configs[1] exactly as written — batch x heads 16 x 16, head_dim 128, seq 8192, causal, fwd + bwd in
    ONE call over all 256 heads — spot-checked against the chunked oracle on two heads drawn at random (the oracle needs
    ~10 s per head; the other heads are covered by the size-independent properties above)."""
    g = torch.Generator(device="cuda").manual_seed(2026)
    B0, B1, d, S = 16, 16, 128, 8192

    def u():
        return (torch.rand((B0, B1, d, S), generator=g, device="cuda") * 4 - 2).half()
    Q, K, V, dO = u(), u(), u(), u()
    tq, tk, tv = (x.requires_grad_(True) for x in (Q, K, V))
    O = fa.causal_1d(tq, tk, tv, "none_front")
    assert _capi.lib.fa_last_path() == 2
    dQ, dK, dV = torch.autograd.grad(O, (tq, tk, tv), dO)
    assert _capi.lib.fa_last_path() == 2
    torch.cuda.synchronize()
    rng = np.random.default_rng(99)
    for h in rng.choice(B0 * B1, size=2, replace=False):
        i, j = divmod(int(h), B1)
        ref = da.forward_backward_chunked(*(x[i, j].detach().cpu().numpy() for x in (Q, K, V, dO)), (S,), (S,),
                                          "none_front", "causal")
        assert max_abs_err(O[i, j].detach().cpu().numpy(), ref["O"]) <= TOL, f"O head {h}"
        for name, gr in (("dQ", dQ), ("dK", dK), ("dV", dV)):
            err = scaled_err(gr[i, j].cpu().numpy(), ref[name])
            print(f"GRADERR C2 full batch head {h} {name}: {err:.3e}")
            assert err <= TOL, f"{name} head {h}: {err}"


D128_SQUARE = [c for c in CASES if c[7] == 128 and c[8] == 128]


@pytest.mark.parametrize("case", D128_SQUARE, ids=lambda c: f"{c[0]}d-{c[1]}-{c[2]}-w{c[3]}s{c[4]}c{c[5]}-q{'x'.join(map(str, c[9]))}-k{'x'.join(map(str, c[10]))}")
def test_grad_precision_1_head_dim128_split_operands(case):
    """fa_set_grad_precision(1): head_dim 128 leaves the fused kernel for the two-kernel backward with dS as hi + lo
    fp16 pairs; same bar."""
    assert _capi.lib.fa_set_grad_precision(1) == 0
    try:
        _run(*case, seed=zlib.crc32(repr(case).encode()) % 1000)
    finally:
        _capi.lib.fa_set_grad_precision(0)


def test_grad_precision_modes_on_the_large_ds_case():
    """1000 queries x 88 keys, causal scale_end (rows that attend one or two keys: |dS| of several units). Mode 2 (dS as one
    fp16 value) is the round-1 behaviour and misses the bar on dK; the default (hi + lo pairs) meets it with margin."""
    case = (1, "causal", "scale_end", 1, 0, 0, (2,), 64, 64, (1000,), (88,))
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) % 1000)
    Q, K, V, dO = da.random_inputs(rng, np.float16, (2,), 64, 64, (1000,), (88,))
    ref = da.attention(Q, K, V, 1, "causal", "scale_end", dO=dO)
    errs = {}
    for mode in (2, 0):
        assert _capi.lib.fa_set_grad_precision(mode) == 0
        try:
            tq, tk, tv = (torch.from_numpy(x).cuda().requires_grad_(True) for x in (Q, K, V))
            O = fa.causal_1d(tq, tk, tv, "scale_end")
            dQ, dK, dV = torch.autograd.grad(O, (tq, tk, tv), torch.from_numpy(dO).cuda())
            errs[mode] = {n: scaled_err(g.cpu().numpy(), ref[n]) for n, g in (("dQ", dQ), ("dK", dK), ("dV", dV))}
        finally:
            _capi.lib.fa_set_grad_precision(0)
    print("GRADERR modes", errs)
    assert max(errs[0].values()) <= TOL
    assert errs[0]["dK"] < 0.6 * errs[2]["dK"]
    assert _capi.lib.fa_set_grad_precision(3) != 0
