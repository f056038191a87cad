"""GPU test of the K/V-ring product backend (C ABI: fa_forward with global index bases,
fa_partial_merge, fa_partial_finalize). On one GPU the ring has a single rank that owns both chunks;
the multi-rank exchange is covered on CPU by tests/test_ring_gloo.py and on 2+ GPUs by
tools/ring_check.py under torchrun."""
import numpy as np
import pytest

from oracle import dense_attention as da
from tests.helpers import max_abs_err

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
from tf_flash_attention_b200 import _capi, ring  # noqa: E402


@pytest.fixture(params=["native", "python"])
def driver(request, monkeypatch):
    """Both drivers of the same schedule: fa_ring_causal_forward / _backward (csrc/fa_ring.cu, the default) and the
    Python loops of ring.py."""
    monkeypatch.setenv("FA_RING_DRIVER", request.param)
    return request.param


@pytest.mark.parametrize("dtype,d,seq,tol", [(np.float16, 128, 2048, 2e-3), (np.float16, 64, 512, 2e-3),
                                              (np.float32, 32, 384, 1e-5), (np.float64, 16, 256, 1e-12)])
def test_single_rank_ring_equals_plain_causal(dtype, d, seq, tol, driver):
    rng = np.random.default_rng(3)
    Q, K, V, _ = da.random_inputs(rng, dtype, (2, 2), d, d, (seq,), (seq,))
    ref = da.attention(Q, K, V, 1, "causal", "none_front")
    O, l, m = ring.ring_causal_1d(*(torch.from_numpy(x).cuda() for x in (Q, K, V)), returning_l_m=True)
    assert max_abs_err(O.cpu().numpy(), ref["O"]) <= tol
    lse = m.cpu().numpy().astype(np.float64) + np.log(l.cpu().numpy().astype(np.float64))
    assert np.max(np.abs(lse - (ref["m"] + np.log(ref["l"])))) <= max(2e-2 if dtype == np.float16 else tol * 10, tol)
    if dtype == np.float16:
        assert _capi.lib.fa_last_path() == 2  # partial blocks ran on the tcgen05 kernel


@pytest.mark.parametrize("dtype,d,seq,tol", [(np.float16, 128, 2048, 2e-3), (np.float16, 64, 512, 2e-3),
                                              (np.float16, 64, 256, 2e-3),   # 128-key chunks: the precise kernels, on key shards
                                              (np.float32, 32, 384, 1e-5), (np.float64, 16, 256, 1e-12)])
def test_single_rank_ring_backward_equals_plain_causal(dtype, d, seq, tol, driver):
    """ring_backward: blocks computed by fa_backward with global index bases and the final (O, l, m), summed with
    fa_grad_accumulate / fa_grad_finalize. fp16 bar: the plain 2e-3 * max(1, |ref|)."""
    rng = np.random.default_rng(4)
    Q, K, V, dO = da.random_inputs(rng, dtype, (2, 2), d, d, (seq,), (seq,))
    ref = da.attention(Q, K, V, 1, "causal", "none_front", dO=dO)
    tq, tk, tv, tdo = (torch.from_numpy(x).cuda() for x in (Q, K, V, dO))
    O, l, m = ring.ring_causal_1d(tq, tk, tv, returning_l_m=True)
    dQ, dK, dV = ring.ring_causal_1d_backward(tq, tk, tv, O, l, m, tdo)
    if dtype == np.float16:
        assert _capi.lib.fa_last_path() == 2
    for name, g in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        err = np.abs(g.cpu().numpy().astype(np.float64) - ref[name]) / np.maximum(1.0, np.abs(ref[name]))
        assert err.max() <= tol, f"{name}: {err.max()}"


@pytest.mark.parametrize("vd", [128, 64])
def test_backward_accumulate_adds_dq_inside_the_launch(vd):
    """fa_backward_accumulate (include/fa_b200.h): dQ of every batch element is added into the fp32 accumulator element
    pb % dq_fold by the fused kernel's own reduce-add (no scratch / convert / fa_grad_accumulate); dK, dV as usual."""
    import ctypes as C
    rng = np.random.default_rng(9)
    B, d, nq, nk = 4, 128, 384, 512
    Q, K, V, dO = da.random_inputs(rng, np.float16, (B,), d, vd, (nq,), (nk,))
    ref = da.attention(Q, K, V, 1, "full", "none_front", dO=dO)
    tq, tk, tv, tdo = (torch.from_numpy(x).cuda() for x in (Q, K, V, dO))
    from tf_flash_attention_b200 import flash_attention as fa
    O, l, m = fa.full_1d(tq, tk, tv, "none_front", returning_l_m=True)
    p = _capi.make_problem(_capi.FA_F16, 1, "full", "none_front", Q.shape, K.shape, V.shape)
    assert _capi.lib.fa_backward_accumulate_supported(C.byref(p), 0) == 1
    assert _capi.lib.fa_backward_accumulate_supported(C.byref(p), 2) == 1
    assert _capi.lib.fa_backward_accumulate_supported(C.byref(p), 5) == 0          # more accumulator elements than problems
    need = _capi.lib.fa_workspace_bytes(C.byref(p), 1)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for fold in (B, 2):
        acc = torch.full((fold, d, nq), 0.5, dtype=torch.float32, device="cuda")    # adds INTO the accumulator
        dk, dv = torch.empty_like(tk), torch.empty_like(tv)
        _capi.lib.fa_launch_count(1)
        _capi.check(_capi.lib.fa_backward_accumulate(C.byref(p), tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), O.data_ptr(),
                                                     l.data_ptr(), m.data_ptr(), tdo.data_ptr(), acc.data_ptr(),
                                                     dk.data_ptr(), dv.data_ptr(), fold, ws.data_ptr(), need, st),
                    "fa_backward_accumulate")
        torch.cuda.synchronize()
        assert _capi.lib.fa_launch_count(1) == 2                                  # statistics pass + the fused kernel
        want = ref["dQ"].reshape(B // fold, fold, d, nq).sum(axis=0) + 0.5
        err = np.abs(acc.cpu().numpy().astype(np.float64) - want) / np.maximum(1.0, np.abs(want))
        assert err.max() <= 2e-3, err.max()
        for name, g in (("dK", dk), ("dV", dv)):
            e2 = np.abs(g.cpu().numpy().astype(np.float64) - ref[name]) / np.maximum(1.0, np.abs(ref[name]))
            assert e2.max() <= 2e-3, (name, e2.max())
    # shapes the fused kernel does not take answer FA_EINVAL_SHAPE (the caller falls back to fa_backward + fa_grad_accumulate)
    p64 = _capi.make_problem(_capi.FA_F16, 1, "full", "none_front", (B, 64, nq), (B, 64, nk), (B, 64, nk))
    assert _capi.lib.fa_backward_accumulate_supported(C.byref(p64), 0) == 0


def test_two_rank_ring_over_the_peer_copy_data_plane():
    """N > 1 on the GPU: two ranks under torchrun run ring_causal_1d and its backward over the fa_ring_* data plane
    (CUDA IPC slots, peer copies, stream-ordered flags) and compare their rows with the dense oracle
    (tools/ring_check.py). Needs two GPUs in the box; skipped on the single-GPU test boxes."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for drv in ("native", "python"):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29533",
                            os.path.join(root, "tools", "ring_check.py")],
                           capture_output=True, text=True, timeout=600, cwd=root, env=dict(os.environ, FA_RING_DRIVER=drv))
        assert r.returncode == 0 and "RING_CHECK PASS" in r.stdout, drv + r.stdout[-2000:] + r.stderr[-2000:]
