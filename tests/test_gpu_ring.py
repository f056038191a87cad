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
