"""The package-side verification groups (`tf_flash_attention_b200/tests`, the port of the reference's
`flash_attention/tests` command line): their masks and dense attention are checked against the oracle on the CPU, their
command line is exercised, and on a GPU a short `TestGroup.verify` runs for both ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import dense_attention as da
from oracle import pattern

torch = pytest.importorskip("torch")
from tf_flash_attention_b200.tests import test_1d, test_2d, test_base  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [((6,), (3,)), ((7,), (20,)), ((33,), (33,)), ((4, 4), (2, 2)), ((3, 10), (6, 5)), ((5, 7), (5, 7))]


def test_case_table_matches_reference_names():
    names = test_base.case_names()
    assert len(names) == 16 and len(set(names)) == 16
    assert "FullAttentionSyncModeNoneFront" in names and "LocalStrideAndCausalAttentionSyncModeScaleEnd" in names
    for n in names:
        kind, sync = test_base.parse_case(n)
        assert kind in test_base.KINDS and sync in test_base.SYNC_MODES
    for group in (test_1d.TestGroup, test_2d.TestGroup):
        assert set(group.SHAPE_TABLE) == {torch.float16, torch.float32, torch.float64}


@pytest.mark.parametrize("q_seq,k_seq", SHAPES)
@pytest.mark.parametrize("sync", test_base.SYNC_MODES)
@pytest.mark.parametrize("kind", list(test_base.KINDS))
def test_group_masks_equal_the_oracle_pattern(kind, sync, q_seq, k_seq):
    q_loc, k_loc = test_base.locations(q_seq, k_seq, sync, torch.device("cpu"))
    mask, window, log2_stride = test_base.attended(kind, q_loc, k_loc, q_seq, k_seq)
    family, strided, causal = test_base.KINDS[kind]
    if family == "local":
        assert window == max(q_seq + k_seq) and log2_stride == (int(np.log2(window)) if strided else 0)
    for formulation in (pattern.tests_mask, pattern.kernel_mask):
        ref = formulation(q_seq, k_seq, sync, family, window or 1, log2_stride or 0, causal)
        assert np.array_equal(mask.numpy(), ref), formulation.__name__


def test_group_dense_attention_equals_the_oracle():
    rng = np.random.default_rng(5)
    Q, K, V, dO = da.random_inputs(rng, np.float64, (2, 3), 8, 5, (3, 10), (6, 5))
    ref = da.attention(Q, K, V, 2, "local", "scale_end", 4, 1, True, dO=dO)
    q_loc, k_loc = test_base.locations((3, 10), (6, 5), "scale_end", torch.device("cpu"))
    diff = (q_loc[0][:, None, :] - k_loc[0][None, :, :]).abs()
    mask = ((diff % 2 == 0) & (diff // 2 < 4)).all(dim=-1) & (q_loc[1][:, None] >= k_loc[1][None, :])
    tq, tk, tv = (torch.from_numpy(x).requires_grad_(True) for x in (Q, K, V))
    O = test_base.dense_attention(tq, tk, tv, mask, 2)
    grads = torch.autograd.grad(O, (tq, tk, tv), torch.from_numpy(dO))
    assert np.abs(O.detach().numpy() - ref["O"]).max() < 1e-12
    for g, n in zip(grads, ("dQ", "dK", "dV")):
        assert np.abs(g.numpy() - ref[n]).max() < 1e-11, n


def _run(module, *argv, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, "-m", f"tf_flash_attention_b200.tests.{module}", *argv], cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=900)


def test_command_line_list_and_usage():
    r = _run("test_1d", "TestGroup.list")
    assert r.returncode == 0 and r.stdout.splitlines()[0] == "Available testcases:" and len(r.stdout.splitlines()) == 17
    r = _run("test_2d", "TestGroup.list", env={"TESTCASE": "CausalAttentionSyncModeScaleEnd"})
    assert r.stdout.splitlines() == ["Available testcases:", "CausalAttentionSyncModeScaleEnd"]
    assert _run("test_1d", "TestGroup.nothing").returncode == 2


@pytest.mark.gpu
@pytest.mark.parametrize("module", ["test_1d", "test_2d"])
def test_groups_verify_on_the_gpu(module):
    r = _run(module, "TestGroup.verify", env={"RUNS": "2", "SEED": "1234"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("Verifying ") == 16 and r.stdout.count("shapes ok") == 48
