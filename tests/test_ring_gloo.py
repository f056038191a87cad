"""N > 1 host logic on CPU: the K/V ring driver (tf_flash_attention_b200/ring.py: zig-zag chunk
ownership, per-step block plan, global index bases, double-buffered send/recv) run with world_size 2
and 4 over the gloo backend. The compute stand-in is the NumPy oracle (test code only); the product
backend (DeviceBackend -> C ABI) is exercised on GPUs by tests/test_gpu_ring.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dense_attention as da
from oracle import pattern
from tf_flash_attention_b200 import ring


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleBackend:
    """NumPy stand-in with the same interface as ring.DeviceBackend (float64 CPU tensors)."""

    def __init__(self, batch, d, v_d, chunk):
        self.batch, self.d, self.v_d, self.chunk = batch, d, v_d, chunk

    def new_acc(self, like_q):
        return [torch.zeros(self.batch, self.v_d, self.chunk, dtype=torch.float64),
                torch.zeros(self.batch, self.chunk, dtype=torch.float64),
                torch.full((self.batch, self.chunk), -np.inf, dtype=torch.float64)]

    def new_out(self, like_q):
        return self.new_acc(like_q)

    def empty_like_kv(self, kv):
        return [torch.empty_like(x) for x in kv]

    def attend_partial(self, q, k, v, q_base, k_base, out):
        qi = q_base + np.arange(self.chunk)
        kj = k_base + np.arange(self.chunk)
        mask = qi[:, None] >= kj[None, :]
        O, l, m = da.forward(q.numpy(), k.numpy(), v.numpy(), mask)
        out[0].copy_(torch.from_numpy(O)); out[1].copy_(torch.from_numpy(l)); out[2].copy_(torch.from_numpy(m))

    def merge(self, part, acc, first):
        o, l, m = (x.numpy() for x in part)
        oa, la, ma = (x.numpy() for x in acc)
        if first:
            oa[...] = 0; la[...] = 0; ma[...] = -np.inf
        m_new = np.maximum(ma, m)
        safe = np.where(np.isfinite(m_new), m_new, 0.0)
        wa = np.where(np.isfinite(ma), np.exp(ma - safe), 0.0)
        wp = np.where(np.isfinite(m), np.exp(m - safe) * l, 0.0)
        oa[...] = oa * wa[:, None, :] + o * wp[:, None, :]
        la[...] = la * wa + wp
        ma[...] = m_new

    def finalize(self, acc, out):
        oa, la, ma = acc
        out[0].copy_(oa / torch.where(la > 0, la, torch.ones_like(la))[:, None, :])
        out[1].copy_(la); out[2].copy_(ma)

    # ---- backward stand-in: one block with the FINAL statistics (kernel/internal_test.cu:413-511 algebra)
    def new_grad_acc(self, like):
        return torch.zeros_like(like, dtype=torch.float64)

    def new_grad_part(self, q, k, v):
        return (torch.empty_like(q), torch.empty_like(k), torch.empty_like(v))

    def grad_partial(self, q, k, v, o, l, m, d_o, q_base, k_base, part):
        qi = q_base + np.arange(self.chunk)
        kj = k_base + np.arange(self.chunk)
        mask = qi[:, None] >= kj[None, :]
        qn, kn, vn, on, ln, mn, don = (x.numpy() for x in (q, k, v, o, l, m, d_o))
        scale = 1.0 / np.sqrt(self.d)
        logit = np.einsum("bcq,bck->bqk", qn, kn) * scale
        P = np.where(mask[None], np.exp(logit - mn[..., None]) / ln[..., None], 0.0)
        D = np.einsum("bcq,bcq->bq", don, on)
        dP = np.einsum("bcq,bck->bqk", don, vn)
        dS = P * (dP - D[..., None]) * scale
        part[0].copy_(torch.from_numpy(np.einsum("bqk,bck->bcq", dS, kn)))
        part[1].copy_(torch.from_numpy(np.einsum("bqk,bcq->bck", dS, qn)))
        part[2].copy_(torch.from_numpy(np.einsum("bqk,bcq->bck", P, don)))

    def grad_add(self, part, acc):
        acc += part

    def grad_finalize(self, acc, like):
        return acc.clone()


class BatchedOracleBackend(OracleBackend):
    """NumPy stand-in for the batched causal schedule (ring.ring_forward_causal / ring_backward_causal): any leading
    batch dims, rules in local coordinates."""

    @staticmethod
    def _mask(rule, nq, nk):
        if rule == "full":
            return np.ones((nq, nk), dtype=bool)
        return np.arange(nq)[:, None] >= np.arange(nk)[None, :]

    def stack2(self, a, b):
        return torch.stack([a, b]).contiguous()

    def new_out_like(self, q, v):
        lead = tuple(q.shape[:-2])
        return [torch.zeros(lead + (v.shape[-2], q.shape[-1]), dtype=torch.float64),
                torch.zeros(lead + (q.shape[-1],), dtype=torch.float64),
                torch.full(lead + (q.shape[-1],), -np.inf, dtype=torch.float64)]

    new_acc_like = new_out_like

    def attend(self, rule, q, k, v, out):
        qn, kn, vn = (x.numpy().reshape((-1,) + tuple(x.shape[-2:])) for x in (q, k, v))
        O, l, m = da.forward(qn, kn, vn, self._mask(rule, q.shape[-1], k.shape[-1]))
        for dst, src in zip(out, (O, l, m)):
            dst.copy_(torch.from_numpy(np.ascontiguousarray(src)).reshape(dst.shape))

    def merge_into(self, part, acc, first, q, k, v):
        flat = lambda xs: [x.reshape((-1,) + tuple(x.shape[x.dim() - (2 if i == 0 else 1):])) for i, x in enumerate(xs)]
        self.merge(flat(part), flat(acc), first)

    def finalize_into(self, acc, out, q, k, v):
        oa, la, ma = acc
        out[0].copy_(oa / torch.where(la > 0, la, torch.ones_like(la)).unsqueeze(-2))
        out[1].copy_(la); out[2].copy_(ma)

    def grad(self, rule, q, k, v, o, l, m, d_o, part):
        mask = self._mask(rule, q.shape[-1], k.shape[-1])
        qn, kn, vn, on, ln, mn, don = (x.numpy() for x in (q, k, v, o, l, m, d_o))
        scale = 1.0 / np.sqrt(q.shape[-2])
        logit = np.einsum("...cq,...ck->...qk", qn, kn) * scale
        P = np.where(mask, np.exp(logit - mn[..., None]) / ln[..., None], 0.0)
        D = np.einsum("...cq,...cq->...q", don, on)
        dP = np.einsum("...cq,...ck->...qk", don, vn)
        dS = P * (dP - D[..., None]) * scale
        part[0].copy_(torch.from_numpy(np.einsum("...qk,...ck->...cq", dS, kn)))
        part[1].copy_(torch.from_numpy(np.einsum("...qk,...cq->...ck", dS, qn)))
        part[2].copy_(torch.from_numpy(np.einsum("...qk,...cq->...ck", P, don)))


def _worker_batched(rank, world, port, seq, d, v_d, batch, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(13)  # same full tensors on every rank
    Q, K, V, dO = da.random_inputs(rng, np.float64, (batch,), d, v_d, (seq,), (seq,))
    mask = pattern.tests_mask((seq,), (seq,), "none_front", "causal")
    O, l, m = da.forward(Q, K, V, mask)
    ref = dict(zip(("dQ", "dK", "dV"), da.backward(Q, K, V, mask, dO)))
    layout = ring.ZigZag(seq, world)
    idx = layout.gather_index(rank)
    c = layout.chunk

    def cm(X):   # this rank's shard, chunk-major
        return torch.from_numpy(np.ascontiguousarray(np.stack([X[..., idx[:c]], X[..., idx[c:]]])))
    backend = BatchedOracleBackend(batch, d, v_d, c)
    O2, l2, m2 = ring.ring_forward_causal(backend, layout, rank, cm(Q), cm(K), cm(V), dist, None)
    got_o = np.concatenate([O2[0].numpy(), O2[1].numpy()], axis=-1)
    errs = [np.abs(got_o - O[:, :, idx]).max()]
    lse = np.concatenate([(m2[h] + torch.log(l2[h])).numpy() for h in range(2)], axis=-1)
    errs.append(np.abs(lse - (m + np.log(l))[:, idx]).max())
    grads = ring.ring_backward_causal(backend, layout, rank, cm(Q), cm(K), cm(V), cm(O), cm(l), cm(m), cm(dO), dist, None)
    for name, g in zip(("dQ", "dK", "dV"), grads):
        got = np.concatenate([g[0].numpy(), g[1].numpy()], axis=-1)
        errs.append(np.abs(got - ref[name][:, :, idx]).max())
    np.save(os.path.join(result_dir, f"cerr_{rank}.npy"), np.array(errs))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_batched_causal_ring_matches_dense_oracle(world, tmp_path):
    """ring_forward_causal / ring_backward_causal: one launch per ring step over chunk-major shards, plain causal / full
    rules in local coordinates (no index bases); forward O and log-sum-exp, then dQ, dK, dV against the dense oracle."""
    seq = 16 * world
    mp.spawn(_worker_batched, args=(world, _free_port(), seq, 8, 6, 2, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert float(np.load(tmp_path / f"cerr_{r}.npy").max()) < 1e-11


def _worker(rank, world, port, seq, d, v_d, batch, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)  # same full tensors on every rank
    Q, K, V, _ = da.random_inputs(rng, np.float64, (batch,), d, v_d, (seq,), (seq,))
    layout = ring.ZigZag(seq, world)
    idx = layout.gather_index(rank)
    c = layout.chunk
    qs = [torch.from_numpy(np.ascontiguousarray(Q[:, :, idx[:c]])), torch.from_numpy(np.ascontiguousarray(Q[:, :, idx[c:]]))]
    kv = [torch.from_numpy(np.ascontiguousarray(X[:, :, sl])) for X in (K, V) for sl in (idx[:c], idx[c:])]
    outs = ring.ring_forward(OracleBackend(batch, d, v_d, c), layout, rank, qs, kv, dist, None, "causal")
    O = np.concatenate([outs[0][0].numpy(), outs[1][0].numpy()], axis=-1)
    ref = da.forward(Q, K, V, pattern.tests_mask((seq,), (seq,), "none_front", "causal"))[0][:, :, idx]
    np.save(os.path.join(result_dir, f"err_{rank}.npy"), np.array([np.abs(O - ref).max()]))
    dist.destroy_process_group()


def _worker_bwd(rank, world, port, seq, d, v_d, batch, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(12)
    Q, K, V, dO = da.random_inputs(rng, np.float64, (batch,), d, v_d, (seq,), (seq,))
    mask = pattern.tests_mask((seq,), (seq,), "none_front", "causal")
    O, l, m = da.forward(Q, K, V, mask)
    ref = dict(zip(("dQ", "dK", "dV"), da.backward(Q, K, V, mask, dO)))
    layout = ring.ZigZag(seq, world)
    idx = layout.gather_index(rank)
    c = layout.chunk

    def halves(X):
        return [torch.from_numpy(np.ascontiguousarray(X[..., idx[:c]])), torch.from_numpy(np.ascontiguousarray(X[..., idx[c:]]))]
    d_q, d_kv = ring.ring_backward(OracleBackend(batch, d, v_d, c), layout, rank, halves(Q), halves(K) + halves(V),
                                   halves(O), halves(l), halves(m), halves(dO), dist, None, "causal")
    got = {"dQ": np.concatenate([x.numpy() for x in d_q], axis=-1),
           "dK": np.concatenate([x.numpy() for x in d_kv[:2]], axis=-1),
           "dV": np.concatenate([x.numpy() for x in d_kv[2:]], axis=-1)}
    errs = [np.abs(got[n] - ref[n][:, :, idx]).max() for n in ("dQ", "dK", "dV")]
    np.save(os.path.join(result_dir, f"berr_{rank}.npy"), np.array(errs))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ring_backward_driver_matches_dense_oracle(world, tmp_path):
    """dQ stays local, the dK / dV accumulators travel with their shard and come home after `world` hops."""
    seq = 16 * world
    mp.spawn(_worker_bwd, args=(world, _free_port(), seq, 8, 6, 2, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert float(np.load(tmp_path / f"berr_{r}.npy").max()) < 1e-12


@pytest.mark.parametrize("world", [2, 4])
def test_ring_driver_matches_dense_oracle(world, tmp_path):
    seq = 16 * world
    mp.spawn(_worker, args=(world, _free_port(), seq, 8, 6, 2, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert float(np.load(tmp_path / f"err_{r}.npy")[0]) < 1e-12


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_step_plan_covers_the_causal_triangle_once(world):
    layout = ring.ZigZag(64 * world, world)
    seen = {}
    per_rank_blocks = []
    for rank in range(world):
        n = 0
        for step in range(world):
            src, plan = ring.step_plan(layout, rank, step, "causal", False)
            for qa, kb, qb, kb_base in plan:
                key = (qb // layout.chunk, kb_base // layout.chunk)
                assert key not in seen
                seen[key] = (rank, step)
                n += 1
        per_rank_blocks.append(n)
    want = {(a, b) for a in range(2 * world) for b in range(2 * world) if b <= a}
    assert set(seen) == want
    # zig-zag balance: every rank launches the same number of blocks
    assert len(set(per_rank_blocks)) == 1
    # ownership is a partition of the sequence
    allidx = np.sort(np.concatenate([layout.gather_index(r) for r in range(world)]))
    assert np.array_equal(allidx, np.arange(64 * world))


def test_batch_head_sharding_is_embarrassingly_parallel():
    """bench.py --gpus N gives every rank the same independent batch; nothing to exchange. This pins the
    host-side invariant the scaling number relies on: flops are additive over ranks."""
    from bench import WORKLOADS, flops_of
    w = WORKLOADS["C2"]
    f1 = sum(flops_of(w, 33558528))
    assert abs(f1 - (4.398583382016e12 + 10.99645845504e12)) / f1 < 1e-9
