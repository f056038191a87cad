"""K/V ring for ONE long 1-D sequence sharded over the GPUs of a node (BASELINE.json configs[4]).

New functionality — the reference is single-GPU. The sequence of length S is cut
into 2G chunks; rank r owns chunks r and 2G-1-r of Q, K and V ("zig-zag", which balances the causal
triangle). In G steps every rank attends its two local query chunks to the K/V shard currently
visiting (global index bases go into fa_problem_t, so the rule is evaluated on global coordinates and
blocks that lie entirely in the future are never launched), folds the partial (O, l, m) into fp32
accumulators (fa_partial_merge) and meanwhile passes the shard on to rank r+1 with NCCL send/recv
(torch.distributed P2P, NVLink 5 / NVSwitch) into the other half of a double buffer.

The driver is written against a tiny backend interface so that the host logic (chunk ownership, step
schedule, index bases, skip rule) can be exercised on CPU with the gloo backend and the NumPy oracle
as the compute stand-in (tests/test_ring_gloo.py); the product backend is `DeviceBackend` below, which
calls the C ABI (fa_forward / fa_partial_merge / fa_partial_finalize).

Backward (`ring_backward`): every block (local query chunk x visiting key chunk) is one fa_backward call
made with the FINAL O, l, m of the whole sequence, so block results simply add. dQ accumulates locally in
fp32; the fp32 dK / dV accumulators travel around the ring together with their K/V shard (one extra hop
at the end brings them home), K/V one step ahead of the compute, the accumulators right behind it.
"""
import ctypes as C

import numpy as np

from . import _capi


class ZigZag:
    """Chunk ownership: 2G chunks of c = S / (2G) positions; rank r owns chunks (r, 2G-1-r)."""

    def __init__(self, seq_len, world):
        if seq_len % (2 * world):
            raise _capi.InvalidArgumentError(_capi.FA_EINVAL_SHAPE,
                                             f"sequence length {seq_len} must be a multiple of 2*world = {2 * world}")
        self.seq_len, self.world, self.chunk = seq_len, world, seq_len // (2 * world)

    def chunks(self, rank):
        return (rank, 2 * self.world - 1 - rank)

    def base(self, chunk_index):
        return chunk_index * self.chunk

    def gather_index(self, rank):
        """Global positions held by `rank`, in local order (for building / checking shards)."""
        a, b = self.chunks(rank)
        return np.concatenate([np.arange(self.base(a), self.base(a) + self.chunk),
                               np.arange(self.base(b), self.base(b) + self.chunk)])


def step_plan(layout, rank, step, rule, is_causal):
    """Blocks (local q chunk slot, visiting k chunk slot, q_base, k_base) to launch at `step`, when the
    K/V shard of rank (rank - step) mod G is visiting. Causal rules drop blocks entirely in the future."""
    src = (rank - step) % layout.world
    plan = []
    causal = rule == "causal" or (rule == "local" and is_causal)
    for qa, qc in enumerate(layout.chunks(rank)):
        for kb, kc in enumerate(layout.chunks(src)):
            if causal and kc > qc:
                continue  # every key of this block is later than every query (none_front, equal lengths)
            plan.append((qa, kb, layout.base(qc), layout.base(kc)))
    return src, plan


class DeviceBackend:
    """Product backend: torch CUDA tensors for memory, the C ABI for every computation."""

    def __init__(self, dtype_code, rule, sync_mode, window_size, log2_stride_size, is_causal, batch_shape, d, v_d,
                 chunk, seq_len):
        import torch
        self.torch = torch
        self.rule, self.chunk, self.seq_len = rule, chunk, seq_len
        self.batch_shape = tuple(batch_shape)
        self.p = _capi.make_problem(dtype_code, 1, rule, sync_mode, self.batch_shape + (d, chunk),
                                    self.batch_shape + (d, chunk), self.batch_shape + (v_d, chunk), window_size,
                                    log2_stride_size, is_causal)
        self.p.q_full_len = self.p.k_full_len = seq_len
        self.v_d = v_d
        self.acc_dtype = torch.float64 if dtype_code == _capi.FA_F64 else torch.float32
        self._ws = None

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def new_acc(self, like_q):
        t, dev = self.torch, like_q.device
        return (t.empty(self.batch_shape + (self.v_d, self.chunk), dtype=self.acc_dtype, device=dev),
                t.empty(self.batch_shape + (self.chunk,), dtype=self.acc_dtype, device=dev),
                t.empty(self.batch_shape + (self.chunk,), dtype=self.acc_dtype, device=dev))

    def new_out(self, like_q):
        t, dev = self.torch, like_q.device
        ldt = t.float32 if like_q.dtype == t.float16 else like_q.dtype
        return (t.empty(self.batch_shape + (self.v_d, self.chunk), dtype=like_q.dtype, device=dev),
                t.empty(self.batch_shape + (self.chunk,), dtype=ldt, device=dev),
                t.empty(self.batch_shape + (self.chunk,), dtype=like_q.dtype, device=dev))

    def attend_partial(self, q, k, v, q_base, k_base, out):
        self.p.q_index_base, self.p.k_index_base, self.p.accumulate = q_base, k_base, 0
        o, l, m = out
        need = _capi.lib.fa_workspace_bytes(C.byref(self.p), 0)
        if need and (self._ws is None or self._ws.numel() < need):
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=q.device)
        _capi.check(_capi.lib.fa_forward(C.byref(self.p), q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
                                         l.data_ptr(), m.data_ptr(), self._ws.data_ptr() if need else None, need,
                                         self._stream()), "fa_forward")

    def merge(self, part, acc, first):
        o, l, m = part
        oa, la, ma = acc
        _capi.check(_capi.lib.fa_partial_merge(C.byref(self.p), o.data_ptr(), l.data_ptr(), m.data_ptr(),
                                               oa.data_ptr(), la.data_ptr(), ma.data_ptr(), int(first),
                                               self._stream()), "fa_partial_merge")

    def finalize(self, acc, out):
        oa, la, ma = acc
        o, l, m = out
        _capi.check(_capi.lib.fa_partial_finalize(C.byref(self.p), oa.data_ptr(), la.data_ptr(), ma.data_ptr(),
                                                  o.data_ptr(), l.data_ptr(), m.data_ptr(), self._stream()),
                    "fa_partial_finalize")

    def empty_like_kv(self, kv):
        return [self.torch.empty_like(x) for x in kv]

    # ---- backward ---------------------------------------------------------------------------
    def new_grad_acc(self, like):
        return self.torch.zeros(like.shape, dtype=self.acc_dtype, device=like.device)

    def new_grad_part(self, q, k, v):
        return (self.torch.empty_like(q), self.torch.empty_like(k), self.torch.empty_like(v))

    def grad_partial(self, q, k, v, o, l, m, d_o, q_base, k_base, part):
        self.p.q_index_base, self.p.k_index_base, self.p.accumulate = q_base, k_base, 0
        need = _capi.lib.fa_workspace_bytes(C.byref(self.p), 1)
        if need and (self._ws is None or self._ws.numel() < need):
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=q.device)
        dq, dk, dv = part
        _capi.check(_capi.lib.fa_backward(C.byref(self.p), q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
                                          l.data_ptr(), m.data_ptr(), d_o.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                          dv.data_ptr(), self._ws.data_ptr() if need else None, need, self._stream()),
                    "fa_backward")

    def grad_add(self, part, acc):
        _capi.check(_capi.lib.fa_grad_accumulate(self.p.dtype, part.data_ptr(), acc.data_ptr(), part.numel(), 0,
                                                 self._stream()), "fa_grad_accumulate")

    def grad_finalize(self, acc, like):
        out = self.torch.empty_like(like)
        _capi.check(_capi.lib.fa_grad_finalize(self.p.dtype, acc.data_ptr(), out.data_ptr(), acc.numel(),
                                               self._stream()), "fa_grad_finalize")
        return out


def ring_forward(backend, layout, rank, q_chunks, kv_chunks, dist=None, group=None, rule="causal", is_causal=False):
    """q_chunks: [Q_a, Q_b] local query chunks, each [batch..., d, c] contiguous.
    kv_chunks: [K_a, K_b, V_a, V_b] local key/value chunks. Returns [(O, l, m) for the two chunks]."""
    world = layout.world
    cur = list(kv_chunks)
    nxt = backend.empty_like_kv(cur) if world > 1 else None
    acc = [backend.new_acc(q_chunks[0]), backend.new_acc(q_chunks[1])]
    part = backend.new_out(q_chunks[0])
    started = [False, False]
    for step in range(world):
        reqs = []
        if world > 1 and step + 1 < world:
            ops = []
            for t_send, t_recv in zip(cur, nxt):
                ops.append(dist.P2POp(dist.isend, t_send, (rank + 1) % world, group))
                ops.append(dist.P2POp(dist.irecv, t_recv, (rank - 1) % world, group))
            reqs = dist.batch_isend_irecv(ops)
        _, plan = step_plan(layout, rank, step, rule, is_causal)
        for qa, kb, q_base, k_base in plan:
            backend.attend_partial(q_chunks[qa], cur[kb], cur[2 + kb], q_base, k_base, part)
            backend.merge(part, acc[qa], not started[qa])
            started[qa] = True
        for r in reqs:
            r.wait()
        if world > 1 and step + 1 < world:
            cur, nxt = nxt, cur
    outs = []
    for qa in range(2):
        out = backend.new_out(q_chunks[qa])
        if not started[qa]:  # cannot happen for causal/full (the diagonal block is local), kept for local rules
            backend.attend_partial(q_chunks[qa], kv_chunks[qa], kv_chunks[2 + qa], layout.base(layout.chunks(rank)[qa]),
                                   layout.base(layout.chunks(rank)[qa]), part)
            backend.merge(part, acc[qa], True)
        backend.finalize(acc[qa], out)
        outs.append(out)
    return outs


def ring_backward(backend, layout, rank, q_chunks, kv_chunks, o_chunks, l_chunks, m_chunks, do_chunks, dist=None,
                  group=None, rule="causal", is_causal=False):
    """Gradients of ring_forward. q/o/l/m/do_chunks: the two local query chunks (O, l, m as ring_forward
    returned them, i.e. final over the whole sequence); kv_chunks: [K_a, K_b, V_a, V_b].
    Returns ([dQ_a, dQ_b], [dK_a, dK_b, dV_a, dV_b]) for the local chunks."""
    world = layout.world
    cur = list(kv_chunks)
    nxt = backend.empty_like_kv(cur) if world > 1 else None
    dq_acc = [backend.new_grad_acc(q) for q in q_chunks]
    dkv_acc = [backend.new_grad_acc(x) for x in cur]          # travels with `cur`
    dkv_nxt = [backend.new_grad_acc(x) for x in cur] if world > 1 else None
    part = backend.new_grad_part(q_chunks[0], cur[0], cur[2])
    for step in range(world):
        reqs = []
        if world > 1 and step + 1 < world:                     # K/V shard for the next step, overlapped
            ops = []
            for t_send, t_recv in zip(cur, nxt):
                ops.append(dist.P2POp(dist.isend, t_send, (rank + 1) % world, group))
                ops.append(dist.P2POp(dist.irecv, t_recv, (rank - 1) % world, group))
            reqs = dist.batch_isend_irecv(ops)
        _, plan = step_plan(layout, rank, step, rule, is_causal)
        for qa, kb, q_base, k_base in plan:
            backend.grad_partial(q_chunks[qa], cur[kb], cur[2 + kb], o_chunks[qa], l_chunks[qa], m_chunks[qa],
                                 do_chunks[qa], q_base, k_base, part)
            backend.grad_add(part[0], dq_acc[qa])
            backend.grad_add(part[1], dkv_acc[kb])
            backend.grad_add(part[2], dkv_acc[2 + kb])
        for r in reqs:
            r.wait()
        if world > 1:
            # the accumulators follow their shard (after the last step this hop returns them to the owner)
            ops = []
            for t_send, t_recv in zip(dkv_acc, dkv_nxt):
                ops.append(dist.P2POp(dist.isend, t_send, (rank + 1) % world, group))
                ops.append(dist.P2POp(dist.irecv, t_recv, (rank - 1) % world, group))
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            dkv_acc, dkv_nxt = dkv_nxt, dkv_acc
            if step + 1 < world:
                cur, nxt = nxt, cur
    d_q = [backend.grad_finalize(dq_acc[i], q_chunks[i]) for i in range(2)]
    d_kv = [backend.grad_finalize(dkv_acc[i], kv_chunks[i]) for i in range(4)]
    return d_q, d_kv


def ring_causal_1d_backward(Q, K, V, O, l, m, dO, sync_mode="none_front", group=None):
    """Gradients of ring_causal_1d: all tensors are this rank's zig-zag shards (chunk r then chunk 2G-1-r);
    O, l, m as returned by ring_causal_1d(..., returning_l_m=True). Returns (dQ, dK, dV) shards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    layout = ZigZag(Q.shape[-1] * world, world)
    c = layout.chunk
    codes = {torch.float16: _capi.FA_F16, torch.float32: _capi.FA_F32, torch.float64: _capi.FA_F64}
    backend = DeviceBackend(codes[Q.dtype], "causal", sync_mode, 1, 0, False, Q.shape[:-2], Q.shape[-2], V.shape[-2], c,
                            layout.seq_len)

    def halves(x):
        return [x[..., :c].contiguous(), x[..., c:].contiguous()]
    d_q, d_kv = ring_backward(backend, layout, rank, halves(Q), halves(K) + halves(V), halves(O), halves(l), halves(m),
                              halves(dO), dist if world > 1 else None, group, "causal")
    return (torch.cat(d_q, dim=-1), torch.cat(d_kv[:2], dim=-1), torch.cat(d_kv[2:], dim=-1))


def ring_causal_1d(Q, K, V, sync_mode="none_front", group=None, returning_l_m=False):
    """causal_1d on a sequence sharded zig-zag over the ranks of `group` (default: world).
    Q, K, V: this rank's shard, torch CUDA tensors `batch_shape + (channel, 2*c)` holding chunk r then chunk
    2G-1-r. Returns O (and l, m) in the same sharded layout."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    c2 = Q.shape[-1]
    layout = ZigZag(c2 * world, world)
    c = layout.chunk
    codes = {torch.float16: _capi.FA_F16, torch.float32: _capi.FA_F32, torch.float64: _capi.FA_F64}
    backend = DeviceBackend(codes[Q.dtype], "causal", sync_mode, 1, 0, False, Q.shape[:-2], Q.shape[-2], V.shape[-2], c,
                            layout.seq_len)
    qs = [Q[..., :c].contiguous(), Q[..., c:].contiguous()]
    kv = [K[..., :c].contiguous(), K[..., c:].contiguous(), V[..., :c].contiguous(), V[..., c:].contiguous()]
    outs = ring_forward(backend, layout, rank, qs, kv, dist if world > 1 else None, group, "causal")
    O = torch.cat([outs[0][0], outs[1][0]], dim=-1)
    if not returning_l_m:
        return O
    return O, torch.cat([outs[0][1], outs[1][1]], dim=-1), torch.cat([outs[0][2], outs[1][2]], dim=-1)
