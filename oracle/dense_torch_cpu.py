"""TEST / BASELINE INFRASTRUCTURE — the reference's "CPU path", multi-threaded.

The reference kernel is GPU-only; its CPU-runnable statement of the same computation is the dense
masked-softmax attention of its own tests (flash_attention/tests/test_1d.py:69-76, test_2d.py:97-109:
einsum -> where(mask, logit, dtype.min) -> softmax -> where(mask, p, 0) -> einsum, gradients by
autodiff, test_base.py:185-194). TensorFlow is not installed in this image, so the same op sequence
is issued with torch CPU ops (all host threads). fp16 inputs are computed in fp32 on the CPU.
Only bench.py's cpu_baseline / --impl reference legs import this."""
import time

import numpy as np
import torch

from . import pattern


def mask_tensor(q_shape, k_shape, sync_mode, rule, window_size=1, log2_stride_size=0, is_causal=False):
    return torch.from_numpy(pattern.tests_mask(q_shape, k_shape, sync_mode, rule, window_size,
                                               log2_stride_size, is_causal))


def vanilla_attention(Q, K, V, mask):
    """Q [B,d,q] K [B,d,k] V [B,vd,k] mask [q,k] -> O [B,vd,q]  (test_1d.py:69-76)."""
    d = Q.shape[1]
    logit = torch.einsum("bcq,bck->bqk", Q, K) / float(np.sqrt(d))
    logit = torch.where(mask, logit, torch.finfo(logit.dtype).min)
    p = torch.softmax(logit, dim=-1)
    p = torch.where(mask, p, torch.zeros((), dtype=p.dtype))
    return torch.einsum("bqk,bck->bcq", p, V)


def time_fwd_bwd(batch, d, v_d, nq, nk, mask, steps=1, warmup=0, seed=0, dtype=torch.float32, backward=True):
    """Seconds per step of vanilla forward (+ autodiff backward) on `batch` independent heads."""
    g = torch.Generator().manual_seed(seed)
    Q = (torch.rand((batch, d, nq), generator=g, dtype=dtype) * 4 - 2).requires_grad_(backward)
    K = (torch.rand((batch, d, nk), generator=g, dtype=dtype) * 4 - 2).requires_grad_(backward)
    V = (torch.rand((batch, v_d, nk), generator=g, dtype=dtype) * 4 - 2).requires_grad_(backward)
    dO = torch.rand((batch, v_d, nq), generator=g, dtype=dtype) * 4 - 2
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        # one head at a time, like a memory-bounded host implementation would
        for b in range(batch):
            O = vanilla_attention(Q[b:b + 1], K[b:b + 1], V[b:b + 1], mask)
            if backward:
                torch.autograd.grad(O, (Q, K, V), dO[b:b + 1])
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
    return float(np.mean(times))
