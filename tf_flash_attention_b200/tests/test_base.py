"""Verification / benchmark harness shipped with the package, with the reference's command line.

The reference ships `flash_attention/tests/{test_base,test_1d,test_2d}.py` (TensorFlow test cases run as
`python -m flash_attention.tests.test_1d TestGroup.{list,verify,benchmark}`, one case selectable with
`TESTCASE=<name>`; test_base.py:294-409). This module offers the same entry points, case names, shape
distributions (test_1d.py:57-66, test_2d.py:85-94), U(-2,2) data (test_base.py:170-173), 20 random shapes per
dtype (test_base.py:108) and tolerances (`1e-6 * N` for float / double, `1e-3 * N` for half, N = number of
summed entries; test_base.py:199-226) against a dense masked-softmax attention.

It is written against torch CUDA tensors (TensorFlow is absent from this image) and is table driven rather
than a class lattice: a case is (rule kind, sync mode); locations, masks and the dense attention are the plain
functions below. The dense side runs in float64 from the dtype-rounded inputs, so the comparison is at least as
strict as the reference's same-dtype dense computation. Needs a CUDA device (there is no CPU fallback).

Environment: TESTCASE=<name> (default all), RUNS=<n> random shapes per dtype (default 20).
"""
import os
import sys
import time

import torch

from .. import flash_attention

random_seed = 1234

SYNC_MODES = ("none_front", "scale_front", "scale_end")
# kind -> (entry-point family, strided window, causal flag of the local rule)
KINDS = {
    "Full": ("full", False, False),
    "Causal": ("causal", False, False),
    "Local": ("local", False, False),
    "LocalStride": ("local", True, False),
    "LocalAndCausal": ("local", False, True),
    "LocalStrideAndCausal": ("local", True, True),
}
_SYNC_TITLE = {"none_front": "NoneFront", "scale_front": "ScaleFront", "scale_end": "ScaleEnd"}
_DTYPES = (torch.float16, torch.float32, torch.float64)
# reference tolerances per summed entry (test_base.py:207-209)
_TOL = {torch.float16: 1e-3, torch.float32: 1e-6, torch.float64: 1e-6}


def case_names():
    """The reference's case list (test_base.py:364-385) plus CausalAttentionSyncModeNoneFront, which the reference
    defines (:317-318) but leaves out of its table."""
    names = ["FullAttentionSyncModeNoneFront", "CausalAttentionSyncModeNoneFront"]
    names += [f"CausalAttentionSyncMode{_SYNC_TITLE[s]}" for s in SYNC_MODES[1:]]
    for kind in ("Local", "LocalStride", "LocalAndCausal", "LocalStrideAndCausal"):
        names += [f"{kind}AttentionSyncMode{_SYNC_TITLE[s]}" for s in SYNC_MODES]
    return names


def parse_case(name):
    kind, _, sync = name.partition("AttentionSyncMode")
    sync_mode = {v: k for k, v in _SYNC_TITLE.items()}[sync]
    return kind, sync_mode


def locations(q_seq, k_seq, sync_mode, device):
    """Per-dimension grid coordinates of every Q and K entry (flattened row-major) and their linear order in the shared
    grid, as the reference's tests define the three sync modes (test_1d.py:9-50, test_2d.py:11-78)."""
    dims = len(q_seq)
    mx = [max(a, b) for a, b in zip(q_seq, k_seq)]

    def one(seq):
        axes = []
        for d in range(dims):
            step = 1 if sync_mode == "none_front" else mx[d] // seq[d]
            r = torch.arange(seq[d], device=device, dtype=torch.int64)
            axes.append((r + 1) * step - 1 if sync_mode == "scale_end" else r * step)
        grid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, dims)
        order = grid[:, 0]
        for d in range(1, dims):
            order = order * mx[d] + grid[:, d]
        return grid, order
    return one(q_seq), one(k_seq)


def attended(kind, q_loc, k_loc, q_seq, k_seq):
    """Boolean [nq, nk] mask plus the (window_size, log2_stride_size) the local cases pass to the op. The window is the
    longest sequence axis of either side and the strided cases use the largest power of two not above it, as the
    reference's VanillaLocalPolicy does (test_base.py:43-67: `max(diff.shape)`, `int(log2(window))`)."""
    family, strided, causal = KINDS[kind]
    (q_grid, q_order), (k_grid, k_order) = q_loc, k_loc
    ahead = q_order[:, None] >= k_order[None, :]
    if family == "full":
        return torch.ones_like(ahead), None, None
    if family == "causal":
        return ahead, None, None
    window = max(list(q_seq) + list(k_seq) + [len(q_seq), 1])
    log2_stride = (window.bit_length() - 1) if strided else 0
    stride = 1 << log2_stride
    diff = (q_grid[:, None, :] - k_grid[None, :, :]).abs()
    near = ((diff % stride == 0) & (diff // stride < window)).all(dim=-1)
    return (near & ahead) if causal else near, window, log2_stride


def dense_attention(Q, K, V, mask, seq_dims):
    """softmax(mask(Q^T K / sqrt(d))) V on flattened sequences, rows without an attended key give 0
    (test_1d.py:69-76, test_2d.py:97-109)."""
    q_seq = Q.shape[-seq_dims:]
    q = Q.flatten(-seq_dims)
    k = K.flatten(-seq_dims)
    v = V.flatten(-seq_dims)
    logit = torch.einsum("...cq,...ck->...qk", q, k) / (Q.shape[-seq_dims - 1] ** 0.5)
    logit = torch.where(mask, logit, torch.finfo(logit.dtype).min)
    p = torch.where(mask, torch.softmax(logit, dim=-1), torch.zeros((), dtype=logit.dtype, device=logit.device))
    return torch.einsum("...qk,...ck->...cq", p, v).unflatten(-1, q_seq)


class Case:
    """One (rule kind, sync mode) pair of a 1-D or 2-D group."""

    def __init__(self, name, seq_dims, shape_table):
        self.name, self.seq_dims, self.shape_table = name, seq_dims, shape_table
        self.kind, self.sync_mode = parse_case(name)

    def flash(self, Q, K, V, window, log2_stride):
        family, _, causal = KINDS[self.kind]
        fn = getattr(flash_attention, f"{family}_{self.seq_dims}d")
        if family == "local":
            return fn(Q, K, V, window_size=window, log2_stride_size=log2_stride, is_causal=causal,
                      sync_mode=self.sync_mode, returning_l_m=True)
        return fn(Q, K, V, sync_mode=self.sync_mode, returning_l_m=True)

    def _random_shape(self, gen, lo, hi, even):
        shape = [int(torch.randint(a, b + 1, (1,), generator=gen)) for a, b in zip(lo, hi)]
        if even:   # half: even innermost length (test_base.py:148-149)
            shape[-1] = shape[-1] // 2 * 2
        return shape

    def data(self, dtype, how, gens, device):
        shape_gen, gen = gens      # shapes are drawn on the host, data on the device
        lo, hi = self.shape_table[dtype]
        sd = self.seq_dims
        if how == "max":
            q_shape = k_shape = list(hi)
        else:
            k_shape = self._random_shape(shape_gen, lo, hi, dtype == torch.float16)
            q_shape = k_shape[:-sd] + self._random_shape(shape_gen, lo[-sd:], hi[-sd:], dtype == torch.float16)
        do_shape = q_shape[:-sd - 1] + k_shape[-sd - 1:-sd] + q_shape[-sd:]

        def u(shape):
            return (torch.rand(shape, generator=gen, device=device, dtype=torch.float32) * 4 - 2).to(dtype)
        Q, K, V, dO = u(q_shape), u(k_shape), u(k_shape), u(do_shape)
        q_loc, k_loc = locations(q_shape[-sd:], k_shape[-sd:], self.sync_mode, device)
        mask, window, log2_stride = attended(self.kind, q_loc, k_loc, q_shape[-sd:], k_shape[-sd:])
        return Q, K, V, dO, mask, window, log2_stride

    def verify(self, runs):
        device = torch.device("cuda", torch.cuda.current_device())
        gen = (torch.Generator().manual_seed(random_seed), torch.Generator(device=device).manual_seed(random_seed))
        for dtype in _DTYPES:
            worst = [0.0, 0.0]
            for _ in range(runs):
                Q, K, V, dO, mask, window, log2_stride = self.data(dtype, "random", gen, device)
                Q, K, V = (t.requires_grad_(True) for t in (Q, K, V))
                O = self.flash(Q, K, V, window, log2_stride)[0]
                grads = torch.autograd.grad(O, (Q, K, V), dO)
                Qd, Kd, Vd = (t.detach().double().requires_grad_(True) for t in (Q, K, V))
                Od = dense_attention(Qd, Kd, Vd, mask, self.seq_dims)
                grads_d = torch.autograd.grad(Od, (Qd, Kd, Vd), dO.double())
                nq, nk = mask.shape
                what = f"{self.name}: {dtype}, random_seed = {random_seed}, Q = {tuple(Q.shape)}, K = {tuple(K.shape)}"
                for label, got, ref, n in (("forward", O, Od, nk), ("backward dQ", grads[0], grads_d[0], nk),
                                           ("backward dK", grads[1], grads_d[1], nq),
                                           ("backward dV", grads[2], grads_d[2], nq)):
                    tol = _TOL[dtype] * n
                    err = (got.detach().double() - ref.detach()).abs()
                    bound = tol + tol * ref.detach().abs()
                    if not bool((err <= bound).all()):
                        raise AssertionError(f"{what}: {label} differs, max abs error {float(err.max()):.3e}, "
                                             f"tolerance {tol:.3e}")
                    worst[0 if label == "forward" else 1] = max(worst[0 if label == "forward" else 1], float(err.max()))
            print(f"  {str(dtype):15s} {runs} shapes ok   max |dO| {worst[0]:.2e}   max |dgrad| {worst[1]:.2e}", flush=True)

    def benchmark(self, runs, burn):
        device = torch.device("cuda", torch.cuda.current_device())
        gen = (torch.Generator().manual_seed(random_seed), torch.Generator(device=device).manual_seed(random_seed))
        report = {}
        for dtype in _DTYPES:
            Q, K, V, dO, mask, window, log2_stride = self.data(dtype, "max", gen, device)
            Q, K, V = (t.requires_grad_(True) for t in (Q, K, V))

            def timed(fn):
                for _ in range(burn):
                    fn()
                torch.cuda.synchronize()
                torch.cuda.reset_peak_memory_stats()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(runs):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) * 1e-3 / runs, torch.cuda.max_memory_allocated()
            flash_out = self.flash(Q, K, V, window, log2_stride)[0]
            dense_out = dense_attention(Q, K, V, mask, self.seq_dims)
            report[dtype] = {
                "forward": (timed(lambda: dense_attention(Q, K, V, mask, self.seq_dims)),
                            timed(lambda: self.flash(Q, K, V, window, log2_stride))),
                "backward": (timed(lambda: torch.autograd.grad(dense_out, (Q, K, V), dO, retain_graph=True)),
                             timed(lambda: torch.autograd.grad(flash_out, (Q, K, V), dO, retain_graph=True))),
            }
        return report


def show_benchmark_report(report):
    for name, per_dtype in report.items():
        for dtype, per_direction in per_dtype.items():
            for direction, (dense, flash) in per_direction.items():
                print(f"{name}: {direction}, {dtype}")
                print(f"{'':28s}{'dense (torch)':28s}flash")
                print(f"{'wall time [s]:':28s}{dense[0]:<28.6g}{flash[0]:.6g}")
                print(f"{'max # of bytes (GPU):':28s}{dense[1]:<28d}{flash[1]}\n")


class TestGroup:
    """All cases of one sequence rank; `list`, `verify`, `benchmark` as in the reference (test_base.py:387-409)."""
    SEQUENCE_DIMS = None
    SHAPE_TABLE = None
    RUNS = 20
    BURNING_RUNS = 6

    def __init__(self):
        self._cases = {n: Case(n, self.SEQUENCE_DIMS, self.SHAPE_TABLE) for n in case_names()}

    def _selected(self):
        name = os.environ.get("TESTCASE", "all")
        return self._cases if name == "all" else {name: self._cases[name]}

    def list(self):
        print("Available testcases:")
        for name in self._selected():
            print(name)

    def verify(self):
        runs = int(os.environ.get("RUNS", self.RUNS))
        for name, case in self._selected().items():
            print(f"Verifying {name}", flush=True)
            case.verify(runs)

    def benchmark(self):
        runs = int(os.environ.get("RUNS", self.RUNS))
        report = {}
        for name, case in self._selected().items():
            print(f"Benchmarking {name}", flush=True)
            report[name] = case.benchmark(runs, self.BURNING_RUNS)
        show_benchmark_report(report)


def main(group_cls, argv=None):
    """`python -m tf_flash_attention_b200.tests.test_1d TestGroup.verify` (also `.list`, `.benchmark`); like the
    reference's `__main__` block the seed is taken from the clock and printed (test_1d.py:141-146) unless SEED is set."""
    global random_seed
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1 or not argv[0].startswith("TestGroup.") or argv[0].split(".", 1)[1] not in ("list", "verify", "benchmark"):
        print("usage: python -m tf_flash_attention_b200.tests.test_{1d,2d} TestGroup.{list|verify|benchmark}", file=sys.stderr)
        return 2
    action = argv[0].split(".", 1)[1]
    if action != "list":
        if not torch.cuda.is_available():
            print("a CUDA device is required (there is no CPU fallback)", file=sys.stderr)
            return 1
        random_seed = int(os.environ.get("SEED", int(time.time())))
        print(f"random seed = {random_seed}")
    getattr(group_cls(), action)()
    return 0
