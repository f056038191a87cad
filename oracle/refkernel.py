"""TEST / BASELINE INFRASTRUCTURE — ctypes driver of the UNMODIFIED reference CUDA kernel built by
oracle/ref_build/Makefile into oracle/_ref/libref_fa.so (harness: oracle/ref_build/ref_fa_harness.cu,
mirroring flash_attention_forward.cc:303-385 / flash_attention_backward.cc:260-341).

Used (a) to generate tests/golden/refkernel_*.npz on a B200 (pins the oracle's numerics against the
real reference), (b) as bench.py's "reference's own CuTe kernel on the same B200" baseline.
Never imported by the product."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_fa.so")
RULES = {"full": 0, "causal": 1, "local": 2}
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{LIB} not built (needs /root/reference; run make -C oracle/ref_build)")
        _lib = C.CDLL(LIB)
        _lib.ref_occupancy_elems.restype = C.c_int64
        _lib.ref_occupancy_elems.argtypes = [C.c_int, C.c_int, C.c_int64]
        vp, i, pi64 = C.c_void_p, C.c_int, C.POINTER(C.c_int64)
        _lib.ref_forward.restype = C.c_int
        _lib.ref_forward.argtypes = [i, i, i, C.c_char_p, i, i, i, i, i, i, pi64, pi64] + [vp] * 7 + [i, vp]
        _lib.ref_backward.restype = C.c_int
        _lib.ref_backward.argtypes = [i, i, i, C.c_char_p, i, i, i, i, i, i, pi64, pi64] + [vp] * 11 + [i, vp]
        _lib.ref_estimate_flops.restype = C.c_int
        _lib.ref_estimate_flops.argtypes = [i, i, i, C.c_char_p, i, i, i, i, i, i, pi64, pi64, i, C.POINTER(C.c_float)]
    return _lib


def _arr(v):
    return (C.c_int64 * len(v))(*[int(x) for x in v])


def estimate_flops(dtype_code, dims, rule, sync_mode, w, s, c, b, d, v_d, qs, ks, smem_optin=232448):
    f = C.c_float(0)
    rc = lib().ref_estimate_flops(dtype_code, dims, RULES[rule], sync_mode.encode(), w, s, int(c), b, d, v_d,
                                  _arr(qs), _arr(ks), smem_optin, C.byref(f))
    if rc:
        raise RuntimeError(f"ref_estimate_flops rc={rc}")
    return f.value


def forward(Q, K, V, dims, rule, sync_mode, w=1, s=0, c=False, smem_optin=232448):
    """torch CUDA tensors [B, ch, seq...] -> (O, l, m) from the reference kernel."""
    import torch
    code = {torch.float16: 0, torch.float32: 1, torch.float64: 2}[Q.dtype]
    qs, ks = tuple(Q.shape[-dims:]), tuple(K.shape[-dims:])
    b, d, v_d = int(np.prod(Q.shape[:-dims - 1])), Q.shape[-dims - 1], V.shape[-dims - 1]
    q = int(np.prod(qs))
    O = torch.empty(tuple(Q.shape[:-dims - 1]) + (v_d,) + qs, dtype=Q.dtype, device=Q.device)
    ldt = torch.float32 if Q.dtype == torch.float16 else Q.dtype
    l = torch.empty(tuple(Q.shape[:-dims - 1]) + qs, dtype=ldt, device=Q.device)
    m = torch.empty(tuple(Q.shape[:-dims - 1]) + qs, dtype=Q.dtype, device=Q.device)
    occ = torch.empty(lib().ref_occupancy_elems(code, b, q), dtype=torch.int32, device=Q.device)
    rc = lib().ref_forward(code, dims, RULES[rule], sync_mode.encode(), w, s, int(c), b, d, v_d, _arr(qs), _arr(ks),
                           Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), l.data_ptr(), m.data_ptr(),
                           occ.data_ptr(), smem_optin, torch.cuda.current_stream().cuda_stream)
    if rc:
        raise RuntimeError(f"ref_forward rc={rc}")
    return O, l, m


def backward(Q, K, V, O, l, m, dO, dims, rule, sync_mode, w=1, s=0, c=False, smem_optin=232448):
    import torch
    code = {torch.float16: 0, torch.float32: 1, torch.float64: 2}[Q.dtype]
    qs, ks = tuple(Q.shape[-dims:]), tuple(K.shape[-dims:])
    b, d, v_d = int(np.prod(Q.shape[:-dims - 1])), Q.shape[-dims - 1], V.shape[-dims - 1]
    q = int(np.prod(qs))
    dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
    occ = torch.empty(lib().ref_occupancy_elems(code, b, q), dtype=torch.int32, device=Q.device)
    rc = lib().ref_backward(code, dims, RULES[rule], sync_mode.encode(), w, s, int(c), b, d, v_d, _arr(qs), _arr(ks),
                            Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), l.data_ptr(), m.data_ptr(),
                            dO.data_ptr(), dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(), occ.data_ptr(), smem_optin,
                            torch.cuda.current_stream().cuda_stream)
    if rc:
        raise RuntimeError(f"ref_backward rc={rc}")
    return dQ, dK, dV


def bench(w, nnz, heads=8, fwd_only=False, iters=2):
    """Times the reference kernel on `heads` heads of workload `w` (bench.py WORKLOADS entry)."""
    import torch
    tdt = {"float16": torch.float16, "float32": torch.float32, "float64": torch.float64}[w["dtype"]]
    dims = w["seq_dims"]
    g = torch.Generator(device="cuda").manual_seed(99)

    def u(shape):
        return (torch.rand(shape, generator=g, device="cuda", dtype=torch.float32) * 4 - 2).to(tdt)
    Q, K = u((heads, w["d"]) + w["q"]), u((heads, w["d"]) + w["k"])
    V, dO = u((heads, w["v_d"]) + w["k"]), u((heads, w["v_d"]) + w["q"])
    args = (dims, w["rule"], w["sync"], w["w"], w["s"], bool(w["c"]))
    O, l, m = forward(Q, K, V, *args)
    if not fwd_only:
        backward(Q, K, V, O, l, m, dO, *args)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(iters):
        e[0].record()
        O, l, m = forward(Q, K, V, *args)
        e[1].record()
        if not fwd_only:
            backward(Q, K, V, O, l, m, dO, *args)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]) / iters
        tb += e[1].elapsed_time(e[2]) / iters
    ff = 2.0 * nnz * (w["d"] + w["v_d"]) * heads
    fb = 2.0 * nnz * (3 * w["d"] + 2 * w["v_d"]) * heads
    out = {"what": "reference's own CuTe/SIMT kernel (unmodified, built for sm_100) on this B200",
           "sample": f"{heads} heads of the same workload (incl. its 4 memsets)",
           "fwd_tflops": ff / (tf * 1e-3) / 1e12, "fwd_ms": tf}
    if not fwd_only:
        out.update(bwd_tflops=fb / (tb * 1e-3) / 1e12, bwd_ms=tb,
                   value=(ff + fb) / ((tf + tb) * 1e-3) / 1e12, unit="TFLOPS")
    return out
