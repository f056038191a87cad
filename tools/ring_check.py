"""Multi-GPU check of the K/V ring (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ring_check.py
Every rank builds the same full random tensors, takes its zig-zag shard, runs ring_causal_1d and ring_causal_1d_backward
over NCCL and compares its rows of O, dQ, dK, dV with the dense oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oracle import dense_attention as da
from tf_flash_attention_b200 import ring

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for dtype, d, seq, tol in ((np.float16, 128, 1024 * world, 2e-3), (np.float32, 32, 128 * world, 1e-5)):
    rng = np.random.default_rng(5)
    Q, K, V, dO = da.random_inputs(rng, dtype, (2,), d, d, (seq,), (seq,))
    full = da.attention(Q, K, V, 1, "causal", "none_front", dO=dO)
    ref = full["O"]
    idx = ring.ZigZag(seq, world).gather_index(rank)
    sh = [torch.from_numpy(np.ascontiguousarray(X[:, :, idx])).cuda() for X in (Q, K, V)]
    O, l, m = ring.ring_causal_1d(*sh, returning_l_m=True)
    torch.cuda.synchronize()
    err = float(np.abs(O.cpu().numpy().astype(np.float64) - ref[:, :, idx]).max())
    print(f"rank {rank}/{world} {np.dtype(dtype).name} seq {seq}: max err {err:.3e} (tol {tol})", flush=True)
    ok &= err <= tol
    # backward ring: dQ local, dK / dV accumulators travel with their shard
    tdo = torch.from_numpy(np.ascontiguousarray(dO[:, :, idx])).cuda()
    grads = ring.ring_causal_1d_backward(*sh, O, l, m, tdo)
    torch.cuda.synchronize()
    gtol = tol
    for name, g in zip(("dQ", "dK", "dV"), grads):
        r = full[name][:, :, idx]
        gerr = float((np.abs(g.cpu().numpy().astype(np.float64) - r) / np.maximum(1.0, np.abs(r))).max())
        print(f"rank {rank}/{world} {np.dtype(dtype).name} seq {seq}: {name} max scaled err {gerr:.3e} (tol {gtol:.1e})", flush=True)
        ok &= gerr <= gtol
t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("RING_CHECK", "PASS" if int(t.item()) else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
