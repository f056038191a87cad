"""Multi-GPU check of the K/V ring (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ring_check.py
Every rank builds the same full random tensors, takes its zig-zag shard, runs ring_causal_1d over NCCL
and compares its rows of O with the dense oracle."""
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
    Q, K, V, _ = da.random_inputs(rng, dtype, (2,), d, d, (seq,), (seq,))
    ref = da.attention(Q, K, V, 1, "causal", "none_front")["O"]
    idx = ring.ZigZag(seq, world).gather_index(rank)
    sh = [torch.from_numpy(np.ascontiguousarray(X[:, :, idx])).cuda() for X in (Q, K, V)]
    O = ring.ring_causal_1d(*sh)
    torch.cuda.synchronize()
    err = float(np.abs(O.cpu().numpy().astype(np.float64) - ref[:, :, idx]).max())
    print(f"rank {rank}/{world} {np.dtype(dtype).name} seq {seq}: max err {err:.3e} (tol {tol})", flush=True)
    ok &= err <= tol
t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("RING_CHECK", "PASS" if int(t.item()) else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
