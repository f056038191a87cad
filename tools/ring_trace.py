"""Developer tool (torchrun, one rank per GPU): per-step host / device timeline of the batched causal ring forward.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ring_trace.py [chunk] [heads]
chunk = positions per chunk (the C5 run on 8 GPUs has chunk 8192, 16 heads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tf_flash_attention_b200 import ring

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = torch.Generator(device="cuda").manual_seed(rank)
Q, K, V = ((torch.rand((1, heads, 128, 2 * chunk), generator=g, device="cuda") * 4 - 2).half() for _ in range(3))
dsetup = ring._causal_ring_setup(Q, V, "none_front", None)
d, rk, layout, backend = dsetup
q2, k2, v2 = (ring._chunk_major(x, layout.chunk) for x in (Q, K, V))
for it in range(4):
    if it == 3:
        backend.trace = ring.StepTrace(torch)
    dist.barrier()
    torch.cuda.synchronize()
    ring.ring_forward_causal(backend, layout, rk, q2, k2, v2, d, None)
    torch.cuda.synchronize()
rep = backend.trace.report()
for r in range(world):
    dist.barrier()
    if r == rank and rank in (0, world - 1):
        print(f"rank {rank}: (mark, host ms, device ms)")
        for row in rep:
            print("   ", row)
        sys.stdout.flush()
dist.destroy_process_group()
