"""all-to-all bandwidth probe (torchrun): per-rank payload of --mb MB split evenly over the ranks,
as one all_to_all_single of raw bytes -- what the record / group exchange of multigpu.py issues."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=230)
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = (args.mb << 20) // world * world
send = torch.empty(n, dtype=torch.uint8, device="cuda")
recv = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    dist.all_to_all_single(recv, send)
torch.cuda.synchronize()
dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.iters):
    dist.all_to_all_single(recv, send)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / args.iters
t = torch.tensor([ms], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    off = n * (world - 1) / world
    print(f"a2a world={world} {args.mb} MB/rank: {t.item():.3f} ms  off-rank {off / 1e9 / (t.item() / 1e3):.1f} GB/s per GPU  "
          f"env NCCL_MIN_P2P_NCHANNELS={os.environ.get('NCCL_MIN_P2P_NCHANNELS')} NCCL_MAX_P2P_NCHANNELS={os.environ.get('NCCL_MAX_P2P_NCHANNELS')}",
          flush=True)
dist.destroy_process_group()
