"""all_to_all_single bandwidth of the remap exchange on this node (torchrun, one rank per GPU):
   torchrun --nproc-per-node 8 scripts/a2a_probe.py [GiB per rank]
Prints GB/s per direction per GPU (bytes a rank sends to its peers / time) and the NCCL_* variables in effect."""
import os
import sys

import torch
import torch.distributed as dist

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G = dist.get_world_size()
n = int(gib * (1 << 30)) // 8 // G * G
src = torch.empty(n, dtype=torch.complex64, device="cuda").view(torch.float32)
dst = torch.empty_like(src)
src.zero_()
for _ in range(2):
    dist.all_to_all_single(dst, src)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    dist.all_to_all_single(dst, src)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
t = torch.tensor([ms], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if dist.get_rank() == 0:
    sent = src.numel() * 4 * (G - 1) / G
    env = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
    print("a2a %d ranks, %.1f GiB per rank: %.2f ms, %.1f GB/s per direction per GPU  %s" % (G, gib, float(t), sent / (float(t) * 1e-3) / 1e9, env), flush=True)
dist.destroy_process_group()
