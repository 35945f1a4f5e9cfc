"""N-rank check (NCCL): sample_sharded over all ranks == the same labels sampled on one rank (Philox stream is keyed by
the global sample index), and every rank receives the full gathered tensor."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from spectrogramgenai_b200.diff_modules import Diffusion
from spectrogramgenai_b200.sharding import sample_sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(42)
d = Diffusion(noise_steps=6, img_size=64, num_classes=27, c_in=4, c_out=4, device=dev, compute_dtype="bf16")
labels = torch.arange(11) % 27
full = sample_sharded(d, labels, 3, seed=5)
single = d.sample(False, labels, 3, seed=5)
ok = torch.equal(full, single) and full.shape == (11, 4, 64, 64) and full.dtype == torch.uint8
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"multi_gpu_check world={world}: sharded == single-rank result on every rank: {bool(flag.item())}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
