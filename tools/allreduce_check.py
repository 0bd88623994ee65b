"""N-GPU check of fav_allreduce (run under torchrun): every rank fills its arena with rank-dependent integers; after the
library's NCCL all-reduce all ranks hold the same sums as torch.distributed computes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, fav
from fav.sweep import CorruptionSweep, SweepConfig
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sw = CorruptionSweep(SweepConfig(corruptions=("gaussian_noise", "contrast"), severities=(1, 3), T=2, block=16), device=local)
g = torch.Generator(device="cpu").manual_seed(rank)
sw.acc.arena.copy_(torch.randint(0, 1 << 40, sw.acc.arena.shape, generator=g, dtype=torch.int64))
want = sw.acc.arena.clone()
dist.all_reduce(want)
sw.acc.allreduce()
torch.cuda.synchronize()
ok = torch.equal(sw.acc.arena, want)
print(f"rank {rank}/{world}: fav_allreduce == torch.distributed: {ok}, words {sw.acc.arena.numel()}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
