"""Data-parallel plumbing: one process per GPU, workers (environments, episodic memories, rollout
buffer) sharded across ranks, model replicated.  The only data-path collectives per optimiser step
are a 3-double all-reduce of the advantage statistics (so normalisation is over the *global*
minibatch, reference trainer.py:285) and ONE sum all-reduce of the flat fp32 gradient arena
(NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU unit tests).  Every rank then applies the same
clip + AdamW, so replicas stay bit-identical."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend, rank=int(os.environ["RANK"]), world_size=world)


def shard_workers(n_workers, rank, world_size):
    """Contiguous worker range owned by ``rank``: [rank*W/G, (rank+1)*W/G)."""
    if n_workers % world_size != 0:
        raise ValueError("n_workers (%d) must be divisible by the number of ranks (%d)" % (n_workers, world_size))
    per = n_workers // world_size
    return range(rank * per, (rank + 1) * per)


class DataParallelContext:
    def __init__(self, device=None):
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank = dist.get_rank() if self.enabled else 0
        self.world_size = dist.get_world_size() if self.enabled else 1
        self.device = device

    def all_reduce_(self, tensor):
        """In-place sum over ranks (no-op for a single rank)."""
        if self.enabled:
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
        return tensor

    def broadcast_(self, tensor, src=0):
        if self.enabled:
            dist.broadcast(tensor, src=src)
        return tensor

    def barrier(self):
        if self.enabled:
            dist.barrier()

    def max_(self, tensor):
        if self.enabled:
            dist.all_reduce(tensor, op=dist.ReduceOp.MAX)
        return tensor
