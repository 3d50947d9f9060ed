"""Data-parallel plumbing: one process per GPU, workers (environments, episodic memories, rollout
buffer) sharded across ranks, model replicated.  Data-path collectives: ONE in-place sum all-reduce per
optimiser step -- the flat fp32 gradient arena with the six loss statistics riding in its tail -- plus
one small all-reduce per epoch of every minibatch's advantage statistics (sum, sum of squares, count as
doubles; the epoch's permutation is known up front), so normalisation is over the *global* minibatch
(reference trainer.py:285).  Every rank then applies the same clip + AdamW, so replicas stay
bit-identical.

On GPUs the exchange goes through libtrxlppo's own C-ABI communicator (``trxl_comm_create`` /
``trxl_allreduce_grads``: NCCL over NVLink/NVSwitch, enqueued on the compute stream, so it is ordered
with the kernels without host synchronisation); torch.distributed only carries the 128-byte rendezvous id.
The CPU unit tests (gloo) and ``TRXL_NATIVE_NCCL=0`` use torch.distributed collectives instead."""
import os

import torch
import torch.distributed as dist

import trxl_native as native


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


_NATIVE_COMMS = {}       # (device index, world size) -> communicator handle, shared by every context of this process


class DataParallelContext:
    def __init__(self, device=None):
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank = dist.get_rank() if self.enabled else 0
        self.world_size = dist.get_world_size() if self.enabled else 1
        self.device = device
        self.n_collectives = 0           # data-path collectives issued through this context
        self._comm = None
        if (self.enabled and device is not None and torch.device(device).type == "cuda" and dist.get_backend() == "nccl"
                and os.environ.get("TRXL_NATIVE_NCCL", "1") != "0"):
            self._comm = self._native_comm(torch.device(device))

    def _native_comm(self, device):
        index = device.index if device.index is not None else torch.cuda.current_device()
        key = (index, self.world_size)
        if key not in _NATIVE_COMMS:
            # rendezvous: rank 0's 128-byte id travels over the existing torch.distributed group
            ident = torch.zeros(native.COMM_ID_BYTES, dtype=torch.uint8, device=device)
            if self.rank == 0:
                ident.copy_(torch.frombuffer(bytearray(native.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(ident, src=0)
            with torch.cuda.device(device):
                _NATIVE_COMMS[key] = native.comm_create(bytes(ident.cpu().numpy().tobytes()), self.rank, self.world_size)
        return _NATIVE_COMMS[key]

    @property
    def native_nccl(self):
        return self._comm is not None

    def all_reduce_(self, tensor):
        """In-place sum over ranks (no-op for a single rank)."""
        if not self.enabled:
            return tensor
        self.n_collectives += 1
        if self._comm is not None and tensor.is_cuda and tensor.is_contiguous() and tensor.dtype in (torch.float32, torch.float64):
            if tensor.dtype == torch.float32:
                native.allreduce_grads(self._comm, tensor)
            else:
                native.allreduce_f64(self._comm, tensor)
        else:
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
