"""Multi-GPU extraction: one process per GPU (torchrun), utterances sharded ``i -> rank i mod G`` with no
data-path collective, then ONE NCCL all-gather of the (N/G, E) fp32 embedding blocks so every rank holds the
full (N,E) matrix for scoring.  This replaces the reference's process-level sharding through
the filesystem (`tools/extract_embedding.sh:39-73`: split list, nj processes, ``cat xvector_*.scp``).
For PLDA, enroll rows are sharded and outputs stay rank-local (no output collective)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/MASTER_*); returns (rank, world, local)."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_indices(n: int, rank: int, world: int):
    """Utterance i -> rank i mod world (deterministic, restores order after the gather)."""
    return list(range(rank, n, world))


def shard_rows(n: int, rank: int, world: int):
    """Contiguous row block [lo, hi) of rank (PLDA enroll sharding)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_embeddings(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """all_gather_into_tensor of per-rank (n_r, E) blocks (padded to equal counts) and re-interleave to the
    original utterance order of `shard_indices`.  Works with nccl (CUDA tensors) and gloo (CPU tensors)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = (n_total + world - 1) // world
    E = local.shape[1]
    padded = torch.zeros((per, E), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * per, E), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    # rank r row j is utterance j*world + r
    full = out.view(world, per, E).transpose(0, 1).reshape(world * per, E)
    return full[:n_total].contiguous()


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
