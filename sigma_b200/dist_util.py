"""Multi-GPU bookkeeping of the hot path.  The forward path has NO data-path collective (the scan is per-sample:
SURVEY.md §8e) — ranks are replicas over disjoint image shards; the only exchange is the timing reduction below.
Training (train.py:103-108 of the reference) wraps the model in torch DDP; nothing custom is needed for that."""
import os

import torch
import torch.distributed as dist


def env_world():
    """(world_size, rank, local_rank) as set by torchrun."""
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")))


def init(backend, device=None):
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return world, rank, local


def shard_seed(base, rank):
    """Every rank draws its own synthetic image shard."""
    return int(base) + int(rank)


def max_over_ranks(value, device="cpu"):
    """MAX all-reduce of a python float (the slowest rank defines the step time)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_images_per_s(batch_per_rank, world, steps, total_ms_max):
    """Whole-job throughput: all ranks' images over the slowest rank's time."""
    return batch_per_rank * world * steps / (total_ms_max * 1e-3)
