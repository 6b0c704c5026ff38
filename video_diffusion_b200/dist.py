"""Multi-GPU plumbing: the hot path shards over independent (video, sample) tasks, one process
per GPU, with NO collective inside the sampling / ELBO loops (the reference launches one
process per GPU the same way: scripts/video_sample.py:577-593, command_launchers.py:32-62).
The only communication is one final gather of uint8 samples / fp32 ELBO arrays over
NCCL (NVLink) -- or gloo in the CPU tests."""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get('RANK', '0')), world
    backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def shard_tasks(n_tasks, rank, world):
    """Task t (one batch of consecutive test videos, like the reference's --task_id) goes to rank t % world."""
    return list(range(rank, n_tasks, world))


def task_video_indices(task_id, batch_size, n_videos):
    """indices = range(task_id*bs, (task_id+1)*bs) clipped to the dataset (scripts/video_sample.py:577-582)."""
    return list(range(task_id * batch_size, min((task_id + 1) * batch_size, n_videos)))


def gather_ragged(local, ids, dst=None):
    """All-gather per-video rows from every rank.

    local: (n_local, ...) tensor (uint8 samples or fp32 ELBO rows); ids: (n_local,) int64 global
    video ids.  Counts differ per rank, so rows are padded to the maximum count.  Returns
    (rows, ids) sorted by video id on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        order = torch.argsort(ids)
        return local[order], ids[order]
    world = dist.get_world_size()
    dev = local.device
    n = torch.tensor([local.shape[0]], device=dev, dtype=torch.long)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    n_max = int(max(int(c.item()) for c in counts))
    pad_rows = torch.zeros((n_max,) + tuple(local.shape[1:]), device=dev, dtype=local.dtype)
    pad_rows[:local.shape[0]] = local
    pad_ids = torch.full((n_max,), -1, device=dev, dtype=torch.long)
    pad_ids[:ids.shape[0]] = ids.to(dev)
    all_rows = [torch.empty_like(pad_rows) for _ in range(world)]
    all_ids = [torch.empty_like(pad_ids) for _ in range(world)]
    dist.all_gather(all_rows, pad_rows)
    dist.all_gather(all_ids, pad_ids)
    rows, idv = torch.cat(all_rows), torch.cat(all_ids)
    keep = idv >= 0
    rows, idv = rows[keep], idv[keep]
    order = torch.argsort(idv)
    return rows[order], idv[order]
