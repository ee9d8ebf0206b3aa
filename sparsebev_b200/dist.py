"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

How the path shards:
  * by SCENE (default; the reference's only strategy is DDP, train.py:90-131): every rank owns whole samples,
    no data-path collective, weak scaling.  `scene_partition` gives each rank its sample indices.
  * by FRAME (partitioning B): rank r owns frames [r*T/N, (r+1)*T/N) of every scene -- features never cross
    NVLink; each rank gathers its own frames for all queries and ONE all-gather per layer assembles the
    sampled features `[T, B, Q, G, P, C]` (frame-major staging so every rank's chunk is contiguous).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for world size 1)."""
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        kw = {'device_id': device} if (backend == 'nccl' and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device='cpu'):
    """Device-timed durations are reported as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def scene_partition(num_scenes, rank, world):
    """Contiguous block of scene indices for this rank (sizes differ by at most one)."""
    base, rem = divmod(num_scenes, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def frame_partition(num_frames, rank, world):
    """Frames [t0, t1) owned by this rank; num_frames must divide evenly so the all-gather chunks are equal."""
    if num_frames % world:
        raise ValueError('num_frames (%d) must be divisible by the world size (%d)' % (num_frames, world))
    per = num_frames // world
    return rank * per, (rank + 1) * per


def all_gather_frames(local, world=None):
    """local [T_local, ...] on every rank -> [T, ...] (frame-major, rank order == frame order)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = world or dist.get_world_size()
    out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out
