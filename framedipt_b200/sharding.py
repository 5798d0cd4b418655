"""Sample-parallel plumbing: independent chains shard across ranks, one process per GPU.

No collective sits inside the timestep loop.  ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the
CPU tests) is used for exactly two things: broadcasting the packed weight blob from rank 0 before sampling and
gathering the finished samples' coordinates afterwards.  The reference has no equivalent (it replicates whole
processes through Hydra's joblib launcher, config/base.yaml:3-4, experiments/utils.py:64-76).
"""
from __future__ import annotations

import torch


def shard_range(n_samples: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [start, end) slice of the sample axis owned by `rank`; remainders go to the lowest ranks."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_samples, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_state_dict(sd: dict[str, torch.Tensor], dist, src: int = 0, device: torch.device | str = "cpu") -> dict[str, torch.Tensor]:
    """One broadcast of all parameters packed into a single fp32 blob (keys sorted; shapes must agree on every rank)."""
    keys = sorted(sd)
    blob = torch.cat([sd[k].detach().reshape(-1).to(torch.float32) for k in keys]).to(device)
    if dist.get_rank() != src:
        blob.zero_()
    dist.broadcast(blob, src)
    out, off = {}, 0
    for k in keys:
        n = sd[k].numel()
        out[k] = blob[off:off + n].reshape(sd[k].shape).cpu()
        off += n
    return out


def gather_samples(local: torch.Tensor, dist, dst: int = 0) -> torch.Tensor | None:
    """Gathers per-rank [B_local, ...] results along the sample axis on `dst` (ragged B_local allowed)."""
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
    bmax = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((bmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if dist.get_rank() == dst else None
    dist.gather(pad, bufs, dst=dst)
    if bufs is None:
        return None
    return torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)], 0)
