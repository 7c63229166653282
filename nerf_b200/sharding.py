"""Ray sharding across the GPUs of one node.

Rays are independent, so the path shards with no data-path collective: rank r renders a
contiguous slice of the flattened ray range (device RNG is keyed on the GLOBAL ray index, so the
image does not depend on the world size).  The one exchange is the final gather of the per-rank
(rays/N, C) outputs into the image — `gather_rows`, a single all_gather over NCCL/NVLink (gloo in
the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size, align=1):
    """Contiguous [start, start+count) of `n_items` for `rank`; every shard but the last is a multiple
    of `align` items (128-row tiles: 1 ray at 128 fine samples, 2 rays at 64 coarse samples)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    per = -(-n_items // world_size)
    per = -(-per // align) * align
    start = min(rank * per, n_items)
    return start, max(0, min(per, n_items - start))


def gather_rows(local, n_total, group=None):
    """all_gather of per-rank row blocks (count_r, C) laid out by shard_range -> (n_total, C) on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = shard_range(n_total, 0, world)[1]
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group) if local.is_cuda else dist.all_gather(
        list(out.view(world, per, *local.shape[1:]).unbind(0)), pad, group=group)
    return out[:n_total]


def render_image_sharded(render_rows, n_rays, channels=3, device=None, group=None):
    """Run `render_rows(start, count) -> (count, channels)` on this rank's shard and gather the image rows."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    start, count = shard_range(n_rays, rank, world)
    local = render_rows(start, count)
    if local.shape[0] != count:
        raise ValueError("render_rows returned the wrong number of rows")
    return gather_rows(local, n_rays, group)
