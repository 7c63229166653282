"""Ray sharding across the GPUs of one node (BASELINE configs[4]: one image, ray tiles over 8 GPUs).

Rays are independent, so the path shards with no data-path collective: rank r renders a contiguous slice of the
flattened ray range (device RNG is keyed on the GLOBAL ray index, so the image does not depend on the world size).
The one exchange is the final gather of the per-rank rgb rows into the image.  Two implementations:

* `PeerImage` (default on NVLink boxes): every rank owns a full-size image buffer allocated by libnerfb200 and exported
  through CUDA IPC; every rank maps all the others.  The compositing epilogue of the fine kernel then stores each
  finished rgb row into ALL the images directly (nb2_render_params.peer_rgb): the gather rides on the kernel's own
  stores over NVLink / NVSwitch, tile by tile, and no gather kernel or collective is launched.  `fence()` (one tiny
  all_reduce) orders "every rank's kernel has finished" before anyone reads the image.
* `gather_rows`: one NCCL all_gather_into_tensor into a pre-allocated buffer (gloo all_gather in the CPU tests) — the
  baseline the fused path is compared with in bench.py.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def shard_range(n_items, rank, world_size, align=1):
    """Contiguous [start, start+count) of `n_items` for `rank`; every shard but the last is a multiple
    of `align` items (128-row tiles: 1 ray at 128 fine samples, 2 rays at 64 coarse samples)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    per = -(-n_items // world_size)
    per = -(-per // align) * align
    start = min(rank * per, n_items)
    return start, max(0, min(per, n_items - start))


class GatherBuffers:
    """Pre-allocated staging for gather_rows (nothing is allocated inside a timed step)."""

    def __init__(self, n_total, channels, world, device, dtype=torch.float32):
        self.per = shard_range(n_total, 0, world)[1]
        self.pad = torch.zeros((self.per, channels), dtype=dtype, device=device)
        self.out = torch.empty((world * self.per, channels), dtype=dtype, device=device)


def gather_rows(local, n_total, group=None, buffers=None):
    """all_gather of per-rank row blocks (count_r, C) laid out by shard_range -> (n_total, C) on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if buffers is None:
        buffers = GatherBuffers(n_total, local.shape[1], world, local.device, local.dtype)
    pad, out = buffers.pad, buffers.out
    if local.shape[0] == buffers.per:
        pad = local                                  # full shard: gather straight from the render output
    else:
        pad[:local.shape[0]] = local
    if local.is_cuda:
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        dist.all_gather(list(out.view(world, buffers.per, *local.shape[1:]).unbind(0)), pad, group=group)
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


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can wrap library-owned device memory without copying."""

    def __init__(self, ptr, shape, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}
        self._owner = owner


class PeerImage:
    """Full-size (n_rows, channels) fp32 image on every rank, each mapped into every other rank (CUDA IPC)."""

    def __init__(self, n_rows, channels, device, group=None):
        if not dist.is_initialized():
            raise _lib.NB2Error("PeerImage needs an initialised process group (one process per GPU)")
        self.group, self.device = group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise _lib.NB2Error("PeerImage: at most 8 GPUs (one NVSwitch node)")
        self.n_rows, self.channels = n_rows, channels
        lib, h = _lib.load(), _lib.handle(self.device)
        self._lib, self._h = lib, h
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(lib.nb2_ipc_alloc(h, n_rows * channels * 4, ctypes.byref(own), handle))
        self._own = own
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._mapped = {}
        self.peer_ptrs = []
        for r, hb in enumerate(handles):
            if r == self.rank:
                continue
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(hb)
            _lib.check(lib.nb2_ipc_open(h, buf, ctypes.byref(p)))
            self._mapped[r] = p
            self.peer_ptrs.append(p.value)
        self.image = torch.as_tensor(_DevArray(own.value, (n_rows, channels), self), device=self.device)
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        dist.barrier(group=group)

    def local_rows(self, start, count):
        """This rank's slice of its own image: the render writes its rows here and into every peer's image."""
        return self.image[start:start + count]

    def fence(self):
        """Every rank's render kernels (and their peer stores) have completed before anyone reads `image`."""
        dist.all_reduce(self._token, group=self.group)

    def close(self):
        if self._lib is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for p in self._mapped.values():
            self._lib.nb2_ipc_close(self._h, p)
        dist.barrier(group=self.group)
        self.image = None
        self._lib.nb2_ipc_free(self._h, self._own)
        self._lib = None
