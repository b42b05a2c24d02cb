"""One process per GPU: shard independent work units, gather the results once.

The renderer has no exchange step (every frame is a pure function of its own (w, camera), every voxel of its own
coordinates), so there is no data-path collective: ranks take a strided share of the units -- the reference's own
idiom for metric evaluation, metrics/metric_utils.py:243 -- and a single all_gather (NCCL over NVLink on GPUs, gloo
in the CPU tests) brings the uint8 frames / sigma slabs together.  0.79 MB per 512^2 RGB frame: bandwidth-trivial.
"""

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world, device)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    use_cuda = torch.cuda.is_available()
    device = torch.device('cuda', local) if use_cuda else torch.device('cpu')
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        kw = dict(device_id=device) if use_cuda else {}
        dist.init_process_group(backend or ('nccl' if use_cuda else 'gloo'), rank=rank, world_size=world, **kw)
    return rank, world, device


def shard_indices(num_items, rank, world):
    """Strided share of range(num_items) for this rank (rank, rank+world, ...)."""
    return list(range(rank, num_items, world))


def slab_range(total, rank, world):
    """Contiguous [first, first+count) share of a flat voxel range (z-slab shard of the sigma grid)."""
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def all_gather_padded(local, num_items, world):
    """Gather per-rank tensors produced for shard_indices(...) back into item order.
    local: [len(shard), ...] on this rank.  Returns [num_items, ...] on every rank."""
    if world == 1:
        return local
    per = (num_items + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous())
    out = out.reshape((world, per) + tuple(local.shape[1:]))
    # item i lives at rank i % world, slot i // world
    idx = torch.arange(num_items, device=local.device)
    return out[idx % world, idx // world]


def all_gather_slabs(local, total, world):
    """Gather contiguous slabs (slab_range) into the flat [.., total] tensor (last dim is the sharded one)."""
    if world == 1:
        return local
    base, rem = divmod(total, world)
    per = base + (1 if rem else 0)
    pad = torch.zeros(tuple(local.shape[:-1]) + (per,), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-1]] = local
    out = torch.empty((world * pad.shape[0],) + tuple(pad.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous())          # concatenation along dim 0 (gloo and nccl agree on this form)
    out = out.reshape((world,) + tuple(pad.shape))
    parts = [out[r][..., :base + (1 if r < rem else 0)] for r in range(world)]
    return torch.cat(parts, dim=-1)


@torch.no_grad()
def render_frames_sharded(G, ws, c, rank, world, batch=8, to_uint8=True, **synthesis_kwargs):
    """Render frames i = rank, rank+world, ... of (ws[i], c[i]) in batches and all-gather them in frame order.
    ws [F, num_ws, w_dim], c [F, 25] (host or device).  Returns uint8 [F, 3, H, W] (or float) on every rank."""
    F = ws.shape[0]
    mine = shard_indices(F, rank, world)
    dev = next(G.parameters()).device
    outs = []
    for i in range(0, len(mine), batch):
        sel = torch.as_tensor(mine[i:i + batch])
        img = G.synthesis(ws[sel].to(dev, non_blocking=True), c=c[sel].to(dev, non_blocking=True), **synthesis_kwargs)
        if isinstance(img, (tuple, list)):
            img = img[0]
        if to_uint8:
            img = (img * 127.5 + 128).clamp(0, 255).to(torch.uint8)
        outs.append(img)
    local = torch.cat(outs) if outs else torch.empty((0, G.img_channels, G.img_resolution, G.img_resolution),
                                                      dtype=torch.uint8 if to_uint8 else torch.float32, device=dev)
    return all_gather_padded(local, F, world)


@torch.no_grad()
def sigma_grid_sharded(G, img_v, seg_v, rank, world, grid_n=256, cube_length=1.0, voxel_origin=(0, 0, 0)):
    """extract_shapes' density grid, flat voxel range split into contiguous z-slabs over the ranks, one all_gather."""
    total = grid_n ** 3
    first, count = slab_range(total, rank, world)
    local = G.synthesis.renderer.sigma_grid(img_v, seg_v, grid_n=grid_n, voxel_origin=voxel_origin,
                                            cube_length=cube_length, first=first, count=count)
    return all_gather_slabs(local, total, world)
