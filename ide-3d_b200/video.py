"""Batched driver for the latent-interpolation video of gen_videos.py (BASELINE config 3).

The reference renders one (grid cell, frame) per `G.synthesis` call and evaluates a scipy spline and a camera pose on the host
in between (gen_videos.py:118-129), i.e. batch 1 and a host round trip per frame.  Every (w, camera) pair of the video is a pure
function of (seeds, frame index), so here they are all computed up front -- `interp_video_inputs`, same arithmetic, same nested
order (frame, grid row, grid column) -- and the frames are rendered by `dist.stream_frames_sharded` in batches, sharded over the
ranks, with the host download overlapping the next batch.  Video encoding (imageio / ffmpeg) is not part of this package; the
result is the uint8 frame grid `layout_grid` (gen_videos.py:24-38) would hand to the writer.
"""

import math

import numpy as np
import torch

INTRINSICS = [4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1]           # gen_videos.py:85


def interp_video_inputs(G, seeds, shuffle_seed=None, w_frames=60 * 4, kind='cubic', grid_dims=(1, 1), num_keyframes=None, wraps=2,
                        psi=1, truncation_cutoff=14, cfg='FFHQ', device=None):
    """-> ws [F * grid_h * grid_w, num_ws, w_dim] (float64, as scipy returns them), c [F * grid_h * grid_w, 25] float32, and
    (F, grid_h, grid_w); row index = (frame * grid_h + yi) * grid_w + xi, the order of gen_videos.py:112-129."""
    import scipy.interpolate
    from .training.volumetric_rendering import LookAtPoseSampler
    grid_w, grid_h = grid_dims
    if num_keyframes is None:
        if len(seeds) % (grid_w * grid_h) != 0:
            raise ValueError('Number of input seeds must be divisible by grid W*H')
        num_keyframes = len(seeds) // (grid_w * grid_h)
    all_seeds = np.zeros(num_keyframes * grid_h * grid_w, dtype=np.int64)
    for idx in range(num_keyframes * grid_h * grid_w):
        all_seeds[idx] = seeds[idx % len(seeds)]
    if shuffle_seed is not None:
        np.random.RandomState(seed=shuffle_seed).shuffle(all_seeds)
    device = device if device is not None else next(G.parameters()).device
    lookat = torch.tensor([0, 0, 0.2] if cfg == 'FFHQ' else [0, 0, 0], dtype=torch.float32, device=device)
    intrinsics = torch.tensor(INTRINSICS, dtype=torch.float32, device=device).reshape(1, 9)

    # keyframe latents (:80-88)
    zs = torch.from_numpy(np.stack([np.random.RandomState(seed).randn(G.z_dim) for seed in all_seeds])).to(device)
    pose0 = LookAtPoseSampler.sample(math.pi / 2, math.pi / 2, lookat, radius=2.7, device=device)
    c0 = torch.cat([pose0.reshape(-1, 16), intrinsics], 1).repeat(len(zs), 1)
    ws = G.mapping(z=zs, c=c0, truncation_psi=psi, truncation_cutoff=truncation_cutoff)
    ws = ws.reshape(grid_h, grid_w, num_keyframes, *ws.shape[1:])

    # all frames of every cell's spline at once (:93-101, :127-128)
    F = num_keyframes * w_frames
    t = np.arange(F) / w_frames
    x = np.arange(-num_keyframes * wraps, num_keyframes * (wraps + 1))
    w_all = np.empty((F, grid_h, grid_w) + tuple(ws.shape[3:]), dtype=np.float64)
    for yi in range(grid_h):
        for xi in range(grid_w):
            y = np.tile(ws[yi][xi].cpu().numpy(), [wraps * 2 + 1, 1, 1])
            w_all[:, yi, xi] = scipy.interpolate.interp1d(x, y, kind=kind, axis=0)(t)

    # camera sweep (:118-124): the pose depends on the frame only, every cell of a frame shares it
    fi = np.arange(F)
    yaw = math.pi / 2 - 0.5 * np.sin(2 * math.pi * fi / F)
    pitch = math.pi / 2 - 0.05 + 0.25 * np.cos(2 * math.pi * fi / F)
    h = torch.from_numpy(yaw).to(torch.float32).reshape(F, 1).to(device)
    v = torch.from_numpy(pitch).to(torch.float32).reshape(F, 1).to(device)
    poses = LookAtPoseSampler.sample(h, v, lookat, radius=2.7, batch_size=F, device=device)
    c = torch.cat([poses.reshape(F, 16), intrinsics.expand(F, 9)], 1)
    c = c.reshape(F, 1, 1, 25).expand(F, grid_h, grid_w, 25)
    return (torch.from_numpy(w_all).reshape(F * grid_h * grid_w, *ws.shape[3:]), c.reshape(F * grid_h * grid_w, 25).contiguous(),
            (F, grid_h, grid_w))


def layout_frames(frames, F, grid_h, grid_w):
    """uint8 [F*grid_h*grid_w, C, H, W] (row order of interp_video_inputs) -> [F, grid_h*H, grid_w*W, C]: layout_grid
    (gen_videos.py:24-38) applied to every frame."""
    n, ch, ih, iw = frames.shape
    assert n == F * grid_h * grid_w
    g = frames.reshape(F, grid_h, grid_w, ch, ih, iw).permute(0, 1, 4, 2, 5, 3)
    return g.reshape(F, grid_h * ih, grid_w * iw, ch)


@torch.no_grad()
def render_interp_video(G, seeds, rank=0, world=1, batch=8, out=None, synthesis_kwargs=None, **kwargs):
    """All frames of gen_interp_video (image_mode='image') as uint8 grids [F, grid_h*H, grid_w*W, 3] on the host of rank 0
    (None on the other ranks).  kwargs: the arguments of `interp_video_inputs`; synthesis_kwargs: extra arguments of
    G.synthesis (default noise_mode='const' like gen_videos.py:129; the depth jitter is drawn per batch as in the reference).
    F * grid cells must be a multiple of world * batch (pad the seed list or pick the batch accordingly)."""
    from . import dist as idist
    ws, c, (F, gh, gw) = interp_video_inputs(G, seeds, **kwargs)
    pin = torch.cuda.is_available()
    ws32 = ws.to(torch.float32)
    c32 = c.cpu()
    if pin:
        ws32, c32 = ws32.pin_memory(), c32.pin_memory()
    skw = dict(noise_mode='const')
    skw.update(synthesis_kwargs or {})
    frames = idist.stream_frames_sharded(G, ws32, c32, rank, world, batch=batch, out=out, **skw)
    return None if frames is None else layout_frames(frames, F, gh, gw)
