#!/usr/bin/env python
"""BASELINE config 4: extract_shapes' 256^3 density grid (cube_size 1), sigma-only query.  Times (a) ide3d_sigma_grid
(points generated in-kernel, one launch), (b) the reference's own loop shape: explicit points in chunks of 100000 through
sample_voxel (extract_shapes.py:144-148).  Prints JSON lines."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ide3d_b200.torch_utils import custom_ops
custom_ops.verbosity = 'none'
from ide3d_b200.compat import random_init_generator
from oracle import renderer as orr

G = random_init_generator('cuda', seed=0)
N = 256
with torch.no_grad():
    z = torch.from_numpy(np.random.RandomState(0).randn(1, G.z_dim)).float().cuda()
    c = torch.tensor([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 2.7, 0, 0, 0, 1, 4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1.]).cuda()[None]
    ws = G.mapping(z, c)
    vws, _ = G.synthesis.split_ws(ws)
    img_v, seg_v = G.synthesis.backbone(vws, noise_mode='const')
    R = G.synthesis.renderer
    tex, seg = R.as_planes(img_v), R.as_planes(seg_v)
    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        return float(np.median(ts)), out
    ms_grid, sg = timed(lambda: R.sigma_grid(tex, seg, grid_n=N, cube_length=1.0))
    samples, _, _ = orr.create_samples(N, [0, 0, 0], 1.0)
    samples = (0.9 * samples).cuda()
    def ref_loop():
        sig = torch.zeros((1, N ** 3, 1), device='cuda')
        head = 0
        while head < N ** 3:
            out = R.sample_voxel(tex, seg, samples[:, head:head + 100000]).reshape(1, -1, 52)
            sig[:, head:head + 100000] = out[:, :, -1:]
            head += 100000
        return sig
    ms_loop, sl = timed(ref_loop, reps=2)
    err = (sg.reshape(-1) - sl.reshape(-1)).abs().max().item()
    npts = N ** 3
    alg = 2 * 96 * 256 * 256 * 4 + npts * 4
    print(json.dumps({'config': 'extract_shapes 256^3 sigma grid', 'sigma_grid_ms': round(ms_grid, 3), 'Mpoints_per_s': round(npts / ms_grid / 1e3, 1),
                      'algorithmic_GBs': round(alg / ms_grid / 1e6, 1), 'reference_loop_shape_ms (168 x sample_voxel[100000,52])': round(ms_loop, 3),
                      'max_abs_diff_between_the_two': err}))
