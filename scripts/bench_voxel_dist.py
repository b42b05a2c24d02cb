#!/usr/bin/env python
"""BASELINE config 4 at N GPUs: extract_shapes.py 256^3 density grid (cube_size 1), flat voxel range split into contiguous z-slabs over
the ranks (dist.sigma_grid_sharded), one all_gather of the slabs.  One process per GPU (plain python for N = 1, torchrun for N > 1).
Rank 0 prints one JSON line: kernel-only time of the local slab, and the whole call including the gather (CUDA events, max over ranks)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from ide3d_b200 import dist as idist
    from ide3d_b200.torch_utils import custom_ops
    from ide3d_b200.compat import random_init_generator
    import torch.distributed as tdist
    custom_ops.verbosity = 'none'
    rank, world, device = idist.init_from_env()
    G = random_init_generator(device=device, seed=0)
    N = 256
    with torch.no_grad():
        z = torch.from_numpy(np.random.RandomState(0).randn(1, G.z_dim)).float().to(device)
        c = torch.tensor([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 2.7, 0, 0, 0, 1, 4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1.], device=device)[None]
        ws = G.mapping(z, c)
        vws, _ = G.synthesis.split_ws(ws)
        img_v, seg_v = G.synthesis.backbone(vws, noise_mode='const')
        R = G.synthesis.renderer
        tex, seg = R.as_planes(img_v), R.as_planes(seg_v)
        first, count = idist.slab_range(N ** 3, rank, world)

        def timed(fn, reps=10):
            fn()
            if world > 1:
                tdist.barrier()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            return float(np.median(ts)), out

        ms_local, _ = timed(lambda: R.sigma_grid(tex, seg, grid_n=N, cube_length=1.0, first=first, count=count))
        ms_all, vol = timed(lambda: idist.sigma_grid_sharded(G, tex, seg, rank, world, grid_n=N, cube_length=1.0))
        t = torch.tensor([ms_local, ms_all], dtype=torch.float64, device=device)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        if rank == 0:
            npts = N ** 3
            print(json.dumps({'config': 'extract_shapes 256^3 sigma grid (BASELINE configs[3])', 'n_gpus': world, 'slab_kernel_ms': float(t[0]),
                              'sharded_call_ms_incl_all_gather': float(t[1]), 'Mpoints_per_s_kernel': npts / float(t[0]) / 1e3,
                              'Mpoints_per_s_call': npts / float(t[1]) / 1e3, 'volume_shape': list(vol.shape), 'finite': bool(torch.isfinite(vol).all()),
                              'algorithmic_bytes': 96 * 256 * 256 * 4 + npts * 4}))
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == '__main__':
    main()
