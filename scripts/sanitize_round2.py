#!/usr/bin/env python
"""One small invocation of every kernel added in round 2, for compute-sanitizer (memcheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_round2.py"""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ide3d_b200 import mesh, render
from ide3d_b200.compat import random_init_generator
from ide3d_b200.torch_utils import custom_ops
from ide3d_b200.torch_utils.ops import filtered_lrelu, upfirdn2d
custom_ops.verbosity = 'none'
dev = 'cuda'
G = random_init_generator(device=dev, seed=0, img_resolution=128, plane_resolution=64, render_size=16, channel_max=64)
c = torch.eye(4, device=dev).reshape(1, 16).repeat(2, 1); c[:, 11] = 2.7
c = torch.cat([c, torch.zeros(2, 9, device=dev)], 1)
with torch.no_grad():
    ws = G.mapping(torch.randn(2, G.z_dim, device=dev), c)
    img = G.synthesis(ws, c=c, render_params=dict(num_steps=20))                                   # style plan, v3 ray-march, FIR / epilogues, skip-add
    img_h = G.synthesis(ws, c=c, render_params=dict(num_steps=12, hierarchical=True, n_importance=8))   # ZVALS mode + sample_pdf
    vws, _ = G.synthesis.split_ws(ws)
    img_v, seg_v = G.synthesis.backbone(vws, noise_mode='const')
    R = G.synthesis.renderer
    sig = R.sigma_grid(img_v[:1], seg_v[:1], grid_n=40, cube_length=1.0)                         # sigma_tc_kernel (grid)
    pts = torch.rand(2, 333, 3, device=dev) - 0.5
    sv = R.sample_voxel(img_v, seg_v, pts, sigma_only=True)                                        # sigma_tc_kernel (points, ragged)
    v, t = mesh.marching_cubes(sig.reshape(40, 40, 40), float(sig.median()))                       # mc_classify / mc_emit
    x = torch.randn(1, 8, 50, 37, device=dev)
    f = upfirdn2d.setup_filter(np.hanning(14)[1:-1].tolist(), device=dev)
    y = filtered_lrelu.filtered_lrelu(x, fu=f, fd=f, b=torch.randn(8, device=dev), up=2, down=2, padding=[10, 11, 10, 11], clamp=1.5)            # thread-staged tile
    y2 = filtered_lrelu.filtered_lrelu(torch.randn(1, 8, 64, 64, device=dev).contiguous(memory_format=torch.channels_last), fu=f, fd=f,
                                       b=torch.randn(8, device=dev), up=2, down=2, padding=[10, 11, 10, 11])                                  # TMA tile, channels-last
# backward kernel
t_, s_ = img_v.detach().clone().requires_grad_(True), seg_v.detach().clone().requires_grad_(True)
heads = [tuple(h[:2]) + tuple(p.detach().clone().requires_grad_(True) for p in h[2:]) for h in R.heads()]
feat, d, _ = render.raymarch(t_, s_, heads, c[:, :16].reshape(-1, 4, 4), resolution=(10, 9), num_steps=13, jitter_seed=3, box_scale=R.box_scale)
(feat.square().sum() + d.sum()).backward()
torch.cuda.synchronize()
print('ok', img.shape, img_h.shape, sig.shape, sv.shape, v.shape, t.shape, y.shape, y2.shape, float(t_.grad.abs().max()))
