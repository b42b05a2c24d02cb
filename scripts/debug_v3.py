#!/usr/bin/env python
"""v2 / v3 ray-march kernels against the oracle on the smoke case and variations (errors printed, nothing asserted)."""
import json, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ide3d_b200 import render
from oracle import camera as ocam, renderer as orr

def case(n, plane, S, res, seed):
    g = torch.Generator().manual_seed(seed)
    up = lambda t: torch.nn.functional.interpolate(t, size=(plane, plane), mode='bicubic', align_corners=True)
    tex, seg = up(torch.randn(n, 96, 6, 6, generator=g)), up(torch.randn(n, 96, 6, 6, generator=g))
    dec = orr.Decoder.random(hidden=64, seed=1, three_head=True)
    H = 64
    heads = [(0, 0, dec.w1[0:H, 0:32], dec.b1[0:H], dec.w2[0:32, 0:H], dec.b2[0:32]),
             (1, 32, dec.w1[H:2 * H, 32:], dec.b1[H:2 * H], dec.w2[32:51, H:2 * H], dec.b2[32:51]),
             (1, 51, dec.w1[2 * H:, 32:], dec.b1[2 * H:], dec.w2[51:52, 2 * H:], dec.b2[51:52])]
    yaw = np.linspace(-0.3, 0.3, n).reshape(n, 1).astype(np.float32) + math.pi / 2
    cam = torch.from_numpy(ocam.look_at_pose(yaw, np.full((n, 1), math.pi / 2, np.float32), [0, 0, 0.2], radius=2.7, batch_size=n))
    ref = orr.render_frames(tex, seg, dec, cam, num_steps=S, resolution=res, jitter_seed=42)
    return tex, seg, heads, cam, ref

only = sys.argv[1:] 
for (n, plane, S, res) in [(2, 32, 24, (16, 16)), (2, 32, 96, (16, 16)), (2, 128, 24, (16, 16)), (1, 32, 8, (8, 8))]:
    tex, seg, heads, cam, (rf, rd, rw) = case(n, plane, S, res, 0)
    for name, env in [('v2', dict(IDE3D_TC_V2='1', IDE3D_TC_TEAMS='2')), ('v3 t2', dict(IDE3D_TC_V2='0', IDE3D_TC_TEAMS='2')), ('v3 t3', dict(IDE3D_TC_V2='0', IDE3D_TC_TEAMS='3')), ('fp32', None)]:
        if only and not any(o in name for o in only):
            continue
        for k in ('IDE3D_TC_V2', 'IDE3D_TC_TEAMS'):
            os.environ.pop(k, None)
        if env:
            os.environ.update(env)
        f, d, w = render.raymarch(tex.cuda(), seg.cuda(), heads, cam.cuda(), resolution=res, num_steps=S, jitter_seed=42, return_weights=True,
                                  precision='fp32' if env is None else 'tc')
        torch.cuda.synchronize()
        print(json.dumps({'case': [n, plane, S, list(res)], 'kernel': name, 'feat_err': float((f.cpu() - rf).abs().max()), 'depth_err': float((d.cpu() - rd).abs().max()),
                          'w_err': float((w.cpu() - rw).abs().max()), 'feat_max': float(rf.abs().max())}), flush=True)
