#!/usr/bin/env python
"""The smoke case of __graft_entry__ (2 frames, 16x16 rays x 24 samples, 32^2 planes, hash jitter) as the FIRST launch of a fresh process,
with / without the weights output; prints the errors against the oracle.  usage: debug_smoke.py <return_weights 0|1> [repeat]"""
import json, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ide3d_b200 import render
from oracle import camera as ocam, renderer as orr
rw = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator().manual_seed(0)
n, plane, S, res = 2, 32, 24, (16, 16)
up = lambda t: torch.nn.functional.interpolate(t, size=(plane, plane), mode='bicubic', align_corners=True)
tex, seg = up(torch.randn(n, 96, 6, 6, generator=g)), up(torch.randn(n, 96, 6, 6, generator=g))
dec = orr.Decoder.random(hidden=64, seed=1, three_head=True)
H = 64
heads = [(0, 0, dec.w1[0:H, 0:32], dec.b1[0:H], dec.w2[0:32, 0:H], dec.b2[0:32]),
         (1, 32, dec.w1[H:2 * H, 32:], dec.b1[H:2 * H], dec.w2[32:51, H:2 * H], dec.b2[32:51]),
         (1, 51, dec.w1[2 * H:, 32:], dec.b1[2 * H:], dec.w2[51:52, 2 * H:], dec.b2[51:52])]
yaw = np.array([[math.pi / 2 - 0.3], [math.pi / 2 + 0.3]], np.float32)
cam = torch.from_numpy(ocam.look_at_pose(yaw, np.full((n, 1), math.pi / 2, np.float32), [0, 0, 0.2], radius=2.7, batch_size=n))
ref_f, ref_d, _ = orr.render_frames(tex, seg, dec, cam, num_steps=S, resolution=res, jitter_seed=42)
td, sd, cd = tex.cuda(), seg.cuda(), cam.cuda()
out = []
for i in range(rep):
    feat, depth, _ = render.raymarch(td, sd, heads, cd, resolution=res, num_steps=S, jitter_seed=42, return_weights=rw)
    torch.cuda.synchronize()
    ef = (feat.cpu() - ref_f).abs()
    out.append((float(ef.max()), float((depth.cpu() - ref_d).abs().max()), int((ef > 1e-4).sum()), int((ef.amax(-1) > 1e-4).sum())))
print(json.dumps({'return_weights': rw, 'lib': os.environ.get('IDE3D_B200_LIB', 'product'), 'env': {k: v for k, v in os.environ.items() if k.startswith('IDE3D_TC')},
                  'runs (feat err, depth err, #values > 1e-4, #rays > 1e-4)': out}))
if os.environ.get('IDE3D_DEBUG_DUMP'):
    f0, d0, _ = render.raymarch(td, sd, heads, cd, resolution=res, num_steps=S, jitter_seed=42, return_weights=False)
    f1, d1, w1 = render.raymarch(td, sd, heads, cd, resolution=res, num_steps=S, jitter_seed=42, return_weights=True)
    f2, d2, w2 = render.raymarch(td, sd, heads, cd, resolution=res, num_steps=S, jitter_seed=42, return_weights=True, precision='fp32')
    torch.cuda.synchronize()
    e0 = (f0.cpu() - ref_f).abs().amax(-1)            # [n, R]
    bad = (e0 > 1e-4).nonzero().tolist()
    print('bad rays (n, ray -> px, py, unit):', [(n_, r_, r_ % 16, r_ // 16, (r_ // 16 // 4) * 4 + (r_ % 16) // 4) for n_, r_ in bad])
    print('rw0 vs rw1 max diff', float((f0 - f1).abs().max()), 'depth', float((d0 - d1).abs().max()))
    print('rw1 vs fp32 max diff', float((f1 - f2).abs().max()), 'weights', float((w1 - w2).abs().max()))
    n_, r_ = bad[0] if bad else (0, 0)
    print('ray', n_, r_, 'feat rw0', f0[n_, r_, :6].tolist(), 'rw1', f1[n_, r_, :6].tolist(), 'ref', ref_f[n_, r_, :6].tolist())
    print('depth rw0', float(d0[n_, r_]), 'rw1', float(d1[n_, r_]), 'ref', float(ref_d[n_, r_]))
    print('weights rw1 sum', float(w1[n_, r_].sum()), 'last 4', w1[n_, r_, -4:, 0].tolist())
