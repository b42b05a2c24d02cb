import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import renderer as orr
from ide3d_b200 import render
sys.path.insert(0, 'tests')
from test_gpu_renderer import _random_case, three_head_from_dense
for name, S, res, th, hid in [('3head', 48, (8, 8), True, 64), ('dense64', 40, (6, 6), False, 64), ('dense128', 33, (5, 7), False, 128), ('3head', 96, (16, 16), True, 64)]:
    tex, seg, dec, cam = _random_case(2, 32, seed=S, hidden=hid, three_head=th)
    heads = three_head_from_dense(dec.w1, dec.b1, dec.w2, dec.b2) if th else render.dense_heads(dec.w1, dec.b1, dec.w2, dec.b2)
    st = orr.render_frames(tex, seg, dec, cam, num_steps=S, resolution=res, return_stages=True)
    for prec in ('fp32', 'tc'):
        f, d, w = render.raymarch(tex.cuda(), seg.cuda(), heads, cam.cuda(), resolution=res, num_steps=S, return_weights=True, precision=prec)
        torch.cuda.synchronize()
        print(f'{name:9s} S={S} {prec:5s} feat err {(f.cpu()-st["rgb"]).abs().max().item():.3e} (mean {(f.cpu()-st["rgb"]).abs().mean().item():.2e}, ref max {st["rgb"].abs().max().item():.2f})  depth {(d.cpu()-st["depth"]).abs().max().item():.2e}  w {(w.cpu()-st["weights"]).abs().max().item():.2e}', flush=True)
