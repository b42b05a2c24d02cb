#!/usr/bin/env python
"""PTI-style step of the renderer at BASELINE configs[1] sizes (8 frames, 64^2 x 96, 256^2 planes): forward (fused kernel) + backward w.r.t.
both tri-planes and the decoder heads -- the backward kernel (ide3d_raymarch_bwd) vs the composed-chain backward it replaces
(render_grad.composed_chain + autograd, every stage materialised slab by slab).  One JSON line."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ide3d_b200 import render, render_grad
from ide3d_b200.torch_utils import custom_ops
custom_ops.verbosity = 'none'


def main():
    from bench import make_labels, make_latents, build_generator, NUM_STEPS, RENDER
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    G = build_generator('cuda')
    with torch.no_grad():
        c = make_labels(8).cuda()[:n]
        ws = G.mapping(make_latents(n, G.z_dim).cuda(), c)
        vws, _ = G.synthesis.split_ws(ws)
        img_v, seg_v = G.synthesis.backbone(vws, noise_mode='const')
    cam = c[:, :16].reshape(-1, 4, 4)
    R = G.synthesis.renderer
    gf = torch.randn(n, RENDER * RENDER, 51, device='cuda')
    gd = torch.randn(n, RENDER * RENDER, 1, device='cuda')

    def step():
        t, s = img_v.detach().clone().requires_grad_(True), seg_v.detach().clone().requires_grad_(True)
        heads = [tuple(h[:2]) + tuple(x.detach().clone().requires_grad_(True) for x in h[2:]) for h in R.heads()]
        feat, d, _ = render.raymarch(t, s, heads, cam, resolution=(RENDER, RENDER), num_steps=NUM_STEPS, jitter_seed=5, box_scale=R.box_scale)
        (feat * gf).sum().add((d * gd).sum()).backward()
        return t.grad, heads[0][2].grad

    def timed(reps, warm):
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        return float(np.median(ts)), out

    ms_k, (gk, wk) = timed(5, 2)
    render_grad.USE_BACKWARD_KERNEL = False
    ms_c, (gc, wc) = timed(2, 1)
    render_grad.USE_BACKWARD_KERNEL = True
    print(json.dumps({'config': f'renderer forward + backward (planes + decoder heads), {n} frames 64^2 x 96, 256^2 planes', 'kernel_path_ms': ms_k,
                      'composed_chain_path_ms': ms_c, 'speedup': ms_c / ms_k, 'max_abs_plane_grad_diff': float((gk - gc).abs().max()),
                      'plane_grad_scale': float(gc.abs().max()), 'max_abs_w1_grad_diff': float((wk - wc).abs().max()), 'w1_grad_scale': float(wc.abs().max())}))


if __name__ == '__main__':
    main()
