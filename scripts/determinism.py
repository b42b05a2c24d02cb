#!/usr/bin/env python
"""Root-cause of the 1-uint8-level run-to-run difference seen in round 1.

Renders the same (ws, c) batch repeatedly in one process and compares the results BITWISE, stage by stage:
  raymarch : the fused kernel alone, identical planes, 6 launches            -> must be bit-equal (it has no atomics)
  planes   : backbone (cuDNN convolutions + FIR / epilogue kernels), 4 runs  -> bit-equal unless cuDNN picks atomics / split-K
  image    : whole synthesis, 4 runs
  streamed : dist.stream_frames_sharded halves (identical inputs as two batches, copy stream overlap)
each under torch.backends.cudnn.{benchmark, deterministic} in {F,T}^2.  Prints one JSON report."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ide3d_b200 import dist as idist                                   # noqa: E402
from ide3d_b200.compat import random_init_generator                    # noqa: E402
from ide3d_b200.torch_utils import custom_ops                          # noqa: E402

custom_ops.verbosity = 'none'


def main():
    small = '--small' in sys.argv
    from bench import make_labels, make_latents
    kwargs = dict(z_dim=32, w_dim=32, img_resolution=128, plane_resolution=64, render_size=32, channel_base=2048, channel_max=64,
                  sr_channels=(32, 32), mapping_kwargs=dict(num_layers=2)) if small else {}
    G = random_init_generator('cuda', seed=0, **kwargs)
    n = 4
    z, c = make_latents(n, G.z_dim).cuda(), make_labels(8)[:n].cuda()
    S = 24 if small else 96
    report = {}
    with torch.no_grad():
        ws = G.mapping(z, c)
        vws, _ = G.synthesis.split_ws(ws)
        for bench in (False, True):
            for det in (False, True):
                torch.backends.cudnn.benchmark, torch.backends.cudnn.deterministic = bench, det
                key = f'benchmark={bench},deterministic={det}'
                planes = [G.synthesis.backbone(vws, noise_mode='const') for _ in range(4)]
                cam = c[:, :16].reshape(-1, 4, 4)
                rm = [G.synthesis.renderer(planes[0][0], planes[0][1], cam, img_size=G.synthesis.render_size, num_steps=S, perturb='hash', seed=3)
                      for _ in range(6)]
                imgs = [G.synthesis(ws, c=c, render_params=dict(num_steps=S), noise_mode='const', perturb='hash', seed=3) for _ in range(4)]
                ws2, c2 = ws.cpu().repeat(2, 1, 1).pin_memory(), c.cpu().repeat(2, 1).pin_memory()
                st = idist.stream_frames_sharded(G, ws2, c2, 0, 1, batch=n, render_params=dict(num_steps=S), noise_mode='const', perturb='hash', seed=3).clone()
                torch.cuda.synchronize()
                report[key] = {
                    'raymarch_bit_equal': all(torch.equal(rm[0][0], r[0]) and torch.equal(rm[0][1], r[1]) for r in rm[1:]),
                    'planes_bit_equal': all(torch.equal(planes[0][0], p[0]) and torch.equal(planes[0][1], p[1]) for p in planes[1:]),
                    'planes_max_abs_diff': max((planes[0][0] - p[0]).abs().max().item() for p in planes[1:]),
                    'image_bit_equal': all(torch.equal(imgs[0], i) for i in imgs[1:]),
                    'image_max_abs_diff': max((imgs[0] - i).abs().max().item() for i in imgs[1:]),
                    'streamed_halves_bit_equal': bool(torch.equal(st[:n], st[n:])),
                    'streamed_halves_max_levels': int((st[:n].int() - st[n:].int()).abs().max()),
                }
    print(json.dumps({'small': small, 'report': report}, indent=1))


if __name__ == '__main__':
    main()
