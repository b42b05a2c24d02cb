#!/usr/bin/env python
"""Kernel-level timing of the fused ray-march at BASELINE configs[1] (8 frames, 64^2 rays x 96 samples, 256^2 planes as the backbone
delivers them): the round-2 kernel's variants (producer teams, gather instruction shape) and the round-1 kernel, same inputs, plus
the reference renderer's op chain on the same GPU (tests/test_gpu_speedup.py).  One JSON line per variant.

    IDE3D_BUILD_TUNING=1 python ide-3d_b200/build.py --force     # the variants are environment switches of a TUNING build
    python scripts/bench_raymarch.py [--only=<substring of the variant name>]
"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from ide3d_b200 import render
from ide3d_b200.torch_utils import custom_ops
custom_ops.verbosity = 'none'


def main():
    from bench import make_labels, make_latents, build_generator, NUM_STEPS, RENDER
    G = build_generator('cuda')
    with torch.no_grad():
        c = make_labels(8).cuda()
        ws = G.mapping(make_latents(8, G.z_dim).cuda(), c)
        vws, _ = G.synthesis.split_ws(ws)
        img_v, seg_v = G.synthesis.backbone(vws, noise_mode='const')
    cam = c[:, :16].reshape(-1, 4, 4)
    R = G.synthesis.renderer
    flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')

    def run():
        return R(img_v, seg_v, cam, img_size=RENDER, num_steps=NUM_STEPS, perturb='hash', seed=5)

    def timed(reps=20, warm=3):
        for _ in range(warm):
            run()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = run(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), float(np.min(ts)), out

    base = None
    variants = [('v1 (round 1)', dict(IDE3D_TC_V1='1'))]
    for coop, st, pf in (('1', '3', '1'), ('1', '3', '0'), ('1', '2', '1'), ('0', '2', '0')):
        variants.append((f'v3 coop={coop} stages={st} prefetch={pf}', dict(IDE3D_TC_V1='0', IDE3D_TC_V2='0', IDE3D_TC_TEAMS='2', IDE3D_TC_RAY_MAJOR='1', IDE3D_TC_COOP=coop,
                                                                            IDE3D_TC_STAGES=st, IDE3D_TC_PREFETCH=pf)))
    for teams in ('2',):
        for rm in ('1',):
            variants.append((f'v2 teams={teams} ray_major={rm}', dict(IDE3D_TC_V1='0', IDE3D_TC_V2='1', IDE3D_TC_TEAMS=teams, IDE3D_TC_RAY_MAJOR=rm)))
    only = [a.split('=', 1)[1] for a in sys.argv[1:] if a.startswith('--only=')]
    for name, env in variants:
        if only and not any(o in name for o in only):
            continue
        for k in ('IDE3D_TC_V1', 'IDE3D_TC_V2', 'IDE3D_TC_TEAMS', 'IDE3D_TC_RAY_MAJOR', 'IDE3D_TC_COOP', 'IDE3D_TC_STAGES', 'IDE3D_TC_PREFETCH'):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            med, mn, out = timed()
        except Exception as e:     # noqa
            print(json.dumps({'variant': name, 'error': str(e)[:300]}), flush=True)
            continue
        feat, depth, _ = out
        if base is None:
            base = (feat.clone(), depth.clone())
        rec = {'variant': name, 'ms_median': round(med, 4), 'ms_min': round(mn, 4), 'frames_per_s': round(8 / med * 1e3, 1),
               'max_abs_feat_diff_vs_v1': float((feat - base[0]).abs().max()), 'max_abs_depth_diff_vs_v1': float((depth - base[1]).abs().max()),
               'finite': bool(torch.isfinite(feat).all())}
        print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
