import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ide3d_b200 import render
from oracle import renderer as orr, camera as ocam
sys.path.insert(0, 'tests')
from test_gpu_renderer import three_head_from_dense
dec = orr.Decoder.random(hidden=64, seed=1)
heads = three_head_from_dense(dec.w1, dec.b1, dec.w2, dec.b2)
n = 8
yaw = math.pi / 2 + np.linspace(-0.5, 0.5, n).reshape(n, 1).astype(np.float32)
cam = torch.from_numpy(ocam.look_at_pose(yaw, np.full((n, 1), math.pi / 2, np.float32), [0, 0, 0.2], radius=2.7, batch_size=n)).cuda()
flush = torch.empty(64 * 1024 * 1024, device='cuda')
for plane in (256,):
    for prec in ('tc',):
        tex = torch.randn(n, plane, plane, 96, device='cuda').permute(0, 3, 1, 2)
        seg = torch.randn(n, plane, plane, 96, device='cuda').permute(0, 3, 1, 2)
        pk = render.PackedDecoder(heads, 'cuda')
        for _ in range(3):
            render.raymarch(tex, seg, pk, cam, resolution=(64, 64), num_steps=96, jitter_seed=1, precision=prec, convert_layout=False)
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); render.raymarch(tex, seg, pk, cam, resolution=(64, 64), num_steps=96, jitter_seed=1, precision=prec, convert_layout=False); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f'plane {plane:4d} {prec:5s} {np.mean(ts):.3f} ms  ({8/np.mean(ts)*1e3:.0f} fps)', flush=True)
