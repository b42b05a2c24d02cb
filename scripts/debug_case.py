import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import renderer as orr, camera as ocam
from ide3d_b200 import render
sys.path.insert(0, 'tests')
from test_gpu_renderer import _random_case, three_head_from_dense
DEV='cuda'
g = torch.Generator(device='cuda').manual_seed(0)
n = 8
up = lambda t: torch.nn.functional.interpolate(t, size=(256, 256), mode='bicubic', align_corners=True).contiguous(memory_format=torch.channels_last)
tex = up(torch.randn(n, 96, 8, 8, device=DEV, generator=g)); seg = up(torch.randn(n, 96, 8, 8, device=DEV, generator=g))
dec = orr.Decoder.random(hidden=64, seed=1)
heads = three_head_from_dense(dec.w1, dec.b1, dec.w2, dec.b2)
yaw = math.pi / 2 + np.linspace(-0.5, 0.5, n).reshape(n, 1).astype(np.float32)
cam = torch.from_numpy(ocam.look_at_pose(yaw, np.full((n, 1), math.pi / 2, np.float32), [0, 0, 0.2], radius=2.7, batch_size=n)).to(DEV)
sub = slice(5, 6)
st = orr.render_frames(tex[sub].cpu().contiguous(), seg[sub].cpu().contiguous(), dec, cam[sub].cpu(), num_steps=96, resolution=(64, 64), return_stages=True)
f, d, w = render.raymarch(tex[sub], seg[sub], heads, cam[sub], resolution=(64, 64), num_steps=96, return_weights=True)
ef = (f.cpu()-st['rgb']).abs().amax(-1)[0]; ew = (w.cpu()-st['weights']).abs().amax((-1,-2))[0]; ed=(d.cpu()-st['depth']).abs()[0,:,0]
print('max feat err', ef.max().item(), 'rays>5e-5:', (ef>5e-5).sum().item(), 'max w err', ew.max().item(), 'depth', ed.max().item())
r = ef.argmax().item()
print('ray', r, 'w err', ew[r].item())
dw = (w.cpu()-st['weights'])[0, r, :, 0]
print('w diff top', dw.abs().topk(4))
print('w ref top', st['weights'][0, r, :, 0].topk(4))
print('sigma at those', st['raw'][0, r, :, -1][dw.abs().topk(4).indices])
print('raw max abs', st['raw'][0,r].abs().max().item(), 'rgb ref max', st['rgb'][0,r].abs().max().item())
print('feat diff', (f.cpu()-st['rgb'])[0, r][:8])
# per-sample decode error: sample_voxel on the oracle's world points
pw = st['points_world'][0, r]   # [S,3]
sv = render.sample_voxel(tex[sub], seg[sub], heads, pw[None].to(DEV))[0].cpu()
print('per-sample decode err max', (sv - st['raw'][0, r]).abs().max().item(), 'at', (sv - st['raw'][0, r]).abs().amax(-1).argmax().item())
