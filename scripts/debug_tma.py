import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ide3d_b200.torch_utils import custom_ops
custom_ops.verbosity = 'none'
from ide3d_b200.torch_utils.ops import upfirdn2d
f = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')
x = torch.randn(1, 2, 64, 64, device='cuda')
y = upfirdn2d.upsample2d(x, f); torch.cuda.synchronize()
print('ok', y.shape, flush=True)
