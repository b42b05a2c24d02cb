#!/usr/bin/env python
"""Micro-benchmark of the StyleGAN ops (BASELINE config 5): 512x512x512-channel tensors, achieved GB/s of algorithmic
bytes (numel_in + numel_out)*sizeof(T) against the measured HBM peak.  Prints one JSON line per case.

    python scripts/bench_ops.py [--small]
"""
import argparse, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ide3d_b200.torch_utils import custom_ops
custom_ops.verbosity = 'none'
from ide3d_b200.torch_utils.ops import bias_act, filtered_lrelu, upfirdn2d


def peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    return float(json.load(open(p))['hbm_gbs']) if os.path.exists(p) else 6650.0


def timeit(fn, reps=10, warm=3):
    flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y = fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--small', action='store_true')
    args = ap.parse_args()
    C = 64 if args.small else 512
    pk = peak()
    f4 = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')
    import scipy.signal
    f12 = torch.as_tensor(scipy.signal.firwin(numtaps=12, cutoff=0.25, width=0.5), dtype=torch.float32, device='cuda')
    for dtype in (torch.float32, torch.float16):
        x = torch.randn(1, C, 512, 512, device='cuda', dtype=dtype)
        b = torch.randn(C, device='cuda', dtype=dtype)
        cases = {
            'bias_act lrelu+clamp': lambda: bias_act.bias_act(x, b, act='lrelu', clamp=256),
            'upfirdn2d upsample2d 4x4': lambda: upfirdn2d.upsample2d(x, f4),
            'upfirdn2d filter2d 4x4': lambda: upfirdn2d.filter2d(x, f4),
            'upfirdn2d downsample2d 4x4': lambda: upfirdn2d.downsample2d(x, f4),
            'filtered_lrelu up2 down2 12-tap': lambda: filtered_lrelu.filtered_lrelu(x, fu=f12, fd=f12, b=b, up=2, down=2, padding=[10, 11, 10, 11], gain=2 ** 0.5, slope=0.2, clamp=256),
        }
        for layout in ('contiguous', 'channels_last'):
            if layout == 'channels_last':
                x = x.contiguous(memory_format=torch.channels_last)
            for name, fn in cases.items():
                try:
                    ms, y = timeit(fn)
                except Exception as e:          # noqa
                    print(json.dumps({'op': name, 'dtype': str(dtype), 'layout': layout, 'error': str(e)[:200]}))
                    continue
                nbytes = (x.numel() + y.numel()) * x.element_size()
                gbs = nbytes / ms / 1e6
                print(json.dumps({'op': name, 'dtype': str(dtype).replace('torch.', ''), 'layout': layout, 'shape': list(x.shape), 'out': list(y.shape),
                                  'ms': round(ms, 4), 'algorithmic_GB': round(nbytes / 1e9, 3), 'achieved_GBs': round(gbs, 1), 'peak_GBs': pk, 'frac': round(gbs / pk, 3)}), flush=True)
                del y


if __name__ == '__main__':
    main()
