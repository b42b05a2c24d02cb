"""2D convolution with optional up/downsampling (reference torch_utils/ops/conv2d_resample.py:46).

cuDNN does the convolution (transposed, stride 2, for up=2 -- the cheap way to convolve a zero-stuffed image);
the FIR resampling runs in ide3d_upfirdn2d.  Padding is applied once, in front, like the reference.
"""

import torch

from . import conv2d_gradfix, upfirdn2d
from .upfirdn2d import _get_filter_size, _parse_padding


def _conv(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    kh, kw = w.shape[2], w.shape[3]
    if not flip_weight and (kw > 1 or kh > 1):
        w = w.flip([2, 3])          # conv2d() is a correlation; flip for a true convolution
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False, fir_epilogue=None, w_transposed=None):
    """w_transposed (extension, up > 1, groups == 1): `w.transpose(0, 1)` already materialised in the layout cuDNN wants -- the
    transposed view is not contiguous, so aten copies the whole weight on every call otherwise (9 MB for a 512x512x3x3 layer).
    fir_epilogue (extension, up > 1 and down == 1 only): keyword arguments of `upfirdn2d.upfirdn2d_epilogue` -- the
    modulated-convolution tail is then applied by the FIR pass that follows the transposed convolution."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert fir_epilogue is None or (up > 1 and down == 1 and not (int(w.shape[2]) == 1 and int(w.shape[3]) == 1))
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1 and isinstance(groups, int) and groups >= 1
    out_channels, in_per_group, kh, kw = [int(v) for v in w.shape]
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _parse_padding(padding)

    # fold the resampling filters' own padding into the single front padding (:84-93)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2

    if kw == 1 and kh == 1 and down > 1 and up == 1:                    # 1x1: decimate first
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:                    # 1x1: convolve first
        x = _conv(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d.upfirdn2d(x=x, f=f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:                                            # filter, then strided conv
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:                                                          # transposed strided conv, then filter
        if groups == 1:
            w = w_transposed if w_transposed is not None else w.transpose(0, 1)
        else:
            w = w.reshape(groups, out_channels // groups, in_per_group, kh, kw).transpose(1, 2)
            w = w.reshape(groups * in_per_group, out_channels // groups, kh, kw)
        px0 -= kw - 1; px1 -= kw - up
        py0 -= kh - 1; py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv(x, w, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        if fir_epilogue is not None:
            assert down == 1
            return upfirdn2d.upfirdn2d_epilogue(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2,
                                                flip_filter=flip_filter, **fir_epilogue)
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
        return x
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:             # plain conv2d with symmetric padding
        return _conv(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    x = upfirdn2d.upfirdn2d(x=x, f=None, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    return _conv(x, w, groups=groups, flip_weight=flip_weight)
