"""GPU parity of the StyleGAN ops (bias_act, upfirdn2d, filtered_lrelu, conv2d_resample) against the recorded reference
outputs and the CPU oracle.  Tolerances: fp32 1e-5 absolute (fast-math transcendentals in bias_act, like the reference's
own --use_fast_math build), fp16 2e-2 / 1e-2 relative, fp64 1e-12."""

import math

import numpy as np
import pytest
import torch

from conftest import T, assert_close, load_golden
from oracle import ops as oops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


# ---------------------------------------------------------------------------------------------- bias_act
def test_bias_act_golden_all_activations():
    from ide3d_b200.torch_utils.ops import bias_act as ba
    g = load_golden('bias_act')
    x, b = T(g['x'], DEV), T(g['b'], DEV)
    for act in ba.activation_funcs:
        assert_close(ba.bias_act(x, b, dim=1, act=act), g[act + '_d'], 1e-5, what=act)
        assert_close(ba.bias_act(x, b, dim=1, act=act, alpha=0.3, gain=1.7, clamp=0.9), g[act + '_c'], 1e-5, what=act + '/clamp')
    assert_close(ba.bias_act(T(g['x2d'], DEV), b, dim=1, act='lrelu'), g['dim1_2d'], 1e-5)
    assert_close(ba.bias_act(x, None, act='swish'), g['nobias'], 1e-5)
    assert ba.bias_act(x, None, act='linear') is x                      # identity short-cut


# fp64: alpha/gain/clamp travel as float32 in the parameter block (bias_act.h:22-24), hence 1e-7 not 1e-15
@pytest.mark.parametrize('dtype,atol,rtol', [(torch.float32, 1e-5, 0), (torch.float16, 2e-3, 1e-2), (torch.float64, 2e-7, 2e-7)])
@pytest.mark.parametrize('shape,dim,cl', [((3, 7, 5, 9), 1, False), ((2, 8, 6, 6), 1, True), ((5, 33), 1, False), ((1031,), 0, False)])
def test_bias_act_layouts_dtypes(dtype, atol, rtol, shape, dim, cl):
    from ide3d_b200.torch_utils.ops import bias_act as ba
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g).to(dtype)
    b = torch.randn(shape[dim], generator=g).to(dtype)
    xd = x.to(DEV)
    if cl:
        xd = xd.contiguous(memory_format=torch.channels_last)
    for act, kw in (('lrelu', dict(clamp=1.0)), ('softplus', {}), ('linear', dict(gain=2.0))):
        y = ba.bias_act(xd, b.to(DEV), dim=dim, act=act, **kw)
        ref = oops.bias_act(x.double(), b.double(), dim, act, **kw)
        assert y.dtype == dtype and y.shape == x.shape
        assert_close(y, ref, atol, rtol, what=f'{act} {dtype}')


def test_bias_act_gradient_forms():
    from ide3d_b200 import _plugins
    from ide3d_b200.torch_utils.ops import bias_act as ba
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, 6, 5, 5, generator=g) * 2
    b = torch.randn(6, generator=g)
    dy = torch.randn(4, 6, 5, 5, generator=g)
    ddx = torch.randn(4, 6, 5, 5, generator=g)
    for act, spec in ba.activation_funcs.items():
        for clamp in (-1.0, 1.2):
            kw = dict(act=act, clamp=None if clamp < 0 else clamp)
            y = oops.bias_act(x, b, 1, **kw)
            d1 = _plugins.bias_act(dy.to(DEV), b.to(DEV), x.to(DEV), y.to(DEV), None, 1, 1, spec.cuda_idx,
                                   spec.def_alpha, spec.def_gain, clamp)
            assert_close(d1, oops.bias_act_grad(dy, x, b, 1, order=1, **kw), 2e-5, what=f'grad1 {act}')
            if spec.has_2nd_grad:
                d2 = _plugins.bias_act(ddx.to(DEV), b.to(DEV), x.to(DEV), y.to(DEV), dy.to(DEV), 2, 1, spec.cuda_idx,
                                       spec.def_alpha, spec.def_gain, clamp)
                assert_close(d2, oops.bias_act_grad(dy, x, b, 1, order=2, ddx=ddx, **kw), 5e-5, what=f'grad2 {act}')
    # autograd through the public op
    xg = x.to(DEV).requires_grad_(True)
    bg = b.to(DEV).requires_grad_(True)
    ba.bias_act(xg, bg, act='lrelu', clamp=1.5).backward(dy.to(DEV))
    xr = x.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    oops.bias_act(xr, br, 1, 'lrelu', clamp=1.5).backward(dy)
    assert_close(xg.grad, xr.grad, 1e-5); assert_close(bg.grad, br.grad, 1e-4)


@pytest.mark.parametrize('dtype,atol,rtol', [(torch.float32, 2e-6, 2e-6), (torch.float16, 4e-3, 1e-2), (torch.float64, 2e-7, 2e-7)])
@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('noise_kind,use_scale,use_bias', [('const', True, True), ('batch', True, True), (None, True, True),
                                                           ('const', False, True), (None, False, False), ('batch', True, False)])
def test_scaled_bias_act_is_fma_then_bias_act(dtype, atol, rtol, cl, noise_kind, use_scale, use_bias):
    """Fused modconv tail == the two reference ops it replaces: fma(x, dcoefs, noise) (networks.py:104-105) -> bias_act (:512)."""
    from ide3d_b200.torch_utils.ops import bias_act as ba
    g = torch.Generator().manual_seed(7)
    N, C, H, W = 3, 16, 12, 20
    x = torch.randn(N, C, H, W, generator=g).to(dtype)
    scale = (torch.rand(N, C, generator=g) + 0.5) if use_scale else None
    noise = None if noise_kind is None else 0.3 * torch.randn((N if noise_kind == 'batch' else 1), 1, H, W, generator=g)
    b = torch.randn(C, generator=g) if use_bias else None
    t = x
    if scale is not None:
        t = t * scale.to(dtype).reshape(N, C, 1, 1)
    if noise is not None:
        t = t + noise.to(dtype)
    want = oops.bias_act(t, None if b is None else b.to(dtype), 1, 'lrelu', None, 1.3, 0.8)
    xd = x.to(DEV)
    if cl:
        xd = xd.contiguous(memory_format=torch.channels_last)
    if noise_kind == 'const':
        nd = noise.reshape(H, W).to(DEV)                     # the [H,W] noise_const * strength form of SynthesisLayer
    else:
        nd = None if noise is None else noise.to(DEV)
    y = ba.scaled_bias_act(xd, None if scale is None else scale.to(DEV), nd, None if b is None else b.to(DEV), act='lrelu',
                           gain=1.3, clamp=0.8)
    assert y.dtype == dtype and y.shape == x.shape and y.stride() == xd.stride()
    assert_close(y, want, atol, rtol)
    # second output: the next layer's style modulation, written by the same pass
    ns = torch.rand(N, C, generator=g) + 0.5
    ya, yb = ba.scaled_bias_act(xd, None if scale is None else scale.to(DEV), nd, None if b is None else b.to(DEV), act='lrelu',
                                gain=1.3, clamp=0.8, next_scale=ns.to(DEV))
    yc = ba.scaled_bias_act(xd, None if scale is None else scale.to(DEV), nd, None if b is None else b.to(DEV), act='lrelu',
                            gain=1.3, clamp=0.8, next_scale=ns.to(DEV), only_next=True)
    assert torch.equal(ya, y) and torch.equal(yb, yc) and yb.stride() == xd.stride()
    assert_close(yb, want.to(torch.float64) * ns.to(dtype).to(torch.float64).reshape(N, C, 1, 1), 2 * atol, 2 * rtol)


def test_scaled_bias_act_autograd_composes_reference_ops():
    from ide3d_b200.torch_utils.ops import bias_act as ba
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 8, 8, 8, generator=g).to(DEV).requires_grad_(True)
    scale = (torch.rand(2, 8, generator=g) + 0.5).to(DEV).requires_grad_(True)
    noise = torch.randn(8, 8, generator=g).to(DEV)
    b = torch.randn(8, generator=g).to(DEV).requires_grad_(True)
    y = ba.scaled_bias_act(x, scale, noise, b, act='lrelu')
    y.square().sum().backward()
    with torch.no_grad():
        y2 = ba.scaled_bias_act(x.detach(), scale.detach(), noise, b.detach(), act='lrelu')
    assert_close(y, y2, 2e-6, 2e-6)
    xr = x.detach().cpu().requires_grad_(True); sr = scale.detach().cpu().requires_grad_(True); br = b.detach().cpu().requires_grad_(True)
    t = xr * sr.reshape(2, 8, 1, 1) + noise.cpu() + br.reshape(1, 8, 1, 1)
    yr = torch.nn.functional.leaky_relu(t, 0.2) * math.sqrt(2)
    yr.square().sum().backward()
    assert_close(x.grad, xr.grad, 2e-5, 1e-5); assert_close(scale.grad, sr.grad, 2e-4, 1e-5); assert_close(b.grad, br.grad, 2e-4, 1e-5)


# ---------------------------------------------------------------------------------------------- upfirdn2d
UPFIR_CASES = {
    'up2_4x4': (2, 1, [2, 1, 2, 1], False, 4.0), 'down2_4x4': (1, 2, [1, 1, 1, 1], False, 1.0),
    'filt_4x4': (1, 1, [1, 1, 1, 1], False, 4.0), 'filt_flip': (1, 1, [2, 1, 2, 1], True, 1.0),
    'asym': ((2, 1), (1, 2), [3, 0, -1, 2], False, 0.5), 'sep8': (2, 2, [3, 4, 4, 3], False, 1.0),
    'up4_down1': ((4, 4), (1, 1), [2, 2, 2, 2], False, 16.0), 'crop': (1, 1, [-1, -2, 0, -1], False, 1.0),
    'ident': (1, 1, 0, False, 1.0),
}


def test_upfirdn2d_golden():
    from ide3d_b200.torch_utils.ops import upfirdn2d as up
    g = load_golden('upfirdn2d')
    x = T(g['x'], DEV)
    for name, (u, d, pad, flip, gain) in UPFIR_CASES.items():
        f = T(g[name + '_f'], DEV) if name + '_f' in g else None
        assert_close(up.upfirdn2d(x, f, up=u, down=d, padding=pad, flip_filter=flip, gain=gain), g[name], 1e-5, what=name)
    f = up.setup_filter([1, 3, 3, 1], device=DEV)
    assert_close(f, oops.setup_filter([1, 3, 3, 1]), 0.0)
    assert_close(up.upsample2d(x, f), g['upsample2d'], 1e-5)
    assert_close(up.downsample2d(x, f), g['downsample2d'], 1e-5)
    assert_close(up.filter2d(x, f), g['filter2d'], 1e-5)


@pytest.mark.parametrize('dtype,atol,rtol', [(torch.float32, 2e-5, 0), (torch.float16, 2e-2, 1e-2), (torch.float64, 2e-7, 2e-7)])
@pytest.mark.parametrize('hw', [(64, 64), (70, 37), (129, 200)])
def test_upfirdn2d_stylegan_shapes(dtype, atol, rtol, hw):
    """The three hot StyleGAN2 specialisations + separable 12-tap passes, on tile-unfriendly sizes, contiguous and
    channels_last, every padding phase."""
    from ide3d_b200.torch_utils.ops import upfirdn2d as up
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, *hw, generator=g).to(dtype)
    f4 = oops.setup_filter([1, 3, 3, 1])
    f12 = oops.setup_filter(np.hanning(14)[1:-1].tolist())
    cases = [dict(f=f4, up=2, padding=[2, 1, 2, 1], gain=4), dict(f=f4, up=2, padding=[1, 2, 2, 1], gain=4),
             dict(f=f4, up=2, padding=[3, 0, 1, 2]), dict(f=f4, padding=[1, 1, 1, 1], gain=4),
             dict(f=f4, down=2, padding=[1, 1, 1, 1]), dict(f=f4, down=2, padding=[2, 0, 0, 2], flip_filter=True),
             dict(f=f12, up=2, padding=[5, 6, 5, 6], gain=4), dict(f=f12, down=2, padding=[5, 5, 5, 5]), dict(f=f12, padding=[6, 5, 6, 5])]
    for kw in cases:
        ref = oops.upfirdn2d(x.double(), kw['f'].double() if dtype == torch.float64 else kw['f'], **{k: v for k, v in kw.items() if k != 'f'})
        for fmt in (torch.contiguous_format, torch.channels_last):
            y = up.upfirdn2d(x.to(DEV).contiguous(memory_format=fmt), kw['f'].to(DEV), **{k: v for k, v in kw.items() if k != 'f'})
            assert y.dtype == dtype
            assert_close(y, ref, atol, rtol, what=str({k: v for k, v in kw.items() if k != 'f'}))


def test_upfirdn2d_backward_is_adjoint():
    from ide3d_b200.torch_utils.ops import upfirdn2d as up
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 3, 12, 10, generator=g)
    f = oops.setup_filter([1, 3, 3, 1])
    dy_shape = oops.upsample2d(x, f).shape
    dy = torch.randn(*dy_shape, generator=g)
    xg = x.to(DEV).requires_grad_(True)
    up.upsample2d(xg, f.to(DEV)).backward(dy.to(DEV))
    xr = x.clone().requires_grad_(True)
    oops.upsample2d(xr, f).backward(dy)
    assert_close(xg.grad, xr.grad, 2e-5)


# ---------------------------------------------------------------------------------------------- filtered_lrelu
FLRELU_CASES = {
    'su2_sd2': dict(up=2, down=2, padding=[5, 6, 5, 6], clamp=1.5),
    'su2_sd1': dict(up=2, down=1, padding=[5, 6, 5, 6], clamp=None),
    'su1_sd2': dict(up=1, down=2, padding=[5, 6, 5, 6], clamp=2.0, slope=0.1, gain=1.3),
    'fu2_fd2': dict(up=2, down=2, padding=[7, 8, 7, 8], clamp=1.0, flip_filter=True),
    'su4_sd2': dict(up=4, down=2, padding=[17, 18, 17, 18], clamp=0.8),
    'plain': dict(up=1, down=1, padding=0, clamp=0.7),
}


def test_filtered_lrelu_golden():
    from ide3d_b200.torch_utils.ops import filtered_lrelu as fl
    g = load_golden('filtered_lrelu')
    x, b = T(g['x'], DEV), T(g['b'], DEV)
    for k, kw in FLRELU_CASES.items():
        fu = T(g[k + '_fu'], DEV) if k + '_fu' in g else None
        fd = T(g[k + '_fd'], DEV) if k + '_fd' in g else None
        assert_close(fl.filtered_lrelu(x, fu=fu, fd=fd, b=b, **kw), g[k], 2e-5, what=k)


@pytest.mark.parametrize('dtype,atol,rtol', [(torch.float32, 3e-5, 0), (torch.float16, 2e-2, 2e-2)])
@pytest.mark.parametrize('hw', [(64, 64), (50, 37)])
def test_filtered_lrelu_fused_shapes(dtype, atol, rtol, hw):
    """The fused separable kernel on tile-unfriendly sizes / odd paddings / both flip settings, vs the oracle."""
    from ide3d_b200 import _plugins
    from ide3d_b200.torch_utils.ops import filtered_lrelu as fl
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(2, 3, *hw, generator=g) * 2).to(dtype)
    b = torch.randn(3, generator=g).to(dtype)
    f12 = oops.setup_filter(np.hanning(14)[1:-1].tolist())
    f8 = oops.setup_filter(np.hanning(10)[1:-1].tolist())
    f24 = oops.setup_filter(np.hanning(26)[1:-1].tolist())
    cases = [dict(fu=f12, fd=f12, up=2, down=2, padding=[10, 11, 10, 11], clamp=1.5),
             dict(fu=f12, fd=f12, up=2, down=2, padding=[5, 6, 4, 7], clamp=None, flip_filter=True, gain=0.9, slope=0.3),
             dict(fu=f8, fd=f8, up=2, down=2, padding=[3, 4, 3, 4], clamp=2.0),
             dict(fu=f24, fd=f12, up=4, down=2, padding=[17, 18, 17, 18], clamp=0.8),
             dict(fu=f12, fd=None, up=2, down=1, padding=[5, 6, 5, 6], clamp=1.0),
             dict(fu=None, fd=f12, up=1, down=2, padding=[5, 6, 5, 6], clamp=1.0),
             dict(fu=None, fd=None, up=1, down=1, padding=0, clamp=0.5)]
    for kw in cases:
        ref = oops.filtered_lrelu(x.float(), b=b.float(), **kw)
        kd = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
        y = fl.filtered_lrelu(x.to(DEV), b=b.to(DEV), **kd)
        assert y.dtype == dtype and y.shape == ref.shape
        assert_close(y, ref, atol, rtol, what=str({k: v for k, v in kw.items() if not isinstance(v, torch.Tensor)}))
    # the fused kernel really ran for the separable case (return code 0), and 2-D filters report "no kernel" (-1)
    xd, bd = x.float().to(DEV), b.float().to(DEV)
    y, so, rc = _plugins.filtered_lrelu(xd, f12.to(DEV), f12.to(DEV), bd, None, 2, 2, 10, 11, 10, 11, 0, 0, 1.4, 0.2, 1.5, False, True)
    assert rc == 0 and so is not None
    _, packed = oops.filtered_lrelu(x.float(), fu=f12, fd=f12, b=b.float(), up=2, down=2, padding=[10, 11, 10, 11], gain=1.4, slope=0.2, clamp=1.5, return_signs=True)
    sh, swb = so.shape[2], so.shape[3]
    # compare the 2-bit codes inside the active region (rows/cols the down-FIR actually reads)
    def unpack(t):
        t = t.cpu().to(torch.int32)
        return torch.stack([(t >> (2 * i)) & 3 for i in range(4)], -1).reshape(*t.shape[:-1], -1)
    mine, ref_codes = unpack(so), unpack(packed)
    sw_active = y.shape[3] * 2 - 1 + 11
    assert torch.equal(mine[:, :, :sh, :sw_active], ref_codes[:, :, :sh, :sw_active])
    f2d = torch.outer(f8, f8).to(DEV)
    assert _plugins.filtered_lrelu(xd, f2d, f2d, bd, None, 2, 2, 7, 8, 7, 8, 0, 0, 1.4, 0.2, 1.5, False, False)[2] == -1


def test_filtered_lrelu_act_sign_tensor_and_backward():
    from ide3d_b200 import _plugins
    from ide3d_b200.torch_utils.ops import filtered_lrelu as fl
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 3, 9, 21, generator=g) * 2
    gain, slope, clamp = 1.4, 0.2, 1.0
    xd = x.to(DEV).clone()
    so = _plugins.filtered_lrelu_act_(xd, None, 0, 0, gain, slope, clamp, True)
    _, packed = oops.filtered_lrelu(x, gain=gain, slope=slope, clamp=clamp, return_signs=True)
    assert so.shape == packed.shape and torch.equal(so.cpu(), packed)
    assert_close(xd, oops.bias_act(x, act='lrelu', alpha=slope, gain=gain, clamp=clamp), 1e-6)
    # sign read: y = x*gain*slope where negative, 0 where clamped
    t = torch.randn(2, 3, 9, 21, generator=g)
    td = t.to(DEV).clone()
    _plugins.filtered_lrelu_act_(td, so, 0, 0, gain, slope, clamp, False)
    pre = x * gain
    neg, cl = pre < 0, torch.where(pre < 0, pre * slope, pre).abs() > clamp
    assert_close(td, torch.where(cl, torch.zeros_like(t), torch.where(neg, t * gain * slope, t * gain)), 1e-6)
    # end-to-end gradient of the op vs autograd through the oracle
    f12 = oops.setup_filter(np.hanning(14)[1:-1].tolist())
    xg = torch.randn(1, 2, 16, 16, generator=g)
    bg = torch.randn(2, generator=g)
    kw = dict(up=2, down=2, padding=[5, 6, 5, 6], clamp=1.5)
    xr, br = xg.clone().requires_grad_(True), bg.clone().requires_grad_(True)
    yr = oops.filtered_lrelu(xr, fu=f12, fd=f12, b=br, **kw)
    dy = torch.randn(*yr.shape, generator=g)
    yr.backward(dy)
    xc, bc = xg.to(DEV).requires_grad_(True), bg.to(DEV).requires_grad_(True)
    fl.filtered_lrelu(xc, fu=f12.to(DEV), fd=f12.to(DEV), b=bc, **kw).backward(dy.to(DEV))
    assert_close(xc.grad, xr.grad, 5e-5); assert_close(bc.grad, br.grad, 5e-4)


# ---------------------------------------------------------------------------------------------- conv2d_resample
def test_conv2d_resample_golden():
    from ide3d_b200.torch_utils.ops import conv2d_resample as cr
    g = load_golden('conv2d_resample')
    x, f = T(g['x'], DEV), T(g['f'], DEV)
    cases = {'up2_k3': dict(w='w3', f=f, up=2, padding=1, flip_weight=False),
             'up2_k3_g2': dict(w='wg', f=f, up=2, padding=1, groups=2, flip_weight=False),
             'same_k3': dict(w='w3', padding=1), 'down2_k3': dict(w='w3', f=f, down=2, padding=1),
             'up2_k1': dict(w='w1', f=f, up=2), 'down2_k1': dict(w='w1', f=f, down=2)}
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for k, kw in cases.items():
            kw = dict(kw)
            kw['w'] = T(g[kw['w']], DEV)
            assert_close(cr.conv2d_resample(x, **kw), g[k], 1e-4, what=k)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_plugin_loader_contract(capsys):
    from ide3d_b200.torch_utils import custom_ops
    custom_ops._cached_plugins.clear()
    custom_ops.verbosity = 'brief'
    p = custom_ops.get_plugin('upfirdn2d_plugin', sources=['upfirdn2d.cpp', 'upfirdn2d.cu'])
    assert 'Setting up PyTorch plugin "upfirdn2d_plugin"... Done.' in capsys.readouterr().out
    assert custom_ops.get_plugin('upfirdn2d_plugin') is p and hasattr(p, 'upfirdn2d')
    with pytest.raises(KeyError):
        custom_ops.get_plugin('no_such_plugin')
    with pytest.raises(RuntimeError, match='f must be float32'):
        p.upfirdn2d(torch.zeros(1, 1, 4, 4, device=DEV), torch.ones(2, 2, device=DEV, dtype=torch.float64), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)


def test_upfirdn2d_channels_last_kernel_matches_contiguous_bitwise_shape():
    """channels_last inputs take the 4-channel-vector kernel (C % 4 == 0) and keep the memory format."""
    from ide3d_b200.torch_utils.ops import upfirdn2d as up
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 8, 33, 20, generator=g)
    f = oops.setup_filter([1, 3, 3, 1])
    xc = x.to(DEV).contiguous(memory_format=torch.channels_last)
    for fn, ref in ((up.upsample2d, oops.upsample2d), (up.downsample2d, oops.downsample2d), (up.filter2d, oops.filter2d)):
        y = fn(xc, f.to(DEV))
        assert y.is_contiguous(memory_format=torch.channels_last)
        assert_close(y, ref(x, f), 1e-5)


@pytest.mark.parametrize('dtype,atol,rtol', [(torch.float32, 2e-5, 0), (torch.float16, 2e-2, 1e-2)])
@pytest.mark.parametrize('hw', [(64, 64), (37, 70), (5, 3)])
@pytest.mark.parametrize('C,lo', [(12, 4), (96, 32)])
def test_upfirdn2d_channels_last_patch_kernel_phases(dtype, atol, rtol, hw, C, lo):
    """channels_last + the 4x4 StyleGAN2 filter: C % 4 == 0 -> upfirdn2d_cl_patch_kernel (L1 gather), C % 32 == 0 ->
    upfirdn2d_cl_tma_kernel (TMA-staged tiles); every padding phase, ragged edges, negative padding (cropping) and sliced
    (non-dense) channel ranges."""
    from ide3d_b200.torch_utils.ops import upfirdn2d as up
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, C, *hw, generator=g).to(dtype)
    f4 = oops.setup_filter([1, 3, 3, 1])
    cases = [dict(up=2, padding=[2, 1, 2, 1], gain=4), dict(up=2, padding=[1, 2, 2, 1], gain=4), dict(up=2, padding=[3, 0, 1, 2]),
             dict(up=2, padding=[2, 1, 1, 2]), dict(padding=[1, 1, 1, 1], gain=4), dict(padding=[2, 1, 0, 3]), dict(padding=[-1, 4, 3, -1]),
             dict(down=2, padding=[1, 1, 1, 1]), dict(down=2, padding=[2, 0, 0, 2], flip_filter=True)]
    xc = x.to(DEV).contiguous(memory_format=torch.channels_last)
    for kw in cases:
        if min(hw) < 4 and kw.get('padding', [0])[0] < 0:
            continue
        ref = oops.upfirdn2d(x.double(), f4, **kw)
        y = up.upfirdn2d(xc, f4.to(DEV), **kw)
        assert y.dtype == dtype and y.is_contiguous(memory_format=torch.channels_last)
        assert_close(y, ref, atol, rtol, what=str(kw))
        ys = up.upfirdn2d(xc[:, lo:C], f4.to(DEV), **kw)              # channel slice: stride_c == 1, pixel pitch C
        assert_close(ys, ref[:, lo:C], atol, rtol, what='slice ' + str(kw))


@pytest.mark.parametrize('dtype,atol,rtol', [(torch.float32, 2e-5, 0), (torch.float16, 2e-2, 1e-2)])
@pytest.mark.parametrize('CI,CW,lo', [(8, 20, 4), (32, 96, 64)])
def test_upsample2d_add_is_upsample_then_add(dtype, atol, rtol, CI, CW, lo):
    """Skip-connection step in one pass == upsample2d (networks.py:841) + add (:844) + ToRGB bias (:707); `y` may be a channel
    slice of a wider NHWC tensor.  NCHW inputs take the composed path and give the same values."""
    from ide3d_b200.torch_utils.ops import upfirdn2d as up
    g = torch.Generator().manual_seed(21)
    img = torch.randn(2, CI, 9, 13, generator=g).to(dtype)
    ywide = torch.randn(2, CW, 18, 26, generator=g).to(dtype)
    b = torch.randn(CI, generator=g).to(dtype)
    f = oops.setup_filter([1, 3, 3, 1])
    want = oops.upsample2d(img.double(), f) + ywide[:, lo:lo + CI].double() + b.double().reshape(1, -1, 1, 1)
    yd = ywide.to(DEV).contiguous(memory_format=torch.channels_last)
    keep = yd.clone()
    out = up.upsample2d_add(img.to(DEV).contiguous(memory_format=torch.channels_last), f.to(DEV), yd[:, lo:lo + CI], b.to(DEV))
    assert out.is_contiguous(memory_format=torch.channels_last) and torch.equal(yd, keep)
    assert_close(out, want, atol, rtol)
    out2 = up.upsample2d_add(img.to(DEV), f.to(DEV), ywide[:, lo:lo + CI].to(DEV), b.to(DEV))          # NCHW: composed
    assert_close(out2, want, atol, rtol)
    out3 = up.upsample2d_add(img.to(DEV).contiguous(memory_format=torch.channels_last), f.to(DEV), yd[:, lo:lo + CI], None)
    assert_close(out3, want - b.double().reshape(1, -1, 1, 1), atol, rtol)


@pytest.mark.parametrize('dtype,atol,rtol', [(torch.float32, 3e-5, 1e-5), (torch.float16, 3e-2, 2e-2)])
@pytest.mark.parametrize('act,noise_kind', [('lrelu', 'const'), ('lrelu', 'batch'), ('linear', None), ('tanh', 'const')])
@pytest.mark.parametrize('C', [8, 64])
def test_upfirdn2d_epilogue_is_fir_then_modconv_tail(dtype, atol, rtol, act, noise_kind, C):
    """FIR + demodulation/noise/bias_act in one pass == conv2d_resample.py:125 followed by networks.py:104-105, :512.
    ('tanh' is not fused by the kernel: the op composes upfirdn2d + scaled_bias_act and must give the same values.)"""
    from ide3d_b200.torch_utils.ops import upfirdn2d as up
    g = torch.Generator().manual_seed(31)
    N, H, W = 2, 19, 23                              # transposed-conv output of a 9x11 layer: (2r+1)
    x = torch.randn(N, C, H, W, generator=g).to(dtype)
    f = oops.setup_filter([1, 3, 3, 1])
    pad = [1, 1, 1, 1]
    scale = torch.rand(N, C, generator=g) + 0.5
    ns = torch.rand(N, C, generator=g) + 0.5
    b = torch.randn(C, generator=g)
    fir = oops.upfirdn2d(x.double(), f, padding=pad, gain=4)
    oh, ow = fir.shape[2:]
    noise = None if noise_kind is None else 0.3 * torch.randn((N if noise_kind == 'batch' else 1), 1, oh, ow, generator=g)
    t = fir * scale.to(dtype).double().reshape(N, C, 1, 1)
    if noise is not None:
        t = t + noise.to(dtype).double()
    want = oops.bias_act(t, b.to(dtype).double(), 1, act, None, 1.2, 2.0)
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
    nd = None if noise is None else (noise.reshape(oh, ow) if noise_kind == 'const' else noise).to(DEV)
    kw = dict(padding=pad, gain=4, scale=scale.to(DEV), noise=nd, b=b.to(DEV), act=act, act_gain=1.2, clamp=2.0)
    y = up.upfirdn2d_epilogue(xd, f.to(DEV), **kw)
    assert y.dtype == dtype and y.is_contiguous(memory_format=torch.channels_last)
    assert_close(y, want, atol, rtol)
    ya, yb = up.upfirdn2d_epilogue(xd, f.to(DEV), next_scale=ns.to(DEV), **kw)
    yc = up.upfirdn2d_epilogue(xd, f.to(DEV), next_scale=ns.to(DEV), only_next=True, **kw)
    assert torch.equal(ya, y) and torch.equal(yb, yc)
    assert_close(yb, want * ns.to(dtype).double().reshape(N, C, 1, 1), 2 * atol, 2 * rtol)
    yn = up.upfirdn2d_epilogue(x.to(DEV), f.to(DEV), **kw)                       # NCHW: composed path
    assert_close(yn, want, atol, rtol)
