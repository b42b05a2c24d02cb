"""CPU: the oracle replays every recorded reference output (tests/golden/*.npz, produced by running the reference's
own functions in the build container -- tests/golden/make_golden.py).  This is what pins the oracle."""

import numpy as np
import pytest
import torch

from conftest import T, assert_close, load_golden
from oracle import camera as ocam, ops as oops, renderer as orr


def test_rays():
    for tag in 'ab':
        g = load_golden('rays_' + tag)
        p, z, d = orr.initial_rays(int(g['n']), int(g['num_steps']), float(g['fov']), tuple(int(v) for v in g['resolution']),
                                   float(g['ray_start']), float(g['ray_end']))
        assert_close(p, g['points'], 2e-6, what='points')
        assert_close(z, g['z_vals'], 2e-6, what='z_vals')
        assert_close(d, g['rays_d_cam'], 2e-6, what='dirs')


def test_transform():
    g = load_golden('transform')
    pj, zo = orr.perturb(T(g['points']), T(g['z_vals']), T(g['rays_d_cam']), T(g['u']))
    pw, dw, ow = orr.to_world(pj, T(g['rays_d_cam']), T(g['camera']))
    assert_close(zo, g['z_jit'], 2e-6)
    assert_close(pw, g['points_world'], 2e-6)
    assert_close(dw, g['dirs_world'], 2e-6)
    assert_close(ow, g['origins_world'], 2e-6)


def test_camera():
    g = load_golden('camera')
    for i, (h, v) in enumerate(zip(g['h'], g['v'])):
        o, _, _ = ocam.sample_camera_positions(n=1, r=float(g['radius']), horizontal_mean=float(h), vertical_mean=float(v), mode=None)
        assert_close(o, g[f'origin{i}'], 2e-6)
        assert_close(ocam.create_cam2world_matrix(-o, o), g[f'c2w{i}'], 2e-6)
        assert_close(ocam.look_at_pose(float(h), float(v), g['lookat'], radius=float(g['radius'])), g[f'lookat{i}'], 2e-6)
    # frontal pose == the hard-coded label of gen_images.py:87
    o, _, _ = ocam.sample_camera_positions(n=1, r=2.7, mode=None)
    m = ocam.create_cam2world_matrix(-o, o)[0]
    assert_close(m, np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2.7], [0, 0, 0, 1]], np.float32), 1e-6)


def test_triplane():
    g = load_golden('triplane')
    assert_close(orr.sample_triplane(T(g['coords']), T(g['grid'])), g['feat'], 3e-6)
    assert_close(orr.sample_triplane_torch(T(g['coords']), T(g['grid'])), g['feat'], 1e-6)


@pytest.mark.parametrize('case,kw', [
    ('softplus', dict(clamp_mode='softplus')), ('relu', dict(clamp_mode='relu')),
    ('lastback', dict(clamp_mode='softplus', last_back=True)),
    ('white', dict(clamp_mode='softplus', white_back=True, max_depth=3.5)),
    ('fillw', dict(clamp_mode='relu', fill_mode='weight'))])
def test_integration(case, kw):
    g = load_golden('integration')
    rgb, dep, w = orr.composite(T(g['rgb_sigma']), T(g['rays_d_cam']), T(g['z_vals']), **kw)
    assert_close(rgb, g[case + '_rgb'], 5e-6)
    assert_close(dep, g[case + '_depth'], 5e-6)
    assert_close(w, g[case + '_weights'], 2e-6)


def test_integration_requires_clamp_mode():
    g = load_golden('integration')
    with pytest.raises(ValueError):
        orr.composite(T(g['rgb_sigma']), T(g['rays_d_cam']), T(g['z_vals']), clamp_mode=None)


def test_sample_pdf():
    g = load_golden('sample_pdf')
    assert_close(orr.sample_pdf(T(g['bins']), T(g['weights']), 8, det=True), g['det'], 2e-6)
    assert_close(orr.sample_pdf(T(g['bins']), T(g['weights']), 8, det=False, u=T(g['u'])), g['rnd'], 2e-6)


def test_chain_and_voxel():
    g = load_golden('chain')
    dec = orr.Decoder(g['w1'], g['b1'], g['w2'], g['b2'])
    res = tuple(int(v) for v in g['resolution'])
    st = orr.render_frames(T(g['planes_tex']), T(g['planes_seg']), dec, T(g['camera']), num_steps=int(g['num_steps']),
                           resolution=res, box_scale=float(g['box_scale']), jitter_u=T(g['u']), return_stages=True)
    assert_close(st['points_world'], g['points_world'], 2e-6)
    assert_close(st['raw'], g['raw'], 2e-5)
    assert_close(st['rgb'], g['rgb'], 1e-5)
    assert_close(st['depth'], g['depth'], 1e-5)
    assert_close(st['weights'], g['weights'], 1e-5)
    d = load_golden('chain_dense')
    dd = orr.Decoder(d['w1'], d['b1'], d['w2'], d['b2'])
    rgb, dep, w = orr.render_frames(T(g['planes_tex']), T(g['planes_seg']), dd, T(g['camera']), num_steps=int(g['num_steps']),
                                    resolution=res, box_scale=float(g['box_scale']), jitter_u=T(g['u']), clamp_mode='relu',
                                    last_back=True)
    assert_close(rgb, d['rgb'], 1e-5); assert_close(dep, d['depth'], 1e-5); assert_close(w, d['weights'], 1e-5)
    v = load_golden('voxel')
    assert_close(orr.sample_voxel(T(g['planes_tex']), T(g['planes_seg']), dec, T(v['points']), float(g['box_scale'])), v['out'], 1e-5)


def test_create_samples_quirk():
    g = load_golden('create_samples')
    s, origin, vs = orr.create_samples(int(g['N']), [0, 0, 0], float(g['cube_length']))
    assert_close(s, g['samples'], 0.0)           # bit-exact
    # the quirk itself: y/x voxel indices are fractional because of true division (extract_shapes.py:84-86)
    raw = (s[0, :, 1] - origin[1]) / vs
    assert (raw - raw.round()).abs().max() > 0.05


def test_hash_uniform_is_uniform_and_deterministic():
    idx = np.arange(200000, dtype=np.uint64)
    u = orr.hash_uniform(idx, 1234567891011)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    assert np.array_equal(u, orr.hash_uniform(idx, 1234567891011))
    assert not np.array_equal(u, orr.hash_uniform(idx, 1234567891012))
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 1e-2


# ---------------------------------------------------------------- ops
def test_bias_act():
    g = load_golden('bias_act')
    x, b = T(g['x']), T(g['b'])
    for act in oops.ACTIVATIONS:
        assert_close(oops.bias_act(x, b, 1, act), g[act + '_d'], 2e-6, what=act)
        assert_close(oops.bias_act(x, b, 1, act, alpha=0.3, gain=1.7, clamp=0.9), g[act + '_c'], 2e-6, what=act)
    assert_close(oops.bias_act(T(g['x2d']), b, 1, 'lrelu'), g['dim1_2d'], 2e-6)
    assert_close(oops.bias_act(x, None, 1, 'swish'), g['nobias'], 2e-6)


UPFIR_CASES = {
    'up2_4x4': (2, 1, [2, 1, 2, 1], False, 4.0), 'down2_4x4': (1, 2, [1, 1, 1, 1], False, 1.0),
    'filt_4x4': (1, 1, [1, 1, 1, 1], False, 4.0), 'filt_flip': (1, 1, [2, 1, 2, 1], True, 1.0),
    'asym': ((2, 1), (1, 2), [3, 0, -1, 2], False, 0.5), 'sep8': (2, 2, [3, 4, 4, 3], False, 1.0),
    'up4_down1': ((4, 4), (1, 1), [2, 2, 2, 2], False, 16.0), 'crop': (1, 1, [-1, -2, 0, -1], False, 1.0),
    'ident': (1, 1, 0, False, 1.0),
}


def test_upfirdn2d():
    g = load_golden('upfirdn2d')
    x = T(g['x'])
    for name, (up, down, pad, flip, gain) in UPFIR_CASES.items():
        f = T(g[name + '_f']) if name + '_f' in g else None
        assert_close(oops.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain), g[name], 3e-6, what=name)
    f = oops.setup_filter([1, 3, 3, 1])
    assert_close(oops.upsample2d(x, f), g['upsample2d'], 3e-6)
    assert_close(oops.downsample2d(x, f), g['downsample2d'], 3e-6)
    assert_close(oops.filter2d(x, f), g['filter2d'], 3e-6)


FLRELU_CASES = {
    'su2_sd2': dict(up=2, down=2, padding=[5, 6, 5, 6], clamp=1.5),
    'su2_sd1': dict(up=2, down=1, padding=[5, 6, 5, 6], clamp=None),
    'su1_sd2': dict(up=1, down=2, padding=[5, 6, 5, 6], clamp=2.0, slope=0.1, gain=1.3),
    'fu2_fd2': dict(up=2, down=2, padding=[7, 8, 7, 8], clamp=1.0, flip_filter=True),
    'su4_sd2': dict(up=4, down=2, padding=[17, 18, 17, 18], clamp=0.8),
    'plain': dict(up=1, down=1, padding=0, clamp=0.7),
}


def test_filtered_lrelu():
    g = load_golden('filtered_lrelu')
    x, b = T(g['x']), T(g['b'])
    for k, kw in FLRELU_CASES.items():
        fu = T(g[k + '_fu']) if k + '_fu' in g else None
        fd = T(g[k + '_fd']) if k + '_fd' in g else None
        assert_close(oops.filtered_lrelu(x, fu=fu, fd=fd, b=b, **kw), g[k], 5e-6, what=k)


def test_conv2d_resample():
    g = load_golden('conv2d_resample')
    x, f = T(g['x']), T(g['f'])
    cases = {'up2_k3': dict(w='w3', f=f, up=2, padding=1, flip_weight=False),
             'up2_k3_g2': dict(w='wg', f=f, up=2, padding=1, groups=2, flip_weight=False),
             'same_k3': dict(w='w3', padding=1), 'down2_k3': dict(w='w3', f=f, down=2, padding=1),
             'up2_k1': dict(w='w1', f=f, up=2), 'down2_k1': dict(w='w1', f=f, down=2)}
    for k, kw in cases.items():
        kw = dict(kw)
        kw['w'] = T(g[kw['w']])
        assert_close(oops.conv2d_resample(x, **kw), g[k], 3e-5, what=k)


def _load_prefixed(g, prefix):
    return {k[len(prefix) + 1:]: torch.from_numpy(np.asarray(g[k])) for k in g if k.startswith(prefix + '/')}


def test_networks_restatement_loads_reference_state_dicts_and_reproduces_outputs():
    """a13: tests/golden/networks.npz holds state_dicts and outputs of the reference's OWN classes (inversion/networks.py
    MappingNetwork, SynthesisBlock, SegSynthesisBlock; generated by make_golden.py::g_networks).  The product's modules load those
    state_dicts unchanged (same parameter names and shapes) and, evaluated with the oracle ops, reproduce the recorded outputs --
    in both forms of modulated_conv2d (eval: the reference's grouped weight-modulated conv vs the product's activation scaling;
    train: activation scaling on both sides) and in both layouts."""
    from ide3d_b200.training import networks as nw
    from oracle.backend import cpu_reference_ops
    g = load_golden('networks')
    T_ = lambda k: torch.from_numpy(np.asarray(g[k]))
    mm = nw.MappingNetwork(z_dim=16, c_dim=25, w_dim=12, num_ws=5, num_layers=2).eval()
    mm.load_state_dict(_load_prefixed(g, 'map'))
    with torch.no_grad(), cpu_reference_ops():
        assert (mm(T_('map_z'), T_('map_c'), truncation_psi=0.7, truncation_cutoff=3) - T_('map_ws')).abs().max() < 2e-6
    for layout in (True, False):
        with torch.no_grad(), cpu_reference_ops(reference_layout=not layout):
            for tag, in_ch in (('b0', 0), ('b1', 16)):
                mb = nw.SynthesisBlock(in_ch, 8, w_dim=12, resolution=16, img_channels=6, is_last=False)
                mb.load_state_dict(_load_prefixed(g, tag))
                x = T_(f'{tag}_x') if in_ch else None
                for mode in ('eval', 'train'):
                    getattr(mb, mode)()
                    img = T_(f'{tag}_img').clone() if in_ch else None
                    xo, io = mb(x, img, T_(f'{tag}_ws'), noise_mode='const')
                    assert (xo - T_(f'{tag}_{mode}_x')).abs().max() < 1e-5 and (io - T_(f'{tag}_{mode}_img')).abs().max() < 1e-5, (tag, mode, layout)
            ms = nw.SegSynthesisBlock(16, 8, w_dim=12, resolution=16, img_channels=6, seg_channels=4, is_last=False).eval()
            ms.load_state_dict(_load_prefixed(g, 'seg'))
            xo, io, so = ms(T_('seg_x'), T_('seg_img').clone(), T_('seg_ws'), condition_img=T_('seg_seg').clone(), noise_mode='const')
            for got, key in ((xo, 'seg_out_x'), (io, 'seg_out_img'), (so, 'seg_out_seg')):
                assert (got - T_(key)).abs().max() < 1e-5, (key, layout)


def test_oracle_hierarchical_matches_reference_stage_composition():
    """chain_hier.npz: two-pass render composed in tests/golden/make_golden.py from the REFERENCE's own stage functions
    (get_initial_rays_trig, transform_sampled_points, sample_from_triplane, fancy_integration, sample_pdf :224-265)."""
    from oracle import renderer as orr
    g = load_golden('chain_hier')
    dec = orr.Decoder(*[T(g[k]) for k in ('w1', 'b1', 'w2', 'b2')])
    res = tuple(int(v) for v in g['resolution'])
    S, NI = int(g['num_steps']), int(g['n_importance'])
    rgb, depth, w, z = orr.render_frames_hierarchical(T(g['planes_tex']), T(g['planes_seg']), dec, T(g['camera']), num_steps=S,
                                                      n_importance=NI, resolution=res, box_scale=float(g['box_scale']),
                                                      jitter_u=T(g['u']), importance_u=T(g['importance_u']))
    assert (z - T(g['z_all'])).abs().max() < 2e-6
    assert (rgb - T(g['rgb'])).abs().max() < 1e-5 and (depth - T(g['depth'])).abs().max() < 1e-5
    assert (w - T(g['weights'])).abs().max() < 1e-5


def test_oracle_hierarchical_composition_properties():
    """The two-pass composition around the pinned sample_pdf: merged depths ascending and inside [ray_start - h, ray_end + h],
    S + n_importance samples per ray, weights a sub-probability; importance depths fall where the coarse weights are."""
    import math
    from oracle import camera as ocam, renderer as orr
    g = torch.Generator().manual_seed(0)
    tex = torch.nn.functional.interpolate(torch.randn(1, 96, 5, 5, generator=g), size=(16, 16), mode='bicubic', align_corners=True)
    seg = torch.nn.functional.interpolate(torch.randn(1, 96, 5, 5, generator=g), size=(16, 16), mode='bicubic', align_corners=True)
    dec = orr.Decoder.random(hidden=64, seed=2)
    cam = torch.from_numpy(ocam.look_at_pose(np.array([[math.pi / 2]], np.float32), np.array([[math.pi / 2]], np.float32), [0, 0, 0.2], radius=2.7, batch_size=1))
    S, NI = 12, 8
    u = torch.rand(1, 16, S, 1, generator=g)
    ui = torch.rand(16, NI, generator=g)
    rgb, depth, w, z = orr.render_frames_hierarchical(tex, seg, dec, cam, num_steps=S, n_importance=NI, resolution=(4, 4), jitter_u=u, importance_u=ui)
    assert z.shape == (1, 16, S + NI, 1) and w.shape == (1, 16, S + NI, 1) and rgb.shape == (1, 16, 51)
    assert bool((z[:, :, 1:] >= z[:, :, :-1]).all())
    h = (3.3 - 2.25) / (S - 1)
    assert z.min() >= 2.25 - h and z.max() <= 3.3 + h
    assert bool((w >= 0).all()) and float(w.sum(2).max()) <= 1 + 1e-5
    st = orr.render_frames(tex, seg, dec, cam, num_steps=S, resolution=(4, 4), jitter_u=u, return_stages=True)
    zc, wc = st['z_vals'].reshape(16, S), st['weights'].reshape(16, S)
    # every importance depth lies in a bin (between two coarse midpoints) and bins with larger weight receive more of them on average
    mids = 0.5 * (zc[:, :-1] + zc[:, 1:])
    fine = orr.sample_pdf(mids, wc[:, 1:-1] + 1e-5, NI, u=ui)
    assert bool((fine >= mids[:, :1]).all()) and bool((fine <= mids[:, -1:]).all())
