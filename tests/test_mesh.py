"""Marching cubes (the step after extract_shapes' sigma grid; render_mesh.py:30-32).  PyMCubes is absent (third-party,
unpinned), so the geometry is pinned by properties of analytic shapes (CPU, oracle) and the CUDA kernels against the oracle bit for bit."""

import numpy as np
import pytest
import torch

from oracle import marching_cubes as omc


def _sphere(n, r, centre=None):
    c = (n - 1) / 2.0 if centre is None else centre
    g = np.arange(n, dtype=np.float32)
    x, y, z = np.meshgrid(g, g, g, indexing='ij')
    return (r - np.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2)).astype(np.float32)      # >= 0 inside


def test_tables_generated_by_product_and_oracle_agree_and_have_the_classic_shape():
    from ide3d_b200.mesh import build_tables
    tri, ntri, ec = build_tables()
    assert int(ntri.sum()) == 820 and int(ntri.max()) == 5 and ntri[0] == 0 and ntri[255] == 0
    assert np.bincount(ntri).tolist() == [2, 16, 50, 80, 76, 32]
    for cfg in range(256):
        mine = [tuple(int(v) for v in tri[cfg, 3 * k:3 * k + 3]) for k in range(ntri[cfg])]
        assert mine == [tuple(t) for t in omc._TABLE[cfg]]
        assert all(v == -1 for v in tri[cfg, 3 * ntri[cfg]:])
        # a cell's loops use each cut edge exactly... once per loop vertex: every cut edge appears, no uncut edge does
        cut = {e for e, (c0, c1) in enumerate(ec) if ((cfg >> c0) & 1) != ((cfg >> c1) & 1)}
        assert {int(v) for t in mine for v in t} == cut
    assert [tuple(e) for e in ec] == [tuple(e) for e in omc._EDGES]


def test_oracle_sphere_is_watertight_with_the_right_area_and_volume():
    n, r = 20, 6.3
    v, t = omc.marching_cubes(_sphere(n, r), 0.0)
    closed, euler, area, vol = omc.mesh_stats(v, t)
    assert closed and euler == 2
    assert abs(area - 4 * np.pi * r * r) / (4 * np.pi * r * r) < 0.03
    assert abs(vol - 4 / 3 * np.pi * r ** 3) / (4 / 3 * np.pi * r ** 3) < 0.03 and vol > 0        # outward normals
    # vertices sit on the iso-level of the trilinear field along their edge: |distance to the sphere| below the grid's curvature error
    d = np.linalg.norm(v - (n - 1) / 2.0, axis=1)
    assert np.abs(d - r).max() < 0.08
    # two separate components and a torus-free check of the ambiguous-face rule: a checkerboard-ish field stays a closed 2-manifold
    rng = np.random.RandomState(0)
    noisy = rng.randn(9, 9, 9).astype(np.float32)
    noisy[0] = noisy[-1] = noisy[:, 0] = noisy[:, -1] = noisy[:, :, 0] = noisy[:, :, -1] = -5.0      # inside region does not touch the border
    vv, tt = omc.marching_cubes(noisy, 0.0)
    assert omc.mesh_stats(vv, tt)[0]
    assert omc.marching_cubes(np.full((4, 4, 4), -1.0, np.float32), 0.0)[1].shape == (0, 3)


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['sphere', 'noise', 'ragged'])
def test_cuda_marching_cubes_matches_oracle_bit_for_bit(case):
    from ide3d_b200 import mesh
    if case == 'sphere':
        vol = _sphere(24, 7.7, centre=11.2)
    elif case == 'noise':
        vol = np.random.RandomState(1).randn(13, 13, 13).astype(np.float32)
    else:
        vol = np.random.RandomState(2).randn(5, 9, 14).astype(np.float32)
        vol[2, 3, 4] = 0.25                                              # a value exactly on the threshold counts as inside
    thr = 0.25 if case == 'ragged' else 0.0
    v, t = mesh.marching_cubes(torch.from_numpy(vol).cuda(), thr)
    vo, to = omc.marching_cubes(vol, thr)
    assert v.shape == vo.shape and t.shape == to.shape
    assert np.array_equal(t.cpu().numpy(), to) and np.array_equal(v.cpu().numpy(), vo)
    if case != 'ragged':
        closed = omc.mesh_stats(v.cpu().numpy(), t.cpu().numpy())[0]
        assert closed == omc.mesh_stats(vo, to)[0]


@pytest.mark.gpu
def test_cuda_marching_cubes_on_a_256_cubed_grid():
    """config-4-sized grid: a sphere in a 256^3 volume -> closed surface, Euler characteristic 2, area within 0.5 %."""
    from ide3d_b200 import mesh
    n, r = 256, 90.5
    g = torch.arange(n, dtype=torch.float32, device='cuda')
    x, y, z = torch.meshgrid(g, g, g, indexing='ij')
    vol = r - torch.sqrt((x - 127.5) ** 2 + (y - 127.5) ** 2 + (z - 127.5) ** 2)
    v, t = mesh.marching_cubes(vol, 0.0)
    closed, euler, area, volu = omc.mesh_stats(v.cpu().numpy(), t.cpu().numpy())
    assert closed and euler == 2 and abs(area / (4 * np.pi * r * r) - 1) < 5e-3 and abs(volu / (4 / 3 * np.pi * r ** 3) - 1) < 5e-3
