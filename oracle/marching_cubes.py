"""Oracle (test infrastructure): marching cubes on the CPU, numpy.

render_mesh.py:30-32 / dnnlib/geometry.py:282-286 call `mcubes.marching_cubes(volume, threshold)` -- PyMCubes (environment.yml:29,
unpinned), a third-party dependency that is absent from the reference tree.  What is restated here is the published algorithm
(Lorensen & Cline, "Marching cubes", SIGGRAPH 1987): per cell the 8 corner signs select the cut edges, the cuts are linked into closed
loops over the cell's faces, the loops are triangulated, vertices sit on the cut edges at the linearly interpolated iso-crossing, and
vertices are shared between neighbouring cells.  PARITY UNPINNED against PyMCubes itself (vertex / triangle order, the treatment of
value == threshold and the choice on ambiguous faces may differ); tests/test_mesh.py pins the geometry instead -- watertightness, Euler
characteristic, vertices on the iso-level, area and volume of analytic shapes -- and the CUDA kernels against this file bit for bit.
Never imported by the product package."""

import numpy as np


def _tables():
    corners = np.array([[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)])
    edge_corner = [(c0, c0 | (1 << a)) for a in range(3) for c0 in range(8) if not (c0 >> a) & 1]
    index = {frozenset(ec): e for e, ec in enumerate(edge_corner)}
    faces = []
    for a in range(3):
        b, c = [(1, 2), (0, 2), (0, 1)][a]
        for val in (0, 1):
            ring = []
            for vb, vc in ((0, 0), (1, 0), (1, 1), (0, 1)):
                p = [0, 0, 0]
                p[a], p[b], p[c] = val, vb, vc
                ring.append(p[0] | (p[1] << 1) | (p[2] << 2))
            faces.append(ring)
    table = []
    for cfg in range(256):
        ins = [(cfg >> i) & 1 for i in range(8)]
        nbr = {}
        for ring in faces:
            f = [ins[c] for c in ring]
            e = [index[frozenset((ring[k], ring[(k + 1) % 4]))] for k in range(4)]
            cut = [k for k in range(4) if f[k] != f[(k + 1) % 4]]
            seg = [(e[cut[0]], e[cut[1]])] if len(cut) == 2 else [(e[(k - 1) % 4], e[k]) for k in range(4) if f[k]] if len(cut) == 4 else []
            for p, q in seg:
                nbr.setdefault(p, []).append(q)
                nbr.setdefault(q, []).append(p)
        done, tris = set(), []
        for start in sorted(nbr):
            if start in done:
                continue
            loop, prev, cur = [start], None, start
            done.add(start)
            while True:
                nxt = nbr[cur][0] if prev is None else [x for x in nbr[cur] if x != prev][0]
                if nxt == start:
                    break
                loop.append(nxt)
                done.add(nxt)
                prev, cur = cur, nxt
            mid = np.array([(corners[edge_corner[k][0]] + corners[edge_corner[k][1]]) / 2.0 for k in loop])
            out_dir = sum((corners[edge_corner[k][1]] - corners[edge_corner[k][0]]) * (1 if ins[edge_corner[k][0]] else -1) for k in loop)
            normal = sum(np.cross(mid[i], mid[(i + 1) % len(loop)]) for i in range(len(loop)))
            if np.dot(normal, out_dir) < 0:
                loop = loop[::-1]
            tris += [(loop[0], loop[i], loop[i + 1]) for i in range(1, len(loop) - 1)]
        table.append(tris)
    return table, edge_corner


_TABLE, _EDGES = _tables()


def marching_cubes(volume, threshold):
    """volume [nx, ny, nz] -> (vertices [V,3] float32, triangles [T,3] int64); inside = value >= threshold; vertices ordered by the id
    of the grid edge they cut ((lower corner flat index) * 3 + axis), triangles by cell (x-major) and table order -- the CUDA path's order."""
    v = np.asarray(volume, np.float32)
    nx, ny, nz = v.shape
    thr = np.float32(threshold)
    ids, pos = [], []
    for x in range(nx - 1):
        for y in range(ny - 1):
            for z in range(nz - 1):
                val = [v[x + (i & 1), y + ((i >> 1) & 1), z + ((i >> 2) & 1)] for i in range(8)]
                cfg = sum((1 << i) for i in range(8) if not val[i] < thr)
                for t in _TABLE[cfg]:
                    for e in t:
                        c0, c1 = _EDGES[e]
                        axis = e >> 2
                        p = [x + (c0 & 1), y + ((c0 >> 1) & 1), z + ((c0 >> 2) & 1)]
                        f0, f1 = val[c0], val[c1]
                        tt = np.float32(0.5) if f1 == f0 else np.float32(thr - f0) / np.float32(f1 - f0)
                        ids.append(((p[0] * ny + p[1]) * nz + p[2]) * 3 + axis)
                        q = [np.float32(p[0]), np.float32(p[1]), np.float32(p[2])]
                        q[axis] = np.float32(q[axis] + tt)
                        pos.append(q)
    if not ids:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int64)
    ids = np.array(ids, np.int64)
    uniq, inverse = np.unique(ids, return_inverse=True)
    vertices = np.zeros((len(uniq), 3), np.float32)
    vertices[inverse] = np.array(pos, np.float32)
    return vertices, inverse.reshape(-1, 3).astype(np.int64)


def mesh_stats(vertices, triangles):
    """(is_closed_2manifold, euler_characteristic, area, signed_volume) of an indexed triangle mesh."""
    tri = np.asarray(triangles)
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])
    und = np.sort(e, 1)
    _, counts = np.unique(und, axis=0, return_counts=True)
    # every undirected edge in exactly two triangles, once in each direction
    key = e[:, 0].astype(np.int64) * (tri.max() + 1) + e[:, 1]
    rev = e[:, 1].astype(np.int64) * (tri.max() + 1) + e[:, 0]
    closed = bool((counts == 2).all()) and len(np.unique(key)) == len(key) and set(key.tolist()) == set(rev.tolist())
    V, E, F = len(vertices), len(counts), len(tri)
    p = np.asarray(vertices, np.float64)
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1).sum()
    vol = np.einsum('ij,ij->i', a, np.cross(b, c)).sum() / 6.0
    return closed, V - E + F, area, vol
