"""Minimal MRC2014 volume writer / reader -- the on-disk output of extract_shapes.py:191-192.

The reference writes the sigma grid through `mrcfile.new_mmap(path, overwrite=True, shape=grid.shape, mrc_mode=2)` and
`mrc.data[:] = grid`.  `mrcfile` is an optional dependency that is not always installed; this module writes the same file
(1024-byte MRC2014 header, mode 2 = little-endian float32, C-order data with x fastest) with numpy only, and `new_mmap`
mirrors the one call the reference makes so that `compat.install()` can stand in for the missing package.
"""

import contextlib
import os
import struct

import numpy as np

_MODES = {0: np.int8, 1: np.int16, 2: np.float32, 6: np.uint16}


def _header(shape, mode, voxel_size, stats):
    nz, ny, nx = [int(v) for v in shape]
    h = bytearray(1024)
    struct.pack_into('<3i', h, 0, nx, ny, nz)                    # NX NY NZ (columns, rows, sections)
    struct.pack_into('<i', h, 12, mode)                          # MODE
    struct.pack_into('<3i', h, 16, 0, 0, 0)                      # NXSTART NYSTART NZSTART
    struct.pack_into('<3i', h, 28, nx, ny, nz)                   # MX MY MZ
    struct.pack_into('<3f', h, 40, nx * voxel_size, ny * voxel_size, nz * voxel_size)   # CELLA
    struct.pack_into('<3f', h, 52, 90.0, 90.0, 90.0)             # CELLB
    struct.pack_into('<3i', h, 64, 1, 2, 3)                      # MAPC MAPR MAPS
    struct.pack_into('<3f', h, 76, *stats[:3])                   # DMIN DMAX DMEAN
    struct.pack_into('<i', h, 88, 1)                             # ISPG: 1 = volume
    struct.pack_into('<i', h, 92, 0)                             # NSYMBT
    struct.pack_into('<i', h, 108, 20140)                        # NVERSION
    h[208:212] = b'MAP '
    h[212:216] = bytes([0x44, 0x44, 0x00, 0x00])                 # MACHST: little endian
    struct.pack_into('<f', h, 216, stats[3])                     # RMS
    struct.pack_into('<i', h, 220, 0)                            # NLABL
    return bytes(h)


def write_mrc(path, volume, voxel_size=0.0, overwrite=True):
    """volume [nz, ny, nx] (numpy or torch, any float dtype) -> float32 MRC2014 file."""
    if hasattr(volume, 'detach'):
        volume = volume.detach().cpu().numpy()
    v = np.ascontiguousarray(volume, dtype='<f4')
    assert v.ndim == 3
    if os.path.exists(path) and not overwrite:
        raise ValueError(f'{path} exists')
    stats = (float(v.min()), float(v.max()), float(v.mean()), float(v.std())) if v.size else (0.0, 0.0, 0.0, 0.0)
    with open(path, 'wb') as f:
        f.write(_header(v.shape, 2, float(voxel_size), stats))
        f.write(v.tobytes())


def read_mrc(path):
    """-> (volume [nz, ny, nx], header dict).  Only what write_mrc / mrcfile volumes need."""
    with open(path, 'rb') as f:
        h = f.read(1024)
        nx, ny, nz, mode = struct.unpack_from('<4i', h, 0)
        nsymbt = struct.unpack_from('<i', h, 92)[0]
        assert h[208:212] == b'MAP ' and mode in _MODES
        f.seek(1024 + nsymbt)
        data = np.frombuffer(f.read(), dtype=np.dtype(_MODES[mode]).newbyteorder('<'), count=nx * ny * nz).reshape(nz, ny, nx)
    hdr = dict(nx=nx, ny=ny, nz=nz, mode=mode, cella=struct.unpack_from('<3f', h, 40), dmin=struct.unpack_from('<f', h, 76)[0],
               dmax=struct.unpack_from('<f', h, 80)[0], dmean=struct.unpack_from('<f', h, 84)[0], ispg=struct.unpack_from('<i', h, 88)[0],
               nversion=struct.unpack_from('<i', h, 108)[0])
    return data, hdr


class _Pending:
    def __init__(self, shape):
        self.data = np.zeros(shape, dtype=np.float32)


@contextlib.contextmanager
def new_mmap(name, shape, mrc_mode=2, fill=None, overwrite=False, extended_header=None, exttyp=None):
    """Stand-in for `mrcfile.new_mmap` as extract_shapes.py:191 uses it: yields an object whose `.data` array is written on
    exit.  (Not memory mapped: the 256^3 float32 grid of config 4 is 64 MB.)"""
    if mrc_mode != 2:
        raise NotImplementedError('ide3d_b200.mrc.new_mmap: only mode 2 (float32) volumes')
    if os.path.exists(name) and not overwrite:
        raise ValueError(f'File {name} already exists; set overwrite=True to overwrite it')
    p = _Pending(tuple(int(v) for v in shape))
    if fill is not None:
        p.data[...] = fill
    yield p
    write_mrc(name, p.data, overwrite=True)
