"""The slice of the reference's `dnnlib` the hot path uses: EasyDict and the network helpers of dnnlib/util.py."""

from . import util
from .util import EasyDict

__all__ = ['util', 'EasyDict']
