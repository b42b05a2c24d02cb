"""conv2d / conv_transpose2d entry points with the reference's names (torch_utils/ops/conv2d_gradfix.py:35,40).

The north star leaves the convolution stack to cuDNN, so these forward to torch.nn.functional.  What the reference's module adds
on top of the 2021 PyTorch it was written for -- arbitrary-order gradients through a custom autograd.Function (:66-198) and the
`no_weight_gradients()` switch that skips the weight-gradient convolution during path-length / R1 regularisation (:21-33, :134,
:178) -- is provided by the autograd of current PyTorch plus one detach:
  * higher-order gradients: aten's convolution is differentiable to any order (convolution_backward has its own double-backward
    formula); `tests/test_host_logic.py::test_conv2d_gradfix_gradient_penalty` runs an R1-style penalty (gradient of a gradient)
    through both entry points and checks it against the explicit formula;
  * `with no_weight_gradients():` -- the weight enters the convolution detached, so no weight gradient is computed or accumulated
    (first or higher order), which is exactly the observable behaviour of the reference's flag.
`enabled` is accepted and ignored (there is one code path).
"""

import contextlib

import torch

enabled = False
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    try:
        yield
    finally:
        weight_gradients_disabled = old


def _w(weight):
    return weight.detach() if weight_gradients_disabled else weight


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    return torch.nn.functional.conv2d(input=input, weight=_w(weight), bias=bias, stride=stride, padding=padding,
                                      dilation=dilation, groups=groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    return torch.nn.functional.conv_transpose2d(input=input, weight=_w(weight), bias=bias, stride=stride, padding=padding,
                                                output_padding=output_padding, groups=groups, dilation=dilation)
