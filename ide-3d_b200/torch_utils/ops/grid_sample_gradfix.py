"""`grid_sample(input, grid)` pinned to bilinear / zeros padding / align_corners=False -- the semantics the
tri-plane gather is defined by (reference torch_utils/ops/grid_sample_gradfix.py:26-29).  The renderer does not
call this (its gather is fused into csrc/raymarch.cu); it is kept for callers that import it."""

import torch

enabled = False  # reference module switch (:22); the custom double-backward op is not needed for inference


def grid_sample(input, grid):
    return torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros',
                                           align_corners=False)
