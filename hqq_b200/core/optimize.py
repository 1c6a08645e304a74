"""``optimize_weights_proximal`` -- the Quantizer.optimize_weights seam (hqq/core/optimize.py:209-259).

Same call signature and return contract as the reference's default (legacy) solver: takes the grouped
float weights plus the *inverse* scale and zero, returns ``(W_q, scale, zero)`` with W_q the un-packed
float levels.  The half-quadratic iterations, the shrinkage operator (optimize.py:96-108) and the
whole-tensor early stop all run inside ``hqq_b200_quantize_ex`` on the GPU; nothing is iterated in Python.
"""
from __future__ import annotations

from typing import Union

import torch
from torch import Tensor

from .. import ops

DEFAULT_OPT_PARAMS = {"lp_norm": 0.7, "beta": 1e1, "kappa": 1.01, "iters": 20}


@torch.inference_mode()
def optimize_weights_proximal_legacy(
    tensor: Tensor,
    scale: Tensor,
    zero: Tensor,
    min_max: list,
    axis: int = 0,
    device: Union[str, None] = None,
    opt_params: dict = DEFAULT_OPT_PARAMS,
    verbose: bool = False,
) -> tuple:
    lp_norm, beta, iters = opt_params["lp_norm"], opt_params["beta"], opt_params["iters"]
    if min_max[0] != 0:
        raise ValueError("hqq_b200: the solver clamps to [0, max]; min_max[0] must be 0 as in Quantizer.quantize")
    home = tensor.device
    dev = torch.device(device) if device is not None else home
    if dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else dev
    W = tensor.to(device=dev)
    if W.dim() != 2:
        raise ValueError("optimize_weights_proximal expects the grouped 2-D weight matrix")
    # `tensor` is already grouped: [R, gs] with per-row meta (axis=1) or [gs, C] with per-column meta (axis=0)
    R, C = W.shape
    gs = C if axis == 1 else R
    W_q8, _, zero_out, trace = ops.quantize(
        W, nbits=8, group_size=gs, axis=axis, round_zero=False, optimize=True, lp_norm=lp_norm, beta=beta, iters=iters,
        scale_init=scale, zero_init=zero, max_level=int(min_max[1]), want_trace=verbose)
    if verbose:
        info = trace["info"].tolist()
        for i, e in enumerate(trace["errors"].tolist()[: info[0]]):
            print(i, round(e, 6))
    W_q = W_q8.to(torch.float32).reshape(R, C)  # nbits=8 "packing" is one level per byte
    zero_new = zero_out.reshape(zero.shape)
    return W_q.to(home), scale.to(home), zero_new.to(home)


# Default: fast with early stopping (the reference's `optimize_weights_proximal = ..._legacy`, optimize.py:259)
optimize_weights_proximal = optimize_weights_proximal_legacy
