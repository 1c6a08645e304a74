"""Tensor-parallel shards of an ALREADY quantised layer.

The reference's serving glue re-slices quantised tensors when a checkpoint is loaded into a tensor-parallel model
(`load_merged_column_weight` / `load_row_parallel_weight` / `load_qkv_weight`, hqq/utils/vllm.py:111-170: unpack, reshape, slice,
repack).  Here the same operation keeps HQQ's own layout: BitPack's slab interleave runs along dim 0 of the grouped level matrix
(hqq/core/bitpack.py), so a shard cannot be cut out of the packed bytes -- it is unpacked, sliced by whole groups and packed again
with the shard's own slab step.  Groups are independent (hqq/core/quantize.py:102-134), so the shard's levels, scales and zeros ARE
the unsharded layer's: a tensor-parallel model built this way computes with exactly the unsharded quantisation, which quantising
each shard on its own does not guarantee (the solver's early stop looks at the whole tensor, hqq/core/optimize.py:239-247).

  column-parallel (q/k/v/gate/up: split N)  rows n0..n1 of W  -> the contiguous groups [n0*K/gs, n1*K/gs)
  row-parallel    (o/down: split K)         columns k0..k1    -> groups j0..j1 of every row; the bias stays on rank 0 only
                                                                 (the partial sums are added by the all-reduce)

`shard_quantized` is plain index arithmetic over `pack` / `unpack` callables (numpy with the oracle's in the CPU tests, torch with
this package's kernels on the GPU); `shard_hqq_linear` applies it to an `HQQLinear`.  axis = 1 only (the path's configuration).
"""
from __future__ import annotations

import copy

_FIELDS = {8: 1, 4: 2, 2: 4, 1: 8, 3: 10}
_PACKING_BITS = {"8bit_u8": 8, "4bit_u8": 4, "3bit_32": 3, "2bit_u8": 2, "1bit_u8": 1}


def shard_bounds(total: int, tp: int, rank: int, multiple: int = 1):
    """[lo, hi) of `rank`'s equal share of `total`; the share must be a whole number of `multiple`s."""
    if tp < 1 or not 0 <= rank < tp:
        raise ValueError(f"bad tensor-parallel rank {rank} of {tp}")
    if total % tp or (total // tp) % multiple:
        raise ValueError(f"{total} does not split into {tp} shards of whole multiples of {multiple}")
    per = total // tp
    return rank * per, (rank + 1) * per


def shard_quantized(W_q, meta: dict, tp: int, rank: int, parallel: str, pack, unpack):
    """(W_q, meta) of one tensor-parallel shard.  `meta` needs nbits, group_size, shape, axis, scale, zero (+ anything else, copied);
    `unpack(W_q, nbits)` returns the level matrix (3-bit: padded rows allowed), `pack(levels, nbits)` the packed tensor."""
    # the STORAGE width comes from the packing (nbits = 1.58 is stored as 2bit_u8, 5 / 6 as 8bit_u8: quantize.py:40-73), meta["nbits"]
    # itself travels on unchanged
    nbits = _PACKING_BITS[meta["packing"]] if "packing" in meta else int(meta["nbits"])
    gs, axis = int(meta["group_size"]), int(meta["axis"])
    N, K = (int(v) for v in meta["shape"])
    if axis != 1:
        raise ValueError("shard_quantized: axis=1 layers only (groups along the input dimension)")
    if parallel not in ("column", "row"):
        raise ValueError("parallel must be 'column' (split the output rows) or 'row' (split the input columns)")
    if K % gs:
        raise ValueError("group_size must divide in_features")
    Gk, R, F = K // gs, N * K // gs, _FIELDS[nbits]
    levels = unpack(W_q, nbits)[:R]                       # [R, gs]; 3-bit carries padded rows (quantize.py:190-195)
    scale, zero = meta["scale"].reshape(R, 1), meta["zero"].reshape(R, 1)
    if parallel == "column":
        n0, n1 = shard_bounds(N, tp, rank)
        lv, s, z = levels[n0 * Gk:n1 * Gk], scale[n0 * Gk:n1 * Gk], zero[n0 * Gk:n1 * Gk]
        shape = (n1 - n0, K)
    else:
        j0, j1 = shard_bounds(Gk, tp, rank)
        lv = levels.reshape(N, Gk, gs)[:, j0:j1].reshape(-1, gs)
        s = scale.reshape(N, Gk)[:, j0:j1].reshape(-1, 1)
        z = zero.reshape(N, Gk)[:, j0:j1].reshape(-1, 1)
        shape = (N, (j1 - j0) * gs)
    rows = shape[0] * shape[1] // gs
    if nbits != 3 and rows % F:
        raise ValueError(f"a shard of {rows} groups cannot be packed {F} to a byte")
    contiguous = (lambda t: t.contiguous()) if hasattr(lv, "contiguous") else (lambda t: __import__("numpy").ascontiguousarray(t))
    out_meta = {k: (copy.copy(v) if not hasattr(v, "shape") else v) for k, v in meta.items()}
    out_meta.update(shape=shape, scale=contiguous(s), zero=contiguous(z))
    return pack(contiguous(lv), nbits), out_meta


def shard_hqq_linear(layer, tp: int, rank: int, parallel: str):
    """A new `HQQLinear` holding `rank`'s shard of the quantised `layer` (on the same device, same compute dtype)."""
    import torch
    from torch import nn

    from .. import ops
    from ..core.quantize import HQQLinear
    if not layer.is_initialized():
        raise ValueError("shard_hqq_linear: the layer is not quantised")
    meta = layer.meta
    if meta.get("quant_scale") or meta.get("quant_zero") or "zero_scale" in meta:
        raise ValueError("shard_hqq_linear: quantised / offloaded scale and zero are deprecated in the reference and not supported here")
    W_q = layer.W_q.data
    if meta.get("view_as_float"):  # the same bytes viewed as the compute dtype (quantize.py:170-173)
        W_q = W_q.view(meta["unpack_view_dtype"])
    unpack = lambda t, nbits: ops.unpack(t, nbits, torch.uint8)  # noqa: E731
    Wq_s, meta_s = shard_quantized(W_q, meta, tp, rank, parallel, ops.pack, unpack)
    meta_s["shape"] = torch.Size(meta_s["shape"])  # the state_dict codec (core/utils.py) encodes torch.Size, like the reference's
    if meta.get("view_as_float"):
        Wq_s = Wq_s.view(layer.compute_dtype)
    cfg = copy.deepcopy(layer.quant_config)
    cfg["offload_meta"] = False
    new = HQQLinear(None, cfg, compute_dtype=layer.compute_dtype, device=layer.device, initialize=False)
    new.W_q = nn.Parameter(Wq_s, requires_grad=False)
    new.meta = meta_s
    if layer.bias is not None:
        if parallel == "column":
            n0, n1 = shard_bounds(int(meta["shape"][0]), tp, rank)
            new.bias = layer.bias[n0:n1].clone()
        else:
            new.bias = layer.bias.clone() if rank == 0 else None
    new.encoded_state_dict = layer.encoded_state_dict
    new.in_features, new.out_features = meta_s["shape"][::-1]
    new.in_gpu, new.ready = True, True
    return new
