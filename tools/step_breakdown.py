"""Where does a decode step's time go?  Captures variants of the 5-launch block into CUDA graphs and times them."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import harness, ops
from hqq_b200._lib import DTYPE_CODE, check, load, ptr, stream_ptr

dev = torch.device("cuda", 0)
model = harness.DecodeModel(harness.LLAMA3_8B, nbits=4, group_size=64, dtype=torch.float16, device=dev, cache_len=256, n_layers=32)
model.capture(warmup=2)
lib, s, b = load(), model.shape, model._bufs
code = DTYPE_CODE[model.dtype]
hd, hq, hkv = s.head_dim, s.n_heads, s.n_kv_heads


def variant(qkv=True, attn=True, o=True, gu=True, down=True, prolog=True, head=False, pair=False, hint=False):
    st = stream_ptr(dev)
    h_cur, h_nxt = b["h"], b["h2"]
    delta = None
    for blk in model.blocks:
        if qkv:
            if prolog:
                ops.decode_linear_fwd(h_cur, (blk["q"], blk["k"], blk["v"]), [b["q"], b["k"], b["v"]], 1, delta, blk["norm1"], h_nxt, s.rms_eps,
                                      l2_hint=(model._kv_hint(blk) if hint else None))
                h_cur, h_nxt = h_nxt, h_cur
            else:
                ops.decode_linear_fwd(h_cur, (blk["q"], blk["k"], blk["v"]), [b["q"], b["k"], b["v"]])
        if attn:
            check(lib.hqq_b200_glue_rope_attn_decode(ptr(b["q"]), ptr(b["k"]), ptr(b["v"]), ptr(model.cos), ptr(model.sin), ptr(blk["k_cache"]),
                                                     ptr(blk["v_cache"]), ptr(model.pos), ptr(b["a"]), hq, hkv, model.cache_len, hd, code, st))
        if o:
            ops.decode_linear_fwd(b["a"], (blk["o"],), [b["o"]])
        if gu and pair:
            ops.decode_linear_fwd(h_cur, (blk["gate"], blk["up"]), [b["act"], b["up"]], 1 | ops.YOP_SILU_MUL_PAIR, b["o"], blk["norm2"], h_nxt, s.rms_eps)
            h_cur, h_nxt = h_nxt, h_cur
        elif gu:
            if prolog:
                ops.decode_linear_fwd(h_cur, (blk["gate"], blk["up"]), [b["gate"], b["up"]], 1, b["o"], blk["norm2"], h_nxt, s.rms_eps)
                h_cur, h_nxt = h_nxt, h_cur
            else:
                ops.decode_linear_fwd(h_cur, (blk["gate"], blk["up"]), [b["gate"], b["up"]])
        if down and pair:
            ops.decode_linear_fwd(b["act"], (blk["down"],), [b["down"]])
        elif down:
            if prolog:
                ops.decode_linear_fwd(b["gate"], (blk["down"],), [b["down"]], 2, b["up"])
            else:
                ops.decode_linear_fwd(b["gate"], (blk["down"],), [b["down"]])
        delta = b["down"]
    if head:
        check(lib.hqq_b200_glue_add_rmsnorm(ptr(h_cur), ptr(delta), ptr(model.final_norm), ptr(b["x"]), s.hidden, s.rms_eps, code, st))
        torch.matmul(b["x"], model.lm_head.t(), out=b["logits"])
        check(lib.hqq_b200_glue_argmax(ptr(b["logits"]), s.vocab, ptr(model.next_tok), code, st))


def time_variant(name, **kw):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side), torch.no_grad():
        variant(**kw)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        variant(**kw)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name:44s} {us:9.1f} us/step  {us / 32:7.2f} us/block", flush=True)


model.pos.fill_(128)
time_variant("full step, unpaired silu", head=True, pair=False)
time_variant("full step, paired silu", head=True, pair=True)
time_variant("full step, paired silu + kv L2 hint", head=True, pair=True, hint=True)
time_variant("qkv+attn + kv L2 hint", o=False, gu=False, down=False, hint=True)
time_variant("blocks only")
time_variant("gate_up paired only", qkv=False, attn=False, o=False, down=False, pair=True)
time_variant("gate_up+down paired", qkv=False, attn=False, o=False, pair=True)
time_variant("blocks, no attention", attn=False)
time_variant("blocks, no prologues (x_op=0)", prolog=False)
time_variant("linears only, no prologues", attn=False, prolog=False)
time_variant("attention only", qkv=False, o=False, gu=False, down=False)
time_variant("qkv only (prolog)", attn=False, o=False, gu=False, down=False)
time_variant("qkv only (x_op 0)", attn=False, o=False, gu=False, down=False, prolog=False)
time_variant("o only", qkv=False, attn=False, gu=False, down=False)
time_variant("gate_up only (prolog)", qkv=False, attn=False, o=False, down=False)
time_variant("gate_up only (x_op 0)", qkv=False, attn=False, o=False, down=False, prolog=False)
time_variant("down only (prolog)", qkv=False, attn=False, o=False, gu=False)
time_variant("down only (x_op 0)", qkv=False, attn=False, o=False, gu=False, prolog=False)
time_variant("qkv+attn", o=False, gu=False, down=False)
time_variant("gate_up+down", qkv=False, attn=False, o=False)
time_variant("head only", qkv=False, attn=False, o=False, gu=False, down=False, head=True)
