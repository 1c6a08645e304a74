"""CPU, world_size 2, gloo: the multi-GPU host logic of the decode harness / bench -- shard shapes,
the column-/row-parallel algebra with its single all-reduce, and the max-over-ranks timing reduction bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hqq_b200 import harness


def test_shard_dims():
    d = harness.shard_dims(harness.LLAMA3_8B, 1)
    assert d["q"] == (4096, 4096) and d["k"] == (1024, 4096) and d["gate"] == (14336, 4096) and d["down"] == (4096, 14336)
    d8 = harness.shard_dims(harness.LLAMA3_8B, 8)
    assert d8["q"] == (512, 4096) and d8["k"] == (128, 4096) and d8["o"] == (4096, 512) and d8["down"] == (4096, 1792)
    # every K handed to a row-parallel shard stays a multiple of the group size and of the 256-k unit of the fused kernel
    for tp in (1, 2, 4, 8):
        for name, (n, k) in harness.shard_dims(harness.LLAMA3_8B, tp).items():
            assert k % 64 == 0 and n % 8 == 0, (tp, name)
    d70 = harness.shard_dims(harness.LLAMA3_70B, 8)
    assert d70["down"] == (8192, 3584) and d70["gate"] == (3584, 8192)
    with pytest.raises(ValueError):
        harness.shard_dims(harness.LLAMA3_8B, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical "unsharded" tensors on every rank
    H, I, = 64, 256
    x = torch.randn(1, H)
    Wg, Wu, Wd = torch.randn(I, H), torch.randn(I, H), torch.randn(H, I)
    full = (torch.nn.functional.silu(x @ Wg.t()) * (x @ Wu.t())) @ Wd.t()
    sl = slice(rank * I // world, (rank + 1) * I // world)
    # column-parallel gate/up: no traffic; row-parallel down: ONE all-reduce of the [1, hidden] partial
    part = (torch.nn.functional.silu(x @ Wg[sl].t()) * (x @ Wu[sl].t())) @ Wd[:, sl].t()
    dist.all_reduce(part)
    ok = torch.allclose(part, full, rtol=1e-4, atol=1e-4)
    # bench.py: time = max over ranks
    t = torch.tensor([10.0 + rank, 5.0 - rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.tolist() == [10.0 + world - 1, 5.0]
    dist.barrier()
    if rank == 0:
        out.put(ok)
    dist.destroy_process_group()


def test_row_parallel_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_layer_sharding_plan_is_complete_balanced_and_deterministic():
    """Quantise-only sharding (BASELINE configs[3]): every rank derives the same size-balanced plan, no collective."""
    from hqq_b200 import harness
    dims = harness.shard_dims(harness.LLAMA3_70B, 1)
    sizes = [dims[n][0] * dims[n][1] for _ in range(80) for n in ("q", "k", "v", "o", "gate", "up", "down")]
    for world in (1, 2, 4, 8):
        plan = harness.assign_layers(sizes, world)
        assert plan == harness.assign_layers(list(sizes), world)
        assert sorted(i for p in plan for i in p) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in plan]
        assert max(loads) - min(loads) <= max(sizes)
    assert harness.assign_layers([5, 1, 1], 2) == [[0], [1, 2]]


def test_batched_decode_step_equals_independent_sequences(monkeypatch):
    """BASELINE configs[4] bs = 32 leg: DecodeModel(batch=B).step() decodes B sequences in lock-step.  Host glue only (KV cache
    per sequence, RoPE at the shared position, GQA attention, per-sequence argmax), checked on CPU with dense stand-ins for the
    HQQLinear layers: every sequence must produce the tokens it produces alone."""
    class Dense:
        def __init__(self, W):
            self.W = W

        def __call__(self, x):
            return x @ self.W.t()

    monkeypatch.setattr(harness.HQQLinear, "from_weights", staticmethod(lambda W, bias, cfg, compute_dtype=None, device=None: Dense(W)))
    monkeypatch.setattr(harness.ops, "linear_fwd_multi", lambda x, layers, outs=None: None)
    shape = harness.LlamaShape(hidden=64, inter=128, n_layers=2, n_heads=4, n_kv_heads=2, vocab=97)
    toks = [5, 9, 11]

    def run(batch, first):
        m = harness.DecodeModel(shape, dtype=torch.float64, device="cpu", cache_len=16, fused=False, seed=1, batch=batch)
        assert m.fused is False and m.blocks[0]["k_cache"].shape[0] == batch
        m.tok.copy_(torch.tensor(first))
        out = []
        for _ in range(5):
            m.step()
            out.append(m.next_tok.clone())
            m.tok.copy_(m.next_tok)
        return torch.stack(out, 1)  # [batch, steps]

    together = run(3, toks)
    for i, t in enumerate(toks):
        alone = run(1, [t])
        assert torch.equal(together[i], alone[0]), (i, together[i], alone[0])
    assert len({tuple(r.tolist()) for r in together}) > 1  # the sequences really differ
    with pytest.raises(ValueError):
        harness.DecodeModel(shape, dtype=torch.float64, device="cpu", cache_len=16, fused=False, batch=0)


class _TwoModeModel:
    """Stand-in for harness.DecodeModel in bench.tokens_agree: a vocabulary of 16 sharded over 2 ranks; the "nccl" mode's logits are
    the "p2p" mode's plus `pert` (what a different summation order does)."""

    def __init__(self, rank, pert):
        self.rank, self.vocab_shard, self.device = rank, 8, torch.device("cpu")
        self.tp_mode, self.graph = "p2p", object()
        self.tok = torch.zeros(1, dtype=torch.long)
        self.next_tok = torch.zeros(1, dtype=torch.long)
        self._bufs = {"logits": torch.zeros(1, 8, dtype=torch.float16)}
        g = torch.Generator().manual_seed(3)
        self.table = (torch.randn(64, 16, generator=g) * 2).half().float()
        self.table[5, 3] = self.table[5, 11] = 9.0   # a dead heat between a token of rank 0's shard and one of rank 1's
        self.table[5, 11] -= 2 ** -7                 # ... that "p2p" decides for token 3 by one fp16 step
        self.pert, self.captures, self.i = pert, 0, 0

    def capture(self, warmup=2):
        self.graph = object()
        self.captures += 1

    def reset_state(self, token=1):
        self.tok.fill_(token)
        self.i = 0

    def decode(self, feed_back=True):
        row = 5 if self.i == 4 else (int(self.tok) * 7 + self.i) % 64
        g = (self.table[row] + (self.pert if self.tp_mode == "nccl" else 0.0)).half().float()
        self._bufs["logits"].copy_(g[self.rank * 8:(self.rank + 1) * 8].view(1, 8))
        self.next_tok.fill_(int(torch.argmax(g)))
        self.i += 1
        if feed_back:
            self.tok.copy_(self.next_tok)


def _agree_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    torch.cuda.synchronize = lambda *a, **k: None  # the stand-in lives on the CPU
    res = {}
    zero = torch.zeros(16)
    tie = torch.zeros(16); tie[11] = 2 ** -6       # last-bits difference: flips the dead heat at step 4, nothing else
    wrong = torch.zeros(16); wrong[2] = 30.0       # a broken exchange: one rank's contribution is off by a lot
    for name, pert in (("same", zero), ("tie", tie), ("wrong", wrong)):
        m = _TwoModeModel(rank, pert)
        agree, detail = bench.tokens_agree(m, torch, n_tokens=8)
        res[name] = (agree, detail, m.tp_mode, m.captures)
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_p2p_vs_nccl_check_accepts_rounding_ties_and_nothing_else():
    """bench.tokens_agree on two gloo ranks with a stand-in model: identical modes agree; a pick that differs because two logits of
    different vocabulary shards are one fp16 step apart is accepted and reported; a gross logit difference is not.  The model is
    left in the mode it came with."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_agree_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    agree, d, mode, captures = res["same"]
    assert agree and d["identical"] and d["max_logit_diff"] == 0.0 and mode == "p2p" and captures == 2 and "differing_picks" not in d
    agree, d, mode, _ = res["tie"]
    assert agree and not d["identical"] and mode == "p2p"
    assert [p["step"] for p in d["differing_picks"]] == [4] and d["differing_picks"][0]["near_tie"]
    assert 0 < d["differing_picks"][0]["margin"] <= 2 * d["differing_picks"][0]["max_logit_diff"] <= 2 ** -5
    agree, d, mode, _ = res["wrong"]
    assert not agree and d["max_logit_diff"] >= 29.0 and mode == "p2p"
