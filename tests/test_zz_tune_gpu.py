"""GPU: DecodeModel.retune / hqq_b200.tune on the default kernels (the candidates themselves are covered, opt-in, by
tests/test_zz_variants_gpu.py).  Written after round 1's GPU budget was spent: non-strict xfail until seen green, and last in the
collection order so that nothing here can affect the validated suite."""
import pytest
import torch

from hqq_b200 import harness

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


@pytest.mark.xfail(strict=False, reason="written after round 1's GPU budget was spent; not yet seen on a GPU")
def test_retune_recaptures_the_same_token_stream():
    """DecodeModel.retune({}) (env reset + hqq_b200_reload_env + new capture) leaves the default kernels producing the same tokens;
    tune.measure / tune.choose_decode run on it and, with no guard survivors, keep the default configuration."""
    from hqq_b200 import tune
    shape = harness.LlamaShape(hidden=1024, inter=2048, n_layers=2, n_heads=8, n_kv_heads=2, vocab=2048)
    m = harness.DecodeModel(shape, dtype=torch.float16, device=DEV, cache_len=128, seed=3)
    m.capture()
    t0, us0 = tune.measure(m, steps=10)
    m.retune({})
    t1, us1 = tune.measure(m, steps=10)
    assert torch.equal(t0, t1) and us0 > 0 and us1 > 0
    rep = tune.choose_decode(m, [{"knobs": {}, "us": us0, "digest": tune.token_digest(t0)}], steps=10)
    assert rep["selected"] == {} and rep["gain"] == 1.0
    t2, _ = tune.measure(m, steps=10)
    assert torch.equal(t0, t2)
    with pytest.raises(ValueError):
        m.retune({"HQQ_B200_PDL": "0"})  # not a result-preserving decode knob
