"""GPU parity: KF model-2 token coder and V22 Rice(k=2) variants vs the CPU oracle (bit-exact)."""
import random

import numpy as np
import pytest

import datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu
FLAGS = [0, 1, 4, 8, 16]


def _mtf_cases():
    c = {}
    for k, d in {**datasets.small_cases(), **datasets.medium_cases()}.items():
        c[k] = O.mtf_encode(O.bbwt_forward(d))
    rnd = random.Random(3)
    c["allzero_9000"] = bytes(9000)
    c["onezero"] = b"\0"
    c["one5"] = b"\x05"
    c["zeros_then_big"] = bytes(5000) + b"\xff" + bytes(4097) + b"\x01\x00"
    c["rand_20000"] = bytes(rnd.randrange(256) for _ in range(20000))
    c["sparse"] = bytes((rnd.randrange(1, 4) if rnd.random() < 0.02 else 0) for _ in range(30000))
    c["long_run_mix"] = bytes(40000) + bytes(rnd.randrange(1, 3) for _ in range(300)) + bytes(70000)
    for name in ("checker", "sine", "pattern", "gradient"):
        c["fx_" + name] = O.mtf_encode(O.bbwt_forward(datasets.fixture(name)[:50000]))
    return c


def test_kf_model2_pack():
    import gpu_util as G
    cases = _mtf_cases()
    names = sorted(cases)
    blocks = [cases[k] for k in names]
    t, off = G.batch(blocks)
    out, out_off, params = G.ctx().rice_kf_encode(t, off, want_params=True)
    got = out.cpu().numpy().tobytes()
    for i, k in enumerate(names):
        want, prm = O.kf_rice_pack(blocks[i], with_params=True)
        assert list(params[i]) == [prm["k0"], prm["k1"], int(prm["use_rice_zero"]), int(prm["use_rice_nz"])], k
        assert got[out_off[i]:out_off[i + 1]] == want, k


@pytest.mark.parametrize("flags", FLAGS)
def test_v22_rice_k2_pack(flags):
    import gpu_util as G
    cases = _mtf_cases()
    names = sorted(cases)
    blocks = [cases[k] for k in names]
    t, off = G.batch(blocks)
    out, out_off, sizes = G.ctx().rice_k2_encode(t, off, flags)
    got = out.cpu().numpy().tobytes()
    for i, k in enumerate(names):
        want = O.v22_rice_pack(blocks[i], flags)
        assert got[out_off[i]:out_off[i + 1]] == want, (k, flags)
        assert list(sizes[i]) == [len(O.v22_rice_pack(blocks[i], f)) for f in FLAGS], k
