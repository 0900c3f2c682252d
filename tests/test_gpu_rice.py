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


def _stress_cases():
    """Shapes that exercise the warp-per-tile kernels' corners: runs crossing step (512 B), tile (4096 B) and several-tile boundaries,
    runs ending exactly ON a boundary, blocks ending in zeros / in a non-zero, values whose tokens exceed 32 bits (unary parts of Rice
    with a small k), windows that overflow the staging area (all 255s), unaligned block starts (odd lengths in front)."""
    rnd = random.Random(17)
    c = {}
    c["odd_front_3"] = b"\x01\x00\x02"                       # makes every later block start unaligned
    for n in (1, 15, 16, 17, 511, 512, 513, 4095, 4096, 4097, 8191, 8192, 8193, 12288 + 5):
        c["zeros_%d" % n] = bytes(n)
        c["zeros_then_one_%d" % n] = bytes(n) + b"\x01"
        c["one_then_zeros_%d" % n] = b"\x07" + bytes(n)
        c["nz_%d" % n] = bytes(rnd.randrange(1, 256) for _ in range(n))
    c["run_ends_on_tile"] = bytes(4096) + b"\x05" + bytes(4095) + b"\x06" + bytes(4096 * 3 - 1) + b"\x01"
    c["run_ends_before_tile"] = b"\x03" + bytes(4094) + b"\x05\x09" + bytes(8190) + b"\x01\x00\x00"
    c["all_255_20000"] = b"\xff" * 20000                     # 66-bit codes: the staging window's worst case
    c["all_254_5000"] = b"\xfe" * 5000
    c["big_then_sparse"] = bytes([255, 0, 254, 0, 0, 253]) * 300 + bytes(9000) + bytes(rnd.choice((0, 0, 0, 200)) for _ in range(9000))
    c["mostly_zero_rare_big"] = bytes((255 if rnd.random() < 0.003 else 0) for _ in range(70000))   # Rice k small vs huge x: long unary parts
    c["ones_and_twos"] = bytes(rnd.choice((1, 2)) for _ in range(10000))
    c["alternating"] = bytes([0, 1]) * 6000 + bytes([0, 0, 0, 3]) * 3000
    c["long_run_1m"] = bytes(1 << 20) + b"\x01" + bytes(300000)
    c["geometric"] = bytes(min(255, int(rnd.expovariate(0.35))) for _ in range(50000))
    c["odd_7"] = bytes(rnd.randrange(3) for _ in range(7))
    return c


def test_rice_stress_shapes_kf_k2_and_dual():
    import gpu_util as G
    cases = _stress_cases()
    names = list(cases)                                      # insertion order: the odd-sized block stays in front
    blocks = [cases[k] for k in names]
    t, off = G.batch(blocks)
    c = G.ctx()
    out, out_off, params = c.rice_kf_encode(t, off, want_params=True)
    got = out[:int(out_off[-1])].cpu().numpy().tobytes()
    want_kf = []
    for i, k in enumerate(names):
        want, prm = O.kf_rice_pack(blocks[i], with_params=True)
        want_kf.append(want)
        assert list(params[i]) == [prm["k0"], prm["k1"], int(prm["use_rice_zero"]), int(prm["use_rice_nz"])], k
        assert got[out_off[i]:out_off[i + 1]] == want, k
    assert G.unbatch(c.rice_kf_decode(out, out_off, off), off) == blocks
    want_k2 = {f: [O.v22_rice_pack(b, f) for b in blocks] for f in FLAGS}
    for flags in FLAGS:
        o2, oo2, sizes = c.rice_k2_encode(t, off, flags)
        g2 = o2[:int(oo2[-1])].cpu().numpy().tobytes()
        for i, k in enumerate(names):
            assert g2[oo2[i]:oo2[i + 1]] == want_k2[flags][i], (k, flags)
            assert list(sizes[i]) == [len(want_k2[f][i]) for f in FLAGS], k
    # both coders from one cost read
    for flags in (0, 1, 16):
        kf, kfo, prm2, k2, k2o, sz2 = c.rice_dual_encode(t, off, flags)
        assert np.array_equal(kfo, out_off) and np.array_equal(prm2, params)
        assert kf[:int(kfo[-1])].cpu().numpy().tobytes() == got
        g2 = k2[:int(k2o[-1])].cpu().numpy().tobytes()
        for i, k in enumerate(names):
            assert g2[k2o[i]:k2o[i + 1]] == want_k2[flags][i], (k, flags)
            assert list(sz2[i]) == [len(want_k2[f][i]) for f in FLAGS], k


def test_rice_capacity_guard_writes_nothing():
    """An undersized output buffer: KOLM_E_CAPACITY comes back and not one byte past (or inside) the buffer was written."""
    import gpu_util as G
    import torch
    from kolmogorovlike_datacompressor_b200 import _lib
    rnd = random.Random(4)
    blk = bytes(rnd.randrange(256) for _ in range(30000))
    t, off = G.batch([blk])
    c = G.ctx()
    for name, enc in (("kf", lambda o: c.rice_kf_encode(t, off, out=o)), ("k2", lambda o: c.rice_k2_encode(t, off, 0, out=o)),
                      ("lz77", lambda o: c.lz77_encode(t, off, 4096, 0, out=o)), ("residual", lambda o: c.residual_encode(t, off, 1, out=o)),
                      ("repair", lambda o: c.repair_encode(t, off, out=o))):
        guard = torch.full((4096 + 256,), 0xAB, dtype=torch.uint8, device="cuda")
        code = 0
        try:
            enc(guard[:1024])
        except _lib.KolmError as e:
            code = e.code
        assert code == -3, name
        torch.cuda.synchronize()
        assert bool((guard[1024:] == 0xAB).all()), name
