"""GPU parity at the large end of BASELINE cfg 5's block-size sweep (4 MiB and 16 MiB blocks): oracle parity for the BBWT
and size-independent round trips through every stage."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _ctx():
    import gpu_util as G
    return G.ctx(max_bytes=40 << 20, max_blocks=64)


def test_bbwt_4mib_block_matches_oracle():
    import torch
    from kolmogorovlike_datacompressor_b200 import synth
    data = synth.s3_mix(8 << 20)[2 << 20:6 << 20].copy()          # S2 gradient + sine + pattern + checker segments: long repeats
    off = np.array([0, data.size], dtype=np.int64)
    t = torch.from_numpy(data).cuda()
    L = _ctx().bbwt_forward(t, off)
    assert L[:data.size].cpu().numpy().tobytes() == O.bbwt_forward(data.tobytes())


def test_roundtrip_16mib_block_and_ragged_neighbours():
    import torch
    from kolmogorovlike_datacompressor_b200 import synth
    big = np.concatenate([synth.s1_text(12 << 20), synth.s2_mixed(4 << 20)])
    small = synth.s1_text(70000, seed=5)
    data = np.concatenate([small[:1], big, small])
    off = np.array([0, 1, 1 + big.size, data.size], dtype=np.int64)
    n = data.size
    t = torch.from_numpy(data).cuda()
    c = _ctx()
    L = c.bbwt_forward(t, off)
    m = c.mtf_encode(L, off)
    pay, poff = c.rice_kf_encode(m, off)
    m2 = c.rice_kf_decode(pay, poff, off)
    assert torch.equal(m2[:n], m[:n])
    L2 = c.mtf_decode(m2, off)
    assert torch.equal(L2[:n], L[:n])
    x = c.bbwt_inverse(L2, off)
    assert torch.equal(x[:n], t[:n])
    # the small neighbour agrees with the oracle bit for bit
    assert L[1 + big.size:n].cpu().numpy().tobytes() == O.bbwt_forward(small.tobytes())
    # LZ77 (V22 parameters) round trip on the same ragged batch
    lz, lzo = c.lz77_encode(t, off, 4096, 0)
    back = c.lz77_decode(lz, lzo, off, 4096)
    assert torch.equal(back[:n], t[:n])
