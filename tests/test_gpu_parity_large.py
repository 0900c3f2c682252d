"""Byte parity with the oracle AT the bench configuration's block sizes: MTF, the KF model-2 payload
with its parameters, all five V22 Rice payloads and sizes on 1 MiB blocks of every corpus kind (the per-block scans split into
groups at this size), the same chain on one 16 MiB block, and a KOLR container at 1 MiB blocks whose per-block (method, payload)
equals the oracle's selection over all ten candidates."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

K2_FLAGS = (0, 1, 4, 8, 16)


def _chain_vs_oracle(c, data, off, blocks_to_check):
    import torch
    n = int(off[-1])
    t = torch.from_numpy(data).cuda()
    L = c.bbwt_forward(t, off)
    m = c.mtf_encode(L, off)
    kf, kfo, prm = c.rice_kf_encode(m, off, want_params=True)
    g_L, g_m = L[:n].cpu().numpy().tobytes(), m[:n].cpu().numpy().tobytes()
    g_kf = kf[:int(kfo[-1])].cpu().numpy().tobytes()
    k2 = {}
    for fl in K2_FLAGS:
        p, o, sz = c.rice_k2_encode(m, off, fl)
        k2[fl] = (p[:int(o[-1])].cpu().numpy().tobytes(), o, sz)
    for b in blocks_to_check:
        a, e = int(off[b]), int(off[b + 1])
        blk = data[a:e].tobytes()
        oL = O.bbwt_forward(blk)
        om = O.mtf_encode(oL)
        assert g_L[a:e] == oL, ("bbwt", b)
        assert g_m[a:e] == om, ("mtf", b)
        okf, oprm = O.kf_rice_pack(om, with_params=True)
        assert g_kf[kfo[b]:kfo[b + 1]] == okf, ("kf payload", b)
        assert [int(v) for v in prm[b]] == [oprm["k0"], oprm["k1"], int(oprm["use_rice_zero"]), int(oprm["use_rice_nz"])], ("kf params", b)
        for i, fl in enumerate(K2_FLAGS):
            want = O.v22_rice_pack(om, fl)
            pay, o, sz = k2[fl]
            assert pay[o[b]:o[b + 1]] == want, ("k2 payload", fl, b)
            assert int(sz[b, i]) == len(want), ("k2 size", fl, b)
    # round trip of the whole batch through the decoders
    m2 = c.rice_kf_decode(kf, kfo, off)
    assert torch.equal(m2[:n], m[:n])
    x = c.bbwt_inverse(c.mtf_decode(m2, off), off)
    assert torch.equal(x[:n], t[:n])


def test_mtf_kf_k2_payloads_1mib_blocks_every_corpus_kind():
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200 import synth
    mib = 1 << 20
    s2 = synth.s2_mixed(4 * mib)
    rnd = np.random.Generator(np.random.PCG64(11)).integers(0, 256, size=mib, dtype=np.uint8)
    ragged = synth.s1_text(mib + 4097, seed=9)                      # not a multiple of the tile, the piece or the scan group
    data = np.concatenate([synth.s1_text(mib), s2, rnd, ragged])
    off = np.array([0, mib, 2 * mib, 3 * mib, 4 * mib, 5 * mib, 6 * mib, 6 * mib + ragged.size], dtype=np.int64)
    _chain_vs_oracle(G.ctx(max_bytes=16 << 20, max_blocks=64), data, off, range(7))


def test_chain_one_16mib_block():
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200 import synth
    data = synth.s3_mix(16 << 20)                                   # text, gradient, sine, pattern, checker, random megabytes in ONE block
    off = np.array([0, data.size], dtype=np.int64)
    _chain_vs_oracle(G.ctx(max_bytes=40 << 20, max_blocks=64), data, off, [0])


def test_kolr_container_1mib_blocks_equals_oracle_selection():
    """compress_blocks_fixed(data, 1 MiB) on one megabyte of every S3 segment kind: each block's (method, payload) in the container
    equals the oracle's selection with all ten candidates (LZ77 and Re-Pair through the oracle's fast exact forms)."""
    from concurrent.futures import ThreadPoolExecutor
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    from kolmogorovlike_datacompressor_b200 import synth
    mib = 1 << 20
    data = synth.s3_mix(8 * mib).tobytes()
    with ThreadPoolExecutor(max_workers=8) as pool:
        futs = [pool.submit(O.encode_block, O.PROFILE_KOLR, data[b * mib:(b + 1) * mib], None, True) for b in range(8)]
        blob = V.compress_blocks_fixed(data, mib)
        names, starts, plens, olens, total, _ = V._parse(blob)
        assert total == len(data) and len(names) == 8
        for b, f in enumerate(futs):
            mid, payload, sizes = f.result()
            assert names[b] == V.KOLR_NAMES[mid], (b, names[b], sizes)
            assert blob[starts[b]:starts[b] + plens[b]] == payload, b
    assert V.decompress(blob) == data


def test_kolm_container_1mib_target_equals_oracle_selection():
    """kolm_final.compress(data, target_block = 1 MiB): CDC boundaries and every block's (method, payload) vs the oracle."""
    from concurrent.futures import ThreadPoolExecutor
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    from kolmogorovlike_datacompressor_b200 import synth
    mib = 1 << 20
    data = synth.s3_mix(6 * mib).tobytes()
    bounds = O.kf_cdc(data, mib // 2, mib, 2 * mib)
    ids = {"raw": 0, "kf_xor": 1, "kf_bbwt": 2, "kf_lz77": 3}
    with ThreadPoolExecutor(max_workers=8) as pool:
        futs = [pool.submit(O.encode_block, O.PROFILE_KOLM, data[a:b], None, True) for a, b in bounds]
        blob = KF.compress(data, mib)
        names, starts, plens, olens, total = KF._parse(blob)
        assert [int(x) for x in olens] == [b - a for a, b in bounds]
        for i, f in enumerate(futs):
            mid, payload, _ = f.result()
            assert ids[names[i]] == mid, i
            assert blob[starts[i]:starts[i] + plens[i]] == payload, i
    assert KF.decompress(blob) == data


def test_chain_1mib_blocks_of_the_s3_mix_every_segment_kind():
    """The S3 mix (cfg 4 / cfg 5 data) at 1 MiB blocks, one block of each of its eight segment kinds: long repeats (fills, ramps,
    checker rows) keep the rotation sort in its rounds for 15-20 doublings with groups far above the local kernel's limit, so
    k_refine_local, the big-group path, the hand-over to the compacted rounds and the stable-partition exit all run."""
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200 import synth
    mib = 1 << 20
    data = synth.s3_mix(8 * mib)
    off = np.arange(0, 8 * mib + 1, mib, dtype=np.int64)
    _chain_vs_oracle(G.ctx(max_bytes=16 << 20, max_blocks=64), data, off, range(8))
