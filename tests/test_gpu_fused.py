"""kolm_encode_blocks / kolm_decode_blocks (SURVEY 8b "fused hot path — what compress() calls") against the oracle's per-block
selection: every block's (method id, payload) and the whole size table, for both profiles, with ragged batches (an empty block, a
one-byte block, unaligned starts), candidate masks, an externally computed candidate, and the decode loop with bad blocks."""
import numpy as np
import pytest

import datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu
BIG = (1 << 62) - 1


def _blocks():
    c = datasets.small_cases()
    names = ["text", "empty", "one", "random_bytes", "repetitive_text", "checker_0", "gradient_200000", "pattern_393316", "sine_1000", "zero_2k",
             "byte_counter", "random_small_alpha", "runs18", "fib", "utf8_mixed", "banana", "desc"]
    blocks = [c[k] for k in names]
    m = datasets.medium_cases()
    blocks += [m["text_big"][:7001], m["sine_20k"][:9000], m["pattern_mix_24k"][:8192], datasets.fixture("checker")[:8191]]
    return blocks


@pytest.mark.parametrize("profile", [1, 2])
def test_encode_blocks_equals_oracle_selection(profile):
    import gpu_util as G
    blocks = _blocks()
    t, off = G.batch(blocks)
    c = G.ctx()
    out, poff, mids, sizes = c.encode_blocks(profile, t, off, want_sizes=True)
    got = out[:int(poff[-1])].cpu().numpy().tobytes()
    for b, blk in enumerate(blocks):
        mid, payload, osz = O.encode_block(profile, blk)
        assert [int(v) for v in sizes[b]] == [int(v) for v in osz], (b, len(blk))
        assert int(mids[b]) == mid, (b, len(blk), osz)
        assert got[poff[b]:poff[b + 1]] == payload, (b, mid)
    # the decode loop brings every block back
    ooff = np.zeros(len(blocks) + 1, dtype=np.int64)
    ooff[1:] = np.cumsum([len(b) for b in blocks])
    back = c.decode_blocks(profile, out, poff[:-1], np.diff(poff), mids, ooff)
    assert back[:int(ooff[-1])].cpu().numpy().tobytes() == b"".join(blocks)


def test_encode_blocks_masks_and_external_candidate():
    import gpu_util as G
    blocks = [b for b in _blocks() if b]
    t, off = G.batch(blocks)
    c = G.ctx()
    # without lz77 (bit 7) and repair (bit 9): the oracle with the same mask
    mask = 0x3FF & ~(1 << 7) & ~(1 << 9)
    out, poff, mids, sizes = c.encode_blocks(2, t, off, cand_mask=mask, want_sizes=True)
    got = out[:int(poff[-1])].cpu().numpy().tobytes()
    for b, blk in enumerate(blocks):
        mid, payload, osz = O.encode_block(2, blk, models_mask=mask)
        assert int(mids[b]) == mid and got[poff[b]:poff[b + 1]] == payload, b
        assert int(sizes[b][7]) == BIG and int(sizes[b][9]) == BIG
    # Re-Pair computed by a separate call and handed in as the external candidate: same container bytes as the all-inline call
    rp, rpo = c.repair_encode(t, off)
    rp = rp.clone()
    addr = np.uint64(rp.data_ptr()) + rpo[:-1].astype(np.uint64)
    nolz = 0x3FF & ~(1 << 7)                                # with LZ77 out of the way Re-Pair wins the repetitive blocks
    full = c.encode_blocks(2, t, off, cand_mask=nolz)
    ext = c.encode_blocks(2, t, off, cand_mask=nolz, ext=(9, np.diff(rpo), addr))
    assert np.array_equal(full[1], ext[1]) and np.array_equal(full[2], ext[2])
    n = int(full[1][-1])
    assert full[0][:n].cpu().numpy().tobytes() == ext[0][:n].cpu().numpy().tobytes()
    assert (np.asarray(full[2]) == 9).any()                  # Re-Pair does win some of these blocks


def test_encode_blocks_capacity_and_arguments():
    import gpu_util as G
    import torch
    from kolmogorovlike_datacompressor_b200 import _lib
    blocks = [b for b in _blocks() if b]
    t, off = G.batch(blocks)
    c = G.ctx()
    guard = torch.full((2048 + 256,), 0xAB, dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.KolmError) as e:
        c.encode_blocks(2, t, off, out=guard[:2048])
    assert e.value.code == -3
    torch.cuda.synchronize()
    assert bool((guard[2048:] == 0xAB).all())
    with pytest.raises(_lib.KolmError) as e:
        c.encode_blocks(2, t, off + 1)                        # off[0] must be 0
    assert e.value.code == -2


def test_decode_blocks_reports_the_first_bad_block():
    import gpu_util as G
    import torch
    from kolmogorovlike_datacompressor_b200 import _lib
    blocks = [b for b in _blocks() if b][:12]
    t, off = G.batch(blocks)
    c = G.ctx()
    out, poff, mids = c.encode_blocks(2, t, off)
    n = int(poff[-1])
    ooff = np.zeros(len(blocks) + 1, dtype=np.int64)
    ooff[1:] = np.cumsum([len(b) for b in blocks])
    # truncate the payloads of blocks 5 and 8 by declaring them shorter: the lowest index comes back
    plen = np.diff(poff).copy()
    victims = [b for b in (5, 8) if int(mids[b]) != 0]
    for b in victims:
        plen[b] = max(0, plen[b] - max(1, plen[b] // 2))
    if victims:
        with pytest.raises(_lib.KolmError) as e:
            c.decode_blocks(2, out, poff[:-1], plen, mids, ooff)
        assert e.value.block == victims[0] and e.value.code in (-4, -5, -7)
    bad = np.array(mids).copy()
    bad[3] = 200                                             # unknown method id: ValueError in the reference
    with pytest.raises(_lib.KolmError) as e:
        c.decode_blocks(2, out, poff[:-1], np.diff(poff), bad, ooff)
    assert e.value.code == -5 and e.value.block == 3
    torch.cuda.synchronize()


def test_dropins_fused_and_stagewise_paths_agree():
    """The drop-ins produce the same container through kolm_encode_blocks (default) and through the stage-by-stage path."""
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    data = datasets.medium_cases()["pattern_mix_24k"] + datasets.medium_cases()["text_big"] + datasets.fixture("sine")[:30000]
    res = {}
    for fused in (True, False):
        for mod in (KF, V):
            eng = mod._engine()
            old = eng.fused
            eng.fused = fused
            try:
                blob = mod.compress(data, 4096) if mod is KF else mod.compress_blocks_fixed(data, 2048)
                assert mod.decompress(blob) == data
                res[(mod.__name__, fused)] = blob
            finally:
                eng.fused = old
    assert res[(KF.__name__, True)] == res[(KF.__name__, False)] == O.kf_compress(data, 4096)
    assert res[(V.__name__, True)] == res[(V.__name__, False)]


def test_repair_early_stop_leaves_the_selection_unchanged():
    """kolm_encode_blocks without the size table stops the Re-Pair rounds of a block once a lower bound on the final payload
    reaches the best other candidate (csrc/repair.cu; tests/test_repair_bound.py checks the bound itself): method ids, offsets and
    payload bytes equal the call that runs every candidate to the end, and the oracle's selection; many blocks do stop early."""
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200 import synth
    mix = synth.s3_mix(8 << 20)
    blocks = [b for b in _blocks() if b]
    for kind in range(8):                                     # 2 KiB (the reference's default block size) ... 8 KiB blocks of every segment kind
        seg = mix[kind << 20:(kind + 1) << 20].tobytes()
        for k in range(24):
            n = (2048, 2048, 2048, 1000, 4096, 8192)[k % 6]
            blocks.append(seg[k * 40960:k * 40960 + n])
        blocks.append(seg[600000:600000 + (12000, 16384, 9001, 16383)[kind % 4]])      # the 16 KiB shape of the Re-Pair kernel
    import random
    rnd = random.Random(3)                                    # "word soup": a dozen random words repeated — Re-Pair beats LZ77 by a few bytes
    for n in (2048, 3000, 5000, 8192, 2048, 6000):
        words = [bytes(rnd.getrandbits(8) for _ in range(rnd.randint(3, 9))) for _ in range(12)]
        blocks.append(b"".join(rnd.choice(words) for _ in range(n))[:n])
    t, off = G.batch(blocks)
    c = G.ctx()
    full = c.encode_blocks(2, t, off, want_sizes=True)
    assert c.encode_blocks_stats()["repair_stopped_early"] == 0
    fast = c.encode_blocks(2, t, off)
    stopped = c.encode_blocks_stats()["repair_stopped_early"]
    assert np.array_equal(full[1], fast[1]) and np.array_equal(full[2], fast[2])
    n = int(full[1][-1])
    got = fast[0][:n].cpu().numpy().tobytes()
    assert full[0][:n].cpu().numpy().tobytes() == got
    assert stopped >= len(blocks) // 3, stopped
    assert (np.asarray(fast[2]) == 9).any()                  # and Re-Pair still wins where it should
    poff, mids = fast[1], fast[2]
    for b in list(range(0, len(blocks), 5)) + list(range(len(blocks) - 6, len(blocks))):
        mid, payload, _ = O.encode_block(2, blocks[b])
        assert int(mids[b]) == mid and got[poff[b]:poff[b + 1]] == payload, b
