"""Host logic: cutting a corpus into containers that respect the KOLM / KOLR header limits."""
import pytest

from kolmogorovlike_datacompressor_b200.corpus import MAX_BLOCKS, MAX_BYTES, container_spans


def test_spans_cover_and_respect_limits():
    for total, blk in ((0, 4096), (1, 4096), (4096 * 7 + 5, 4096), (16 << 30, 1 << 20), (16 << 30, 64 << 10), (5 << 30, 2048), ((4 << 30) - 1, 1 << 20)):
        spans = container_spans(total, blk)
        assert spans[0][0] == 0 and spans[-1][1] == total
        for (a, b), (c, _d) in zip(spans, spans[1:]):
            assert b == c
        for a, b in spans:
            assert b - a <= MAX_BYTES and -(-(b - a) // blk) <= MAX_BLOCKS
            assert a % blk == 0                       # fixed blocking of a span == fixed blocking of the whole input


def test_spans_examples():
    assert container_spans(16 << 30, 1 << 20) == [(i * 4095 << 20, min(16 << 30, (i + 1) * 4095 << 20)) for i in range(5)]   # u32 byte limit binds
    assert len(container_spans(16 << 30, 64 << 10)) == 5                                                                    # 65 535 blocks of 64 KiB ~ 4 GiB
    assert len(container_spans(1 << 30, 2048)) == 9                                                                         # block limit binds: 65 535 * 2048 bytes
    assert container_spans(10, 4, max_bytes=9, max_blocks=100) == [(0, 8), (8, 10)]
    with pytest.raises(ValueError):
        container_spans(10, 8, max_bytes=7)


def test_fixed_plan_equals_the_block_lists():
    """dist._fixed_plan (block ranges, partition and per-rank runs of a corpus of fixed-size blocks by arithmetic) against the
    per-block lists it replaces: corpus_blocks -> _virtual_bounds -> partition_blocks -> _my_runs."""
    import random
    from kolmogorovlike_datacompressor_b200 import dist as kd
    rnd = random.Random(12)
    for _ in range(400):
        sizes = [rnd.choice([0, 0, 1, 5, 100, 2047, 2048, 2049, 8192, 30000, 70001]) for _ in range(rnd.randint(0, 6))]
        bs = rnd.choice([1, 7, 100, 2048, 8192])
        if sum(sizes) // bs > 30000:
            continue
        blocks = kd.corpus_blocks(sizes, bs)
        for world in (1, 2, 3, 8):
            want_parts = kd.partition_blocks(kd._virtual_bounds(blocks), world)
            for rank in range(world):
                parts, nblocks, runs = kd._fixed_plan(sizes, bs, world, rank)
                assert parts == want_parts and nblocks == len(blocks), (sizes, bs, world)
                b0, b1 = want_parts[rank]
                want = []
                for k, i, j in kd._my_runs(blocks, b0, b1):
                    lo, hi = blocks[i][1], blocks[j - 1][2]
                    want.append((k, lo, hi, [(a - lo, b - lo) for _, a, b in blocks[i:j]]))
                assert runs == want, (sizes, bs, world, rank)
