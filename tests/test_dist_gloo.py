"""World-size-2 gloo test (CPU) of the block-sharding host logic: partitioning, table all_gather, payload send/recv,
ordering.  The per-block encoder is injected (the CPU oracle) because this container has no GPU."""
import os
import socket

import pytest
import torch.multiprocessing as mp

import datasets


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from kolmogorovlike_datacompressor_b200 import dist as kd
    from oracle import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = datasets.medium_cases()["text_big"] + datasets.fixture("pattern")[:30000]
    bounds = O.kf_cdc(data, 1024, 2048, 4096)

    def enc(d, bs):
        out = []
        for a, b in bs:
            mid, payload, _ = O.encode_block(O.PROFILE_KOLM, d[a:b])
            out.append((mid, payload))
        return out

    got = kd.sharded_encode(data, bounds, enc)
    if rank == 0:
        want = enc(data, bounds)
        q.put(("ok", got == want, len(bounds), kd.partition_blocks(bounds, world)))
    else:
        q.put(("none", got is None))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_encode_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    ok = [r for r in res if r[0] == "ok"][0]
    assert ok[1] is True
    parts = ok[3]
    assert parts[0][0] == 0 and parts[-1][1] == ok[2] and parts[0][1] == parts[1][0]
    assert [r for r in res if r[0] == "none"][0][1] is True


def test_partition_balanced_and_total():
    from kolmogorovlike_datacompressor_b200 import dist as kd
    bounds = [(i * 1000, (i + 1) * 1000) for i in range(37)]
    for world in (1, 2, 3, 4, 8, 64):
        parts = kd.partition_blocks(bounds, world)
        assert parts[0][0] == 0 and parts[-1][1] == 37
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        sizes = [e - s for s, e in parts]
        assert max(sizes) - min(sizes) <= 1 or world > 37
    assert kd.partition_blocks([], 4) == [(0, 0)] * 4


def _worker_area(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    from kolmogorovlike_datacompressor_b200 import dist as kd
    from oracle import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = datasets.fixture("pattern")[:50000] + datasets.medium_cases()["text_big"]
    bounds = [(a, min(len(data), a + 2048)) for a in range(0, len(data), 2048)]

    def area_fn(d, bs):
        enc = [O.encode_block(O.PROFILE_KOLR, d[a:b])[:2] for a, b in bs]
        return (np.array([m for m, _ in enc], dtype=np.int64), np.array([len(p) for _, p in enc], dtype=np.int64),
                np.frombuffer(b"".join(p for _, p in enc), dtype=np.uint8))

    got = kd.sharded_encode_area(data, bounds, area_fn)
    if rank == 0:
        want = area_fn(data, bounds)
        q.put(("ok", all(np.array_equal(g, w) for g, w in zip(got, want)), len(bounds)))
    else:
        q.put(("none", got is None))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_encode_area_three_ranks():
    """Array form used by dist.compress_kolm / compress_kolr_*: ids, lengths and the payload area arrive in block order on rank 0."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_area, args=(r, 3, port, q)) for r in range(3)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r for r in res if r[0] == "ok"][0][1] is True
    assert sum(1 for r in res if r[0] == "none" and r[1]) == 2


def _worker_dec(rank, world, port, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    from kolmogorovlike_datacompressor_b200 import dist as kd, kolm_final as KF
    from oracle import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = datasets.fixture("pattern")[:40000] + datasets.medium_cases()["text_big"] + bytes(3000)
    blob = O.kf_compress(data, 2048)
    names, starts, plens, olens, total = KF._parse(blob)
    ids = {"raw": 0, "kf_xor": 1, "kf_bbwt": 2, "kf_lz77": 3}

    def dec(b, nm, st, pl, ol):
        out = b"".join(O.decode_model(O.PROFILE_KOLM, ids[n], b[s:s + l], o) for n, s, l, o in zip(nm, st, pl, ol))
        return torch.frombuffer(bytearray(out), dtype=torch.uint8)

    got = kd._sharded_decode(blob, names, starts, plens, olens, dec, piece=7000)
    q.put(("ok", got == data and total == len(data)) if rank == 0 else ("none", got is None))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_decode_two_ranks():
    """dist.decompress_*: block ranges decoded on different ranks arrive in order on rank 0 (pieces smaller than a range)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_dec, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] is True for r in res)


def _corpus_parts():
    a = datasets.fixture("pattern")[:30000] + datasets.medium_cases()["text_big"]
    b = datasets.fixture("checker")[:9000]
    c = b""
    d = datasets.fixture("sine")[:12345] + bytes(5000)
    return [a, b, c, d]


def _kolr_area_fn(d, bs):
    import numpy as np
    from oracle import oracle as O
    enc = [O.encode_block(O.PROFILE_KOLR, bytes(d[a:b]))[:2] for a, b in bs]
    return (np.array([m for m, _ in enc], dtype=np.int64), np.array([len(p) for _, p in enc], dtype=np.int64),
            np.frombuffer(b"".join(p for _, p in enc), dtype=np.uint8))


def _worker_corpus(rank, world, port, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    from kolmogorovlike_datacompressor_b200 import dist as kd, kolm_final_researched_v2_2 as V
    from oracle import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = _corpus_parts()
    loads = []

    def load(k, a, b):
        loads.append((k, a, b))
        return parts[k][a:b]

    st = {}
    got = kd.compress_kolr_fixed_corpus([len(p) for p in parts], load, 2048, stats=st, area_fn=_kolr_area_fn)
    ids = {n: i for i, n in enumerate(V.KOLR_NAMES)}

    def dec(span, nm, stt, pl, ol):
        raw = span.numpy().tobytes()
        out = b"".join(O.decode_model(O.PROFILE_KOLR, ids[n], raw[s:s + l], o) for n, s, l, o in zip(nm, stt, pl, ol))
        return torch.frombuffer(bytearray(out), dtype=torch.uint8) if out else torch.empty(0, dtype=torch.uint8)

    back = kd.decompress_kolr_corpus(got, decode_fn=dec)
    shard = kd.decompress_kolr_corpus(got, decode_fn=dec, gather=False)
    shard_ok = all(parts[k][a:b] == y.numpy().tobytes() for k, a, b, y in shard)
    loaded = sum(b - a for _, a, b in loads)
    if rank == 0:
        want = []
        for p in parts:
            bounds = [(a, min(len(p), a + 2048)) for a in range(0, len(p), 2048)]
            want.append(V._assemble(len(p), bounds, V.MODE_FIXED, 2048, encoded=_kolr_area_fn(p, bounds) if bounds else None))
        q.put(("ok", got == want, back == parts, shard_ok, loaded, sorted(st)))
    else:
        q.put(("none", got is None and back is None, shard_ok, loaded))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_corpus_of_containers_sharded(world):
    """dist.compress_kolr_fixed_corpus / decompress_kolr_corpus: four containers (one empty) cut into block ranges over the ranks;
    every rank loads only its own bytes; rank 0 gets byte-identical containers and the round trip, gathered and sharded."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_corpus, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    ok = [r for r in res if r[0] == "ok"][0]
    assert ok[1] is True and ok[2] is True and ok[3] is True
    assert {"encode_s", "table_allgather_s", "payload_exchange_s", "d2h_s", "assemble_s"} <= set(ok[5])
    others = [r for r in res if r[0] == "none"]
    assert len(others) == world - 1 and all(r[1] and r[2] for r in others)
    total = sum(len(p) for p in _corpus_parts())
    assert ok[4] + sum(r[3] for r in others) == total          # every input byte was loaded exactly once, by the rank that encodes it
    assert max([ok[4]] + [r[3] for r in others]) < total * 0.75
