"""GPU parity of the host-facing BlockPipeline calls: the streaming call (copies overlapped across batches) returns exactly what the
single-batch call returns, and both agree with the oracle."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_streaming_pipeline_matches_single_batch_calls():
    """encode_host_many (copies overlapped across batches) returns exactly what encode_host returns batch by batch, and the
    KF payloads equal the oracle's encode_model_bbwt_mtf."""
    import torch
    from kolmogorovlike_datacompressor_b200 import synth
    from kolmogorovlike_datacompressor_b200.pipeline import BlockPipeline
    import gpu_util as G
    G._ctx.clear()
    torch.cuda.empty_cache()
    pipe = BlockPipeline(1 << 20, 64)
    batches = []
    for seed, cuts in ((1, [0, 70000, 70000, 200001, 300000]), (2, [0, 1, 4097, 150000]), (3, [0, 262144, 524288, 1 << 20])):
        data = np.concatenate([synth.s1_text(cuts[-1] // 2, seed=seed), synth.s2_mixed(cuts[-1] - cuts[-1] // 2)])
        batches.append((torch.from_numpy(data).pin_memory() if seed != 2 else data, np.array(cuts, dtype=np.int64)))
    want = []
    for data, off in batches:
        r = pipe.encode_host(data, off)
        want.append({k: (v.clone() if isinstance(v, torch.Tensor) else np.array(v)) for k, v in r.items() if k not in ("h2d_bytes", "d2h_bytes", "chunks")})
    got = 0
    for r, w, (data, off) in zip(pipe.encode_host_many(iter(batches)), want, batches):
        for k, v in w.items():
            if isinstance(v, torch.Tensor):
                assert torch.equal(r[k], v), k
            else:
                assert np.array_equal(np.asarray(r[k]), v), k
        raw = data.numpy().tobytes() if isinstance(data, torch.Tensor) else data.tobytes()
        b = 0                                                # first block against the oracle
        pay = r["kf_payload"][int(r["kf_off"][b]):int(r["kf_off"][b + 1])].numpy().tobytes()
        blk = raw[off[b]:off[b + 1]]
        assert pay == O.kf_rice_pack(O.mtf_encode(O.bbwt_forward(blk)))
        got += 1
    assert got == len(batches)
    del pipe
    torch.cuda.empty_cache()
