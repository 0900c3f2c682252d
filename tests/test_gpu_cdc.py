"""GPU-assisted content-defined chunking: `Engine.cdc_boundaries` (kernel `k_cdc_candidates` + host walk)
must return exactly the boundaries of the reference's cdc_fast_boundaries (kolm_final.py:161-194) and
cdc_fast_boundaries_strict (kolm_final_researched_v2-2.py:210-309) — checked against the oracle, across piece boundaries."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _data():
    from kolmogorovlike_datacompressor_b200 import synth
    rng = np.random.default_rng(11)
    return np.concatenate([synth.s3_mix(3 << 20), rng.integers(0, 256, 300001, dtype=np.uint8), synth.s1_text(500000, seed=9)]).tobytes()


@pytest.mark.parametrize("params", [(32, 64, 128), (100, 300, 1000), (4096, 8192, 16384), (65536, 131072, 262144)])
def test_gpu_cdc_matches_oracle(params):
    from kolmogorovlike_datacompressor_b200.engine import Engine
    mn, avg, mx = params
    data = _data()
    eng = Engine.shared()
    for piece in (64 << 20, (1 << 20) + 7):                  # one piece / many pieces with history across the seams
        assert eng.cdc_boundaries("kf", data, mn, avg, mx, piece=piece) == O.kf_cdc(data, mn, avg, mx)
        assert eng.cdc_boundaries("v22", data, mn, avg, mx, piece=piece) == O.v22_cdc(data, mn, avg, mx)


def test_gpu_cdc_degenerate_and_dropin():
    from kolmogorovlike_datacompressor_b200.engine import Engine
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF, kolm_final_researched_v2_2 as V
    eng = Engine.shared()
    zeros = bytes(3 << 20)                                   # every position (or none) passes: the sparse list overflows -> host scan
    assert eng.cdc_boundaries("kf", zeros, 4096, 8192, 16384) == O.kf_cdc(zeros, 4096, 8192, 16384)
    assert eng.cdc_boundaries("v22", zeros, 4096, 8192, 16384) == O.v22_cdc(zeros, 4096, 8192, 16384)
    data = _data()
    old_kf, old_v = KF.GPU_CDC_MIN_BYTES, V.GPU_CDC_MIN_BYTES
    try:
        KF.GPU_CDC_MIN_BYTES = V.GPU_CDC_MIN_BYTES = 1      # force the GPU scan inside the drop-in functions
        assert KF.cdc_fast_boundaries(data, 4096, 8192, 16384) == O.kf_cdc(data, 4096, 8192, 16384)
        assert V.cdc_fast_boundaries_strict(data, 2048, 4096, 8192) == O.v22_cdc(data, 2048, 4096, 8192)
        with pytest.raises(ValueError):
            V.cdc_fast_boundaries_strict(data, 10, 32, 64)
    finally:
        KF.GPU_CDC_MIN_BYTES, V.GPU_CDC_MIN_BYTES = old_kf, old_v
