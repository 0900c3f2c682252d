"""The reference's benchmark_compare experiment (benchmark_compare.py:157-232) on the drop-in: every coder round-trips and
the byte counts equal what the CPU oracle produces for the same data sets."""
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_benchmark_compare_rows_match_oracle():
    from kolmogorovlike_datacompressor_b200 import benchmark_compare as BC
    sets = BC.data_sets()
    rows = BC.run_benchmarks(sets)
    assert len(rows) == 15 and all(r["valid"] for r in rows)
    by = {(r["dataset"], r["algorithm"]): r for r in rows}
    for name, data in sets.items():
        assert by[(name, "kolm_final")]["bytes"] == len(O.kf_compress(data, 8192)), name
        assert by[(name, "baseline_lz77")]["bytes"] == len(O.lz77_encode(data, 255, 127)), name
        m = O.mtf_encode(O.bbwt_forward(data))
        assert BC.mtf_encode(BC.bbwt_forward(data)) == list(m), name
