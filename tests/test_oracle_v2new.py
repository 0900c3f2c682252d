"""Oracle restatement of the V2 bit-plane pipeline (method 10) against golden vectors generated from the
unmodified Python reference (tests/golden/make_golden_v2new.py): decode of every vector; encode of the vectors the reference
produced with its own model choice (circuit_map_automaton_forward(parallel=False))."""
import json
import os

import pytest

from oracle import oracle as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "v2new.json")))


@pytest.mark.parametrize("name", sorted(G))
def test_v2new_decode_matches_reference(name):
    v = G[name]
    data, pay = bytes.fromhex(v["input_hex"]), bytes.fromhex(v["payload_hex"])
    assert O.v2new_decode(pay, len(data)) == data


@pytest.mark.parametrize("name", sorted(k for k, v in G.items() if not v["forced"]))
def test_v2new_encode_matches_reference(name):
    v = G[name]
    assert O.v2new_encode(bytes.fromhex(v["input_hex"])) == bytes.fromhex(v["payload_hex"])


def test_v2new_error_cases():
    v = G["text"]
    data, pay = bytes.fromhex(v["input_hex"]), bytes.fromhex(v["payload_hex"])
    for bad in (pay[:2], pay[:len(pay) // 2], bytes([pay[0] | 7]) + pay[1:]):     # short header, truncated planes, param_len > 4
        with pytest.raises(O.OracleError):
            O.v2new_decode(bad, len(data))
    assert O.v2new_decode(b"", 0) == b""


def test_v2new_all_modes_covered():
    assert {(v["mode"]) for v in G.values()} == {0, 1, 2, 3, 4, 5}


@pytest.mark.parametrize("name", sorted(k for k, v in G.items() if v["forced"]))
def test_v2new_forced_encode_matches_reference(name):
    v = G[name]
    assert O.v2new_encode(bytes.fromhex(v["input_hex"]), force=(v["mode"], v["param"])) == bytes.fromhex(v["payload_hex"])
