#!/usr/bin/env python3
"""Golden vectors for the V2 bit-plane pipeline (method 10) from the UNMODIFIED Python reference.

encode_new_pipeline is dead in the shipped reference: circuit_map_automaton_forward defaults to parallel=True and uses
names the file never imports.  The reference SOURCE is not touched here; the generator only rebinds the
module global `circuit_map_automaton_forward` to
  (a) the reference's own function called with parallel=False (its documented deterministic mode), or
  (b) a wrapper that forces one of the reference's models (so every inverse model is exercised),
then calls the reference's encode_new_pipeline / decode_new_pipeline.  Outputs tests/golden/v2new.json:
    name -> {"input_hex", "orig_len", "mode", "param", "payload_hex", "forced": bool}

    python tests/golden/make_golden_v2new.py          # a few minutes on 8 cores
"""
from __future__ import annotations

import json
import os
import sys
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_loader  # noqa: E402
import datasets  # noqa: E402

FORCED = [(1, 1), (1, 3), (1, 4), (2, 0), (2, 1), (2, 2), (2, 3), (3, 0), (4, 0), (5, 0), (5, 1)]


def _job(args):
    name, data, forced = args
    V = ref_loader.load_v22()
    orig = V.circuit_map_automaton_forward
    if forced is None:
        V.circuit_map_automaton_forward = lambda blk: orig(blk, parallel=False)
    else:
        mode, param = forced
        model = {1: V.ModelDeltaK, 2: V.ModelGrayFamily, 3: V.ModelInterleave, 4: V.ModelBM3, 5: V.ModelMorpho}[mode]()
        V.circuit_map_automaton_forward = lambda blk: (model.forward(blk, param), {"mode": mode, "param": param})
    try:
        payload = V.encode_new_pipeline(data)
    finally:
        V.circuit_map_automaton_forward = orig
    assert V.decode_new_pipeline(payload, len(data)) == data, name
    mode = (payload[0] >> 5) & 7 if payload else 0
    plen = payload[0] & 7 if payload else 0
    param = int.from_bytes(payload[1:1 + plen], "little") if payload else 0
    return name, {"input_hex": data.hex(), "orig_len": len(data), "mode": mode, "param": param,
                  "payload_hex": payload.hex(), "forced": forced is not None}


def main():
    assert ref_loader.available(), "reference not mounted"
    small = datasets.small_cases()
    jobs = []
    for name, data in small.items():
        jobs.append((name, data, None))                    # the reference's own model choice
    pool = ["text", "sine_1000", "gradient_200000", "byte_counter", "random_bytes", "checker_100000", "runs18", "one", "ab", "banana"]
    for i, (mode, param) in enumerate(FORCED):
        for j in range(2):
            nm = pool[(2 * i + j) % len(pool)]
            n = [777, 1024, 3, 64, 1, 2, 1000][(i + j) % 7]
            jobs.append((f"forced_m{mode}_p{param}_{nm}_{n}", small[nm][:n], (mode, param)))
    with Pool(8) as p:
        res = dict(p.map(_job, jobs, chunksize=1))
    with open(os.path.join(HERE, "v2new.json"), "w") as f:
        json.dump(res, f, indent=0, sort_keys=True)
    modes = {}
    for v in res.values():
        modes[(v["mode"], v["param"])] = modes.get((v["mode"], v["param"]), 0) + 1
    print(len(res), "vectors; (mode,param) histogram:", sorted(modes.items()))


if __name__ == "__main__":
    main()
