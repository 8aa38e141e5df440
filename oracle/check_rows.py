#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (checker): compares rows of a score table with the CPU oracle, bit for bit.

bench.py runs this in a subprocess after its timed regions (the benchmark process itself never loads anything from
oracle/).  Input: an .npz with  kp_codes, kp_off, normalized  and, per row i,  text_i (packed uint32 document), m_i
(strings), row_i (float64 scores the GPU path produced).  Prints one JSON line: {"rows": n, "mismatching_rows": k}."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle  # noqa: E402  (oracle/oracle.py: run as a script, this directory is the import root)


def main(path):
    z = np.load(path)
    codes, off, normalized = z["kp_codes"], z["kp_off"], bool(z["normalized"])
    n = int(z["n_rows"])
    bad = 0
    for i in range(n):
        exp = oracle.OracleEASA(text=z["text_%d" % i], m=int(z["m_%d" % i])).score_many(codes, off, normalized)
        got = np.ascontiguousarray(z["row_%d" % i], dtype=np.float64)
        bad += int(not np.array_equal(exp.view(np.uint64), got.view(np.uint64)))
    print(json.dumps({"rows": n, "mismatching_rows": bad}))


if __name__ == "__main__":
    oracle.build()
    main(sys.argv[1])
