#!/usr/bin/env python3
"""Tuning sweep of the GLOBAL prefix-doubling sort (option no_doc_sort = 1) on the bench workload
(1000 docs x 50 KB): radix-sort kernel shape (rs_variant) and round-0 window (key_chars).  Prints per-stage device ms (CUDA events) of the best of 3 builds.
usage (GPU box): python profiles/sweep_build.py [--docs 1000] [--doc-bytes 50000]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ast-text-analysis_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from east import _capi

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=1000)
ap.add_argument("--doc-bytes", type=int, default=50000)
ap.add_argument("--variants", default="0,1,2,3,4,5,6,7,8,9")
ap.add_argument("--key-chars", default="0")
ap.add_argument("--batch", default="0", help="sort_batch_elems values (0 = default 3Mi, -1 = whole)")
a = ap.parse_args()
packed, ms, _ = synth.packed_collection(a.docs, a.doc_bytes)
doc_off = np.zeros(a.docs + 1, dtype=np.int64); np.cumsum([len(p) for p in packed], out=doc_off[1:])
doc_m = np.array(ms, dtype=np.int32)
dev = torch.from_numpy(np.concatenate(packed).view(np.int32)).cuda()
_capi.set_option("no_doc_sort", 1)
for _ in range(3):
    _capi.DeviceIndex.build_dev(dev.data_ptr(), doc_off, doc_m).close()
import itertools
for kc, bt in itertools.product([int(x) for x in a.key_chars.split(",")], [int(x) for x in a.batch.split(",")]):
    _capi.set_option("key_chars", kc)
    _capi.set_option("sort_batch_elems", bt)
    for v in [int(x) for x in a.variants.split(",")]:
        _capi.set_option("rs_variant", v)
        best = None
        for _ in range(3):
            _capi.set_option("time_kernels", 0); _capi.set_option("time_kernels", 1)
            t0 = time.perf_counter()
            idx = _capi.DeviceIndex.build_dev(dev.data_ptr(), doc_off, doc_m)
            wall = (time.perf_counter() - t0) * 1e3
            info = idx.info(); idx.close(); st = dict(idx.build_timings)
            ks = _capi.kernel_stats()
            one = ks["k_rs_onesweep"]
            row = (wall, one["ms"], one["launches"], one["bytes"] / (one["ms"] * 1e-3) / 1e9, st, info["rounds"])
            if best is None or row[0] < best[0]:
                best = row
        _capi.set_option("time_kernels", 0)
        print("batch=%d key_chars=%d rs_variant=%d  build_wall=%.2f ms  onesweep=%.3f ms in %d launches (%.0f GB/s)  rounds=%d  stages=%s" % (
            bt, kc, v, best[0], best[1], best[2], best[3], best[5], {k: round(x, 2) for k, x in best[4].items()}))
