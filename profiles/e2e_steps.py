#!/usr/bin/env python3
"""Wall time of consecutive end-to-end steps (east_table_host / east_table_host_u8) of the bench workload, one line per step.
usage (GPU box): python profiles/e2e_steps.py [--width 1|4] [--steps 12]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ast-text-analysis_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from east import _capi, utils
from east.asts import utils as au

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=1000)
ap.add_argument("--doc-bytes", type=int, default=50000)
ap.add_argument("--keyphrases", type=int, default=1000)
ap.add_argument("--width", type=int, default=1)
ap.add_argument("--steps", type=int, default=12)
a = ap.parse_args()
packed, ms, cols = synth.packed_collection(a.docs, a.doc_bytes)
if a.width == 1:
    packed = [au.pack_strings_collection_u8(c) for c in cols]
doc_off = np.zeros(a.docs + 1, dtype=np.int64); np.cumsum([len(p) for p in packed], out=doc_off[1:])
doc_m = np.array(ms, dtype=np.int32)
host = torch.empty(int(doc_off[-1]), dtype=torch.uint8 if a.width == 1 else torch.int32).pin_memory()
text = host.numpy() if a.width == 1 else host.numpy().view(np.uint32)
text[:] = np.concatenate(packed)
codes, off = _capi.pack_keyphrases([utils.prepare_text(k) for k in synth.keyphrases(a.keyphrases)])
out_t = torch.empty(a.docs * a.keyphrases, dtype=torch.float64).pin_memory()
out = out_t.numpy().reshape(a.docs, a.keyphrases)
for i in range(a.steps):
    t0 = time.perf_counter()
    idx = _capi.DeviceIndex.build_host_and_score(text, doc_off, doc_m, codes, off, out)
    t1 = time.perf_counter()
    idx.close()
    t2 = time.perf_counter()
    print("step %2d: call %.3f ms, close %.3f ms, pipelined=%s" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, "?"))
