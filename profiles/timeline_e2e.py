#!/usr/bin/env python3
"""Launch timeline of ONE end-to-end step (east_table_host, pinned host text -> table on the host) of the bench
workload: start offset and duration of every kernel, from the library's own event pairs.
usage (GPU box): EAST_DEBUG_TIMELINE=1 python profiles/timeline_e2e.py [--docs 1000] [--doc-bytes 50000] 2> timeline.txt
(the events sit between the kernels of a stream: the step runs ~3 % slower than untimed)"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ast-text-analysis_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from east import _capi, utils

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=1000)
ap.add_argument("--doc-bytes", type=int, default=50000)
ap.add_argument("--keyphrases", type=int, default=1000)
a = ap.parse_args()
packed, ms, _ = synth.packed_collection(a.docs, a.doc_bytes)
doc_off = np.zeros(a.docs + 1, dtype=np.int64); np.cumsum([len(p) for p in packed], out=doc_off[1:])
doc_m = np.array(ms, dtype=np.int32)
host = torch.empty(int(doc_off[-1]), dtype=torch.int32).pin_memory()
text = host.numpy().view(np.uint32); text[:] = np.concatenate(packed)
codes, off = _capi.pack_keyphrases([utils.prepare_text(k) for k in synth.keyphrases(a.keyphrases)])
out_t = torch.empty(a.docs * a.keyphrases, dtype=torch.float64).pin_memory()
out = out_t.numpy().reshape(a.docs, a.keyphrases)
for _ in range(4):
    _capi.DeviceIndex.build_host_and_score(text, doc_off, doc_m, codes, off, out).close()
_capi.set_option("time_kernels", 1)
t0 = time.perf_counter()
idx = _capi.DeviceIndex.build_host_and_score(text, doc_off, doc_m, codes, off, out)
t1 = time.perf_counter()
idx.close()
_capi.set_option("time_kernels", 0)
sys.stderr.write("[east] step wall %.3f ms (with event pairs)\n" % ((t1 - t0) * 1e3))
