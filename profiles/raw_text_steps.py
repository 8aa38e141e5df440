#!/usr/bin/env python3
"""east_table_texts_host (raw UTF-8 in, table out) on the bench workload: wall time per step and per-kernel device time.
usage (GPU box): python profiles/raw_text_steps.py [--steps 6]"""
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
ap.add_argument("--steps", type=int, default=6)
a = ap.parse_args()
docs = synth.documents(a.docs, a.doc_bytes)
buf, off = _capi.concat_utf8(docs)
pinned = torch.empty(buf.size, dtype=torch.uint8).pin_memory(); raw = pinned.numpy(); raw[:] = buf
codes, koff = _capi.pack_keyphrases([utils.prepare_text(k) for k in synth.keyphrases(a.keyphrases)])
out_t = torch.empty(a.docs * a.keyphrases, dtype=torch.float64).pin_memory()
out = out_t.numpy().reshape(a.docs, a.keyphrases)
for i in range(a.steps):
    if i == a.steps - 1:
        _capi.set_option("time_kernels", 1)
    t0 = time.perf_counter()
    idx = _capi.DeviceIndex.table_from_texts((raw, off), codes, koff, out, True)
    t1 = time.perf_counter()
    idx.close()
    print("step %d: %.3f ms" % (i, (t1 - t0) * 1e3))
ks = _capi.kernel_stats(); _capi.set_option("time_kernels", 0)
for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms"])[:8]:
    print("  %-24s %2d launches %8.3f ms" % (k, v["launches"], v["ms"]))
