#!/usr/bin/env python3
"""Build (and score) the bench workload a few times: a short target for ncu captures.
usage: python profiles/run_build_once.py [docs] [doc_bytes] [iters] [option=value ...]
(EAST_BENCH_E2E=two_calls: east_build_dev + east_score_table_dev instead of east_table_dev)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ast-text-analysis_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from east import _capi, utils
docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nbytes = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    _capi.set_option(k, int(v))
packed, ms, _ = synth.packed_collection(docs, nbytes)
doc_off = np.zeros(docs + 1, dtype=np.int64); np.cumsum([len(p) for p in packed], out=doc_off[1:])
doc_m = np.array(ms, dtype=np.int32)
dev = torch.from_numpy(np.concatenate(packed).view(np.int32)).cuda()
kps = [utils.prepare_text(k) for k in synth.keyphrases(1000)]
codes, off = _capi.pack_keyphrases(kps)
kp_dev = torch.from_numpy(codes.view(np.int32).copy()).cuda()
out = torch.empty(docs * 1000, dtype=torch.float64, device="cuda")
two_calls = os.environ.get("EAST_BENCH_E2E", "") == "two_calls"
for _ in range(iters):
    if two_calls:
        idx = _capi.DeviceIndex.build_dev(dev.data_ptr(), doc_off, doc_m)
        idx.score_table_dev(kp_dev.data_ptr(), off, out.data_ptr(), True)
    else:   # east_table_dev: the per-document kernel scores its document itself
        idx = _capi.DeviceIndex.build_dev_and_score(dev.data_ptr(), doc_off, doc_m, kp_dev.data_ptr(), codes, off, out.data_ptr(), True)
    info = idx.info(); idx.close()
    print(info, [(n, round(m, 3)) for n, m in idx.build_timings + getattr(idx, "score_timings", [])])
torch.cuda.synchronize()
