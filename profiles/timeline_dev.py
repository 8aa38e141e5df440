#!/usr/bin/env python3
"""Launch timeline of ONE device-resident step (east_table_dev) of the bench workload.
usage (GPU box): EAST_DEBUG_TIMELINE=1 EAST_DEBUG_TIMING=1 python profiles/timeline_dev.py 2> timeline.txt"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ast-text-analysis_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from east import _capi, utils
docs, nbytes = 1000, 50000
packed, ms, _ = synth.packed_collection(docs, nbytes)
doc_off = np.zeros(docs + 1, dtype=np.int64); np.cumsum([len(p) for p in packed], out=doc_off[1:])
doc_m = np.array(ms, dtype=np.int32)
dev = torch.from_numpy(np.concatenate(packed).view(np.int32)).cuda()
codes, off = _capi.pack_keyphrases([utils.prepare_text(k) for k in synth.keyphrases(1000)])
kp_dev = torch.from_numpy(codes.view(np.int32).copy()).cuda()
out = torch.empty(docs * 1000, dtype=torch.float64, device="cuda")
def step():
    t0 = time.perf_counter()
    idx = _capi.DeviceIndex.build_dev_and_score(dev.data_ptr(), doc_off, doc_m, kp_dev.data_ptr(), codes, off, out.data_ptr(), True)
    t1 = time.perf_counter()
    idx.close()
    return (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3
for _ in range(4):
    step()
torch.cuda.synchronize()
for i in range(4):
    sys.stderr.write("[east] ---- step %d\n" % i)
    a, b = step()
    sys.stderr.write("[east] step wall: call %.3f ms, close %.3f ms\n" % (a, b))
_capi.set_option("time_kernels", 1)
a, b = step()
_capi.set_option("time_kernels", 0)
sys.stderr.write("[east] step wall with event pairs: call %.3f ms\n" % a)
