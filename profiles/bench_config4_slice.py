#!/usr/bin/env python3
"""A one-GPU slice of BASELINE configs[3] (100k keyphrases x 100k docs of ~10 KB, docs sharded over 8 GPUs):
K keyphrases x D documents on ONE GPU, build + score, device-timed.  The full job is 12 500 docs per GPU; the
slice runs D docs and reports scores/s, which is independent of D once D >> 148 SMs.
usage: python profiles/bench_config4_slice.py [K=100000] [D=1000] [doc_bytes=10000]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ast-text-analysis_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from east import _capi, utils
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
nbytes = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
packed, ms, _ = synth.packed_collection(D, nbytes)
doc_off = np.zeros(D + 1, dtype=np.int64); np.cumsum([len(p) for p in packed], out=doc_off[1:])
doc_m = np.array(ms, dtype=np.int32)
dev = torch.from_numpy(np.concatenate(packed).view(np.int32)).cuda()
kps = [utils.prepare_text(k) for k in synth.keyphrases(K)]
codes, off = _capi.pack_keyphrases(kps)
kp_dev = torch.from_numpy(codes.view(np.int32).copy()).cuda()
out = torch.empty(D * K, dtype=torch.float64, device="cuda")
res = []
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx = _capi.DeviceIndex.build_dev(dev.data_ptr(), doc_off, doc_m)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    idx.score_table_dev(kp_dev.data_ptr(), off, out.data_ptr(), True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    st = dict(idx.score_timings); idx.close()
    res.append((t1 - t0, t2 - t1, st["score"]))
b, sc, dev_ms = min(res, key=lambda r: r[1])
print(json.dumps({"K": K, "D": D, "doc_bytes": nbytes, "query_suffixes": int(off[-1]), "build_s": b, "score_wall_s": sc,
                  "score_device_ms": dev_ms, "scores_per_s_score_only": D * K / (dev_ms * 1e-3),
                  "scores_per_s_build_and_score": D * K / (b + sc), "checksum": float(out.sum().item())}))
