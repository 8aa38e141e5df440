import sys, os
sys.path.insert(0, "ast-text-analysis_b200"); sys.path.insert(0, ".")
import numpy as np
import synth
from east import _capi, utils
packed, ms, _ = synth.packed_collection(5, 6000, first_seed=5)
idx = _capi.DeviceIndex(packed, ms)
kps = [utils.prepare_text(k) for k in synth.keyphrases(40)]
codes, off = _capi.pack_keyphrases(kps)
t = idx.score_table(codes, off, True)
print(idx.info(), float(t.sum()))
a = idx.array(0, _capi.ANNTAB); print(int(a.sum()))
from east.asts import utils as au
deep = [au.pack_strings_collection(["AB" * 600, "B" * 700]), au.pack_strings_collection(["XABXAC", "HI"]), au.pack_strings_collection(["A" * 2000] * 3)]
idx2 = _capi.DeviceIndex(deep, [2, 2, 3]); print(idx2.info(), int(idx2.array(2, _capi.LCPTAB).sum()))
# one-call entries: the per-document kernel scores from shared memory; speculative runs of a pipelined build
doc_off = np.zeros(len(packed) + 1, dtype=np.int64); np.cumsum([len(p) for p in packed], out=doc_off[1:])
text = np.ascontiguousarray(np.concatenate(packed), dtype=np.uint32)
out = np.zeros((len(packed), len(kps)))
idx3 = _capi.DeviceIndex.build_host_and_score(text, doc_off, ms, codes, off, out)
print("fused == two calls:", bool(np.array_equal(out.view(np.uint64), t.view(np.uint64))))
many, mm, _ = synth.packed_collection(300, 700, first_seed=9)
many.append(au.pack_strings_collection(["中文", "AB"])); mm.append(2)          # a broken layout in the last run
many.append(au.pack_strings_collection(["E" * 5000, "END"])); mm.append(2)    # a bucket too large
doc_off = np.zeros(len(many) + 1, dtype=np.int64); np.cumsum([len(p) for p in many], out=doc_off[1:])
text = np.ascontiguousarray(np.concatenate(many), dtype=np.uint32)
out = np.zeros((len(many), len(kps)))
_capi.set_option("pipeline_chunk", 30000)
idx4 = _capi.DeviceIndex.build_host_and_score(text, doc_off, mm, codes, off, out)
print("pipelined:", idx4.stat("pipelined"), "miss:", idx4.stat("pipeline_miss"), float(out.sum()))
idx5 = _capi.DeviceIndex.build_host_and_score(text[:doc_off[300]].copy(), doc_off[:301], mm[:300], codes, off, out[:300])
print("pipelined:", idx5.stat("pipelined"), float(out[:300].sum()))
