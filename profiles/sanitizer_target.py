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
# ---- round 2 paths: one byte per code point (pipelined in-kernel decode + device expansion), raw texts (tokenizer),
# device keyphrase preparation with odd keyphrases, sampled alphabet with a miss, persistence, co-occurrence (TMA cluster kernel)
cols300 = [utils.text_to_strings_collection(d) for d in synth.documents(300, 700, first_seed=9)]
p8 = [au.pack_strings_collection_u8(c) for c in cols300]
m8 = [len(c) for c in cols300]
off8 = np.zeros(len(p8) + 1, dtype=np.int64); np.cumsum([len(p) for p in p8], out=off8[1:])
t8 = np.ascontiguousarray(np.concatenate(p8), dtype=np.uint8)
out8 = np.zeros((300, len(kps)))
idx6 = _capi.DeviceIndex.build_host_and_score(t8, off8, m8, codes, off, out8)
print("u8 pipelined:", idx6.stat("pipelined"), bool(np.array_equal(out8.view(np.uint64), out[:300].view(np.uint64))))
idx7 = _capi.DeviceIndex.build_host_u8(t8[:off8[7]].copy(), off8[:8], m8[:7]); print("u8 small:", idx7.info()["doc_sorted"])
odd = kps + ["E", "E", "Q" * 300, "中文A", "ABਁC"]
c2, o2 = _capi.pack_keyphrases(odd)
raw_out = np.zeros((302, len(odd)))
texts = synth.documents(300, 700, first_seed=9) + ["", "привет мир – «ёжик» №5 привет"]
idx8 = _capi.DeviceIndex.table_from_texts(texts, c2, o2, raw_out)
print("raw texts:", idx8.info()["doc_sorted"], float(raw_out.sum()), idx8.strings_collection(301))
_capi.set_option("alphabet_sample", 3000)
import torch
late = many[:300] + [au.pack_strings_collection(["0123456789 QUIZ", "ZEBRA9 J"])]
doc_off = np.zeros(len(late) + 1, dtype=np.int64); np.cumsum([len(p) for p in late], out=doc_off[1:])
td = torch.from_numpy(np.concatenate(late).astype(np.uint32).view(np.int32)).cuda()
kd = torch.from_numpy(c2.view(np.int32).copy()).cuda()
od = torch.zeros((301, len(odd)), dtype=torch.float64, device="cuda")
idx9 = _capi.DeviceIndex.build_dev_and_score(td.data_ptr(), doc_off, mm[:300] + [2], kd.data_ptr(), c2, o2, od.data_ptr(), True)
print("sampled alphabet miss:", idx9.stat("alphabet_miss"), float(od.sum().item()))
_capi.set_option("alphabet_sample", 0)
idx9.save("/tmp/east_sanitizer.idx"); idx10 = _capi.DeviceIndex.load("/tmp/east_sanitizer.idx")
print("loaded:", bool(np.array_equal(idx10.array(300, _capi.SUFTAB), idx9.array(300, _capi.SUFTAB))))
S = np.random.default_rng(1).random((700, 300))
C = _capi.cooc_host(S, 0.5); B = S >= 0.5
print("cooc exact:", bool(np.array_equal(C, (B.T.astype(np.int64) @ B.astype(np.int64)).astype(np.int32))))
