"""Small target for compute-sanitizer --tool racecheck: the per-document kernel on the one-byte text path (phase 1's block
scan of string ends), the tokenizer, the keyphrase preparation."""
import sys
sys.path.insert(0, "ast-text-analysis_b200"); sys.path.insert(0, ".")
import numpy as np
import synth
from east import _capi, utils
from east.asts import utils as au
docs = synth.documents(300, 600, first_seed=9)
cols = [utils.text_to_strings_collection(d) for d in docs]
p8 = [au.pack_strings_collection_u8(c) for c in cols]
m8 = [len(c) for c in cols]
off8 = np.zeros(len(p8) + 1, dtype=np.int64); np.cumsum([len(p) for p in p8], out=off8[1:])
t8 = np.ascontiguousarray(np.concatenate(p8), dtype=np.uint8)
codes, off = _capi.pack_keyphrases([utils.prepare_text(k) for k in synth.keyphrases(30)])
out = np.zeros((300, 30))
_capi.set_option("pipeline_chunk", 30000)
idx = _capi.DeviceIndex.build_host_and_score(t8, off8, m8, codes, off, out)
print("u8 pipelined:", idx.stat("pipelined"), float(out.sum()))
out2 = np.zeros((300, 30))
idx2 = _capi.DeviceIndex.table_from_texts(docs, codes, off, out2)
print("raw:", bool(np.array_equal(out.view(np.uint64), out2.view(np.uint64))))
