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
