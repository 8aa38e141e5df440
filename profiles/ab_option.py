#!/usr/bin/env python3
"""A/B of one library option on the bench workload: wall time per step of the device-resident call (east_table_dev) and
of the host-buffer call (east_table_host_u8), option off / on in alternating rounds.
usage (GPU box): python profiles/ab_option.py no_alphabet_guess [--steps 40] [--rounds 3]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ast-text-analysis_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from east import _capi, utils
from east.asts import utils as au

ap = argparse.ArgumentParser()
ap.add_argument("option")
ap.add_argument("--docs", type=int, default=1000)
ap.add_argument("--doc-bytes", type=int, default=50000)
ap.add_argument("--keyphrases", type=int, default=1000)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--rounds", type=int, default=3)
a = ap.parse_args()
packed, ms, cols = synth.packed_collection(a.docs, a.doc_bytes)
packed8 = [au.pack_strings_collection_u8(c) for c in cols]
doc_off = np.zeros(a.docs + 1, dtype=np.int64); np.cumsum([len(p) for p in packed], out=doc_off[1:])
doc_m = np.array(ms, dtype=np.int32)
host = torch.empty(int(doc_off[-1]), dtype=torch.uint8).pin_memory()
text8 = host.numpy(); text8[:] = np.concatenate(packed8)
dev = torch.from_numpy(np.concatenate(packed).view(np.int32)).cuda()
codes, off = _capi.pack_keyphrases([utils.prepare_text(k) for k in synth.keyphrases(a.keyphrases)])
kp_dev = torch.from_numpy(codes.view(np.int32).copy()).cuda()
out_dev = torch.empty(a.docs * a.keyphrases, dtype=torch.float64, device="cuda")
out_t = torch.empty(a.docs * a.keyphrases, dtype=torch.float64).pin_memory()
out = out_t.numpy().reshape(a.docs, a.keyphrases)

def step_dev():
    _capi.DeviceIndex.build_dev_and_score(dev.data_ptr(), doc_off, doc_m, kp_dev.data_ptr(), codes, off, out_dev.data_ptr(), True).close()

def step_host():
    _capi.DeviceIndex.build_host_and_score(text8, doc_off, doc_m, codes, off, out).close()

def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / a.steps

for r in range(a.rounds):
    for v in (0, 1):
        _capi.set_option(a.option, v)
        print("round %d  %s=%d  device %.3f ms  host-buffer %.3f ms" % (r, a.option, v, timed(step_dev), timed(step_host)), flush=True)
_capi.set_option(a.option, 0)
