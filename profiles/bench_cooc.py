#!/usr/bin/env python3
"""Co-occurrence count C = B B^T (keyphrase graph, BASELINE configs[4]) on one GPU: tcgen05 kernel vs AND+POPC.
usage: python profiles/bench_cooc.py [K] [D]   (default 10000 keyphrases x 100000 documents)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ast-text-analysis_b200")); sys.path.insert(0, ROOT)
import torch
from east import _capi
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
g = torch.Generator(device="cuda").manual_seed(1)
S = torch.rand((D, K), dtype=torch.float64, device="cuda", generator=g)
C = torch.empty((K, K), dtype=torch.int32, device="cuda")
res = {}
for variant, name in ((0, "tcgen05"), (3, "tcgen05_cp_async"), (2, "tcgen05_simple"), (1, "and_popc")):
    _capi.set_option("cooc_variant", variant)
    for _ in range(2):
        _capi.cooc_dev(S.data_ptr(), D, K, 0.5, C.data_ptr())
    _capi.set_option("time_kernels", 0); _capi.set_option("time_kernels", 1)
    for _ in range(3):
        _capi.cooc_dev(S.data_ptr(), D, K, 0.5, C.data_ptr())
    ks = _capi.kernel_stats(); _capi.set_option("time_kernels", 0)
    res[name] = {k: v["ms"] / v["launches"] for k, v in ks.items()}
    res[name + "_checksum"] = int(C.to(torch.int64).sum().item())
main = res["tcgen05"].get("k_cooc_umma_tma") or res["tcgen05"].get("k_cooc_umma_pipe")
cpa = res["tcgen05_cp_async"].get("k_cooc_umma_pipe")
simple = res["tcgen05_simple"].get("k_cooc_umma")
flops = 2.0 * K * K * D
print(json.dumps({"K": K, "D": D, "ms": res, "tcgen05_TOPS": flops / (main * 1e-3) / 1e12 if main else None,
                  "tcgen05_cp_async_TOPS": flops / (cpa * 1e-3) / 1e12 if cpa else None,
                  "tcgen05_simple_TOPS": flops / (simple * 1e-3) / 1e12 if simple else None,
                  "and_popc_equiv_TOPS": flops / (res["and_popc"]["k_cooc_popc"] * 1e-3) / 1e12,
                  "checksums_equal": res["tcgen05_checksum"] == res["and_popc_checksum"] == res["tcgen05_cp_async_checksum"]}))
