"""Driver entry points: build() compiles everything for sm_100a, smoke() runs one tiny
build + score on cuda:0 and checks it against the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ast-text-analysis_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def build():
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> ast-text-analysis_b200/lib/libeast_b200.so
    (Makefile in the package), the C oracle -> oracle/liboracle.so, and -- only where /root/reference
    exists -- the py3-patched reference copy -> oracle/_ref/ (test/bench infrastructure)."""
    subprocess.check_call(["make", "-C", PKG, "-j4"])
    from oracle import oracle
    oracle.build()
    if os.path.isdir("/root/reference/east") and not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "east")):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "make_ref.py")])
    import east  # noqa: F401
    from east import _capi
    _capi.load()


def smoke():
    """README usage of the reference (README.rst:147-152) on cuda:0, plus a 3-document batch checked
    against the CPU oracle bit for bit."""
    import numpy as np
    import east  # noqa: F401
    from east import _capi, utils
    from east.asts import base
    from oracle import oracle
    import synth

    ast = base.AST.get_ast(["XABXAC", "HI"])
    assert ast.score("ABCI") == 0.1875 and ast.score("NOPE") == 0
    assert ast.suftab.tolist() == [1, 4, 2, 5, 7, 8, 0, 3, 6, 9]

    packed, ms, _ = synth.packed_collection(3, 5000)
    idx = _capi.DeviceIndex(packed, ms)
    kps = [utils.prepare_text(k) for k in synth.keyphrases(16)]
    codes, off = _capi.pack_keyphrases(kps)
    table = idx.score_table(codes, off, True)
    for d in range(3):
        o = oracle.OracleEASA(text=packed[d], m=ms[d])
        for which, name in ((_capi.SUFTAB, "suftab"), (_capi.LCPTAB, "lcptab"), (_capi.ANNTAB, "anntab")):
            assert np.array_equal(idx.array(d, which), getattr(o, name)), name
        exp = o.score_many(codes, off, True)
        assert np.array_equal(table[d].view(np.uint64), exp.view(np.uint64))
    print("smoke ok: launches=%d" % _capi.launch_count())


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
