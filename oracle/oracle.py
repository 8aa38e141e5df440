"""ctypes front-end of the CPU oracle (oracle/east_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of east_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg; the product
package (`east`) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
SRC_PATH = os.path.join(HERE, "east_oracle.c")
TERMINATOR_BASE = 0x0A00

_lib = None


def build(force=False):
    """gcc -O2 -shared east_oracle.c -> oracle/liboracle.so (git-ignored, travels with gpurun)."""
    if (not force and os.path.exists(LIB_PATH)
            and (not os.path.exists(SRC_PATH) or os.path.getmtime(LIB_PATH) >= os.path.getmtime(SRC_PATH))):
        return LIB_PATH
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-o", LIB_PATH, SRC_PATH])
    return LIB_PATH


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        u32p, i32p, i64p, f64p = (ctypes.POINTER(t) for t in
                                  (ctypes.c_uint32, ctypes.c_int32, ctypes.c_int64, ctypes.c_double))
        L.oracle_pack.restype = ctypes.c_int64
        L.oracle_pack.argtypes = [u32p, i64p, ctypes.c_int32, u32p]
        L.oracle_build.restype = ctypes.c_int
        L.oracle_build.argtypes = [u32p, ctypes.c_int32, ctypes.c_int32] + [i32p] * 6
        L.oracle_suffix_array.restype = ctypes.c_int
        L.oracle_suffix_array.argtypes = [u32p, ctypes.c_int32, i32p]
        L.oracle_lcp.restype = ctypes.c_int
        L.oracle_lcp.argtypes = [u32p, ctypes.c_int32, i32p, i32p]
        L.oracle_score.restype = ctypes.c_double
        L.oracle_score.argtypes = ([u32p, ctypes.c_int32, ctypes.c_int32] + [i32p] * 6 +
                                   [u32p, ctypes.c_int32, ctypes.c_int, f64p, i64p,
                                    ctypes.POINTER(ctypes.c_int)])
        L.oracle_score_many.restype = ctypes.c_int
        L.oracle_score_many.argtypes = ([u32p, ctypes.c_int32, ctypes.c_int32] + [i32p] * 6 +
                                        [u32p, i64p, ctypes.c_int32, ctypes.c_int, f64p, i64p])
        _lib = L
    return _lib


def codepoints(s):
    """unicode string -> uint32 code points"""
    if not s:
        return np.zeros(0, dtype=np.uint32)
    return np.frombuffer(s.encode("utf-32-le", errors="surrogatepass"), dtype=np.uint32).copy()


def pack(strings_collection):
    """east/asts/utils.py:25-40 + easa.py:19 -> uint32 T with terminators 0x0A00+i."""
    chars = codepoints("".join(strings_collection))
    off = np.zeros(len(strings_collection) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in strings_collection], out=off[1:])
    out = np.zeros(int(off[-1]) + len(strings_collection), dtype=np.uint32)
    if chars.size == 0:
        chars = np.zeros(1, dtype=np.uint32)
    n = lib().oracle_pack(_p(chars, ctypes.c_uint32), _p(off, ctypes.c_int64),
                          len(strings_collection), _p(out, ctypes.c_uint32))
    assert n == out.size
    return out


class OracleEASA(object):
    """CPU restatement of east.asts.easa.EnhancedAnnotatedSuffixArray (easa.py:12-400)."""

    def __init__(self, strings_collection=None, text=None, m=None):
        if text is None:
            text = pack(strings_collection)
            m = len(strings_collection)
        self.text = np.ascontiguousarray(text, dtype=np.uint32)
        self.n = int(self.text.size)
        self.m = int(m)
        n = self.n
        self.suftab = np.zeros(n, dtype=np.int32)
        self.lcptab = np.zeros(n, dtype=np.int32)
        self.childtab_up = np.zeros(n, dtype=np.int32)
        self.childtab_down = np.zeros(n, dtype=np.int32)
        self.childtab_next_l_index = np.zeros(n, dtype=np.int32)
        self.anntab = np.zeros(n, dtype=np.int32)
        rc = lib().oracle_build(_p(self.text, ctypes.c_uint32), n, self.m,
                                *[_p(a, ctypes.c_int32) for a in self._arrays()])
        if rc:
            raise MemoryError("oracle_build failed")
        self.probes = 0

    def _arrays(self):
        return (self.suftab, self.lcptab, self.childtab_up, self.childtab_down,
                self.childtab_next_l_index, self.anntab)

    def score(self, query, normalized=True, return_suffix_scores=False):
        q = codepoints(query.replace(" ", ""))
        L = int(q.size)
        if L == 0:
            raise ZeroDivisionError("float division by zero")
        ss = np.zeros(L, dtype=np.float64)
        probes = ctypes.c_int64(0)
        status = ctypes.c_int(0)
        r = lib().oracle_score(_p(self.text, ctypes.c_uint32), self.n, self.m,
                               *[_p(a, ctypes.c_int32) for a in self._arrays()],
                               _p(q, ctypes.c_uint32), L, int(bool(normalized)),
                               _p(ss, ctypes.c_double), ctypes.byref(probes), ctypes.byref(status))
        self.probes += probes.value
        if status.value == 2:
            raise IndexError("reference _lcp_value reads childtab_up[n]")
        if return_suffix_scores:
            qs = query.replace(" ", "")
            return r, {qs[s:]: float(ss[s]) for s in range(L)}
        return r

    def score_many(self, kp_codes, kp_off, normalized=True):
        """kp_codes: uint32 concatenation of space-stripped keyphrases, kp_off[K+1]."""
        K = len(kp_off) - 1
        out = np.zeros(K, dtype=np.float64)
        probes = ctypes.c_int64(0)
        kp_codes = np.ascontiguousarray(kp_codes, dtype=np.uint32)
        kp_off = np.ascontiguousarray(kp_off, dtype=np.int64)
        lib().oracle_score_many(_p(self.text, ctypes.c_uint32), self.n, self.m,
                                *[_p(a, ctypes.c_int32) for a in self._arrays()],
                                _p(kp_codes, ctypes.c_uint32), _p(kp_off, ctypes.c_int64), K,
                                int(bool(normalized)), _p(out, ctypes.c_double), ctypes.byref(probes))
        self.probes += probes.value
        return out


def interval_score(text, m, sa, query_codes, normalized=True):
    """Second, independent restatement: SA-interval narrowing in pure Python.
    Small cases only.  This is the formulation the CUDA scorer uses; keeping it here lets the
    CPU tests prove A.5 == child-table walk without a GPU."""
    n = len(text)
    L = len(query_codes)
    if L == 0:
        raise ZeroDivisionError
    result = 0
    for s in range(L):
        lo, hi, d, frac, nodes, parent_f = 0, n - 1, 0, 0, 0, n - m
        while s + d < L:
            c = int(query_codes[s + d])
            rows = [r for r in range(lo, hi + 1) if sa[r] + d < n and int(text[sa[r] + d]) == c]
            if not rows:
                break
            lo2, hi2 = rows[0], rows[-1]
            size = hi2 - lo2 + 1
            if d == 0 or size != hi - lo + 1:
                nodes += 1
                frac = frac + float(size) / float(parent_f)
            lo, hi, parent_f, d = lo2, hi2, size, d + 1
        if d > 0:
            r = (frac + d) - nodes
            if normalized:
                r = r / d
            result = result + r
    return result / L
