"""GPU: parity of the ONE-CALL entries (east_table_host / east_table_dev -- the path bench.py times) at the sizes
the benchmark runs them: ~50 KB documents x 1 000 keyphrases, the pipelined host build, documents at the limits of
the per-document kernel, and the device-side keyphrase preparation against its host variant.  All bit-exact
against the CPU oracle (oracle/east_oracle.c), through the C ABI."""
import numpy as np
import pytest

from conftest import ARRAY_NAMES

pytestmark = pytest.mark.gpu


def _capi():
    from east import _capi
    return _capi


def _concat(packed):
    doc_off = np.zeros(len(packed) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in packed], out=doc_off[1:])
    return np.ascontiguousarray(np.concatenate(packed), dtype=np.uint32), doc_off


def _keyphrases(K, extra=()):
    import synth
    from east import utils
    kps = [utils.prepare_text(k) for k in synth.keyphrases(K)] + list(extra)
    return _capi().pack_keyphrases(kps)


def _table_host(packed, ms, codes, off, normalized=True):
    text, doc_off = _concat(packed)
    out = np.full((len(packed), len(off) - 1), -1.0)
    idx = _capi().DeviceIndex.build_host_and_score(text, doc_off, ms, codes, off, out, normalized)
    return idx, out


def _table_dev(packed, ms, codes, off, normalized=True, host_copy=True):
    import torch
    text, doc_off = _concat(packed)
    text_t = torch.from_numpy(text.view(np.int32)).cuda()
    kp_t = torch.from_numpy(codes.view(np.int32).copy()).cuda()
    out_t = torch.full((len(packed), len(off) - 1), -1.0, dtype=torch.float64, device="cuda")
    idx = _capi().DeviceIndex.build_dev_and_score(text_t.data_ptr(), doc_off, ms, kp_t.data_ptr(), codes if host_copy else None,
                                                  off, out_t.data_ptr(), normalized)
    torch.cuda.synchronize()
    return idx, out_t.cpu().numpy()


def _oracle_rows(oracle_mod, packed, ms, docs, codes, off, normalized=True):
    return {d: oracle_mod.OracleEASA(text=packed[d], m=ms[d]).score_many(codes, off, normalized) for d in docs}


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def _check_arrays(idx, doc, o, tag):
    capi = _capi()
    for which, name in zip(range(6), ARRAY_NAMES):
        got = idx.array(doc, which)
        exp = getattr(o, name)
        assert np.array_equal(got, exp), (tag, name, np.nonzero(got != exp)[0][:5])


def test_one_call_entries_at_benchmark_document_size(oracle_mod):
    # BASELINE configs[1] shape, 8 documents of it: every row of both one-call entries against the oracle
    import synth
    packed, ms, _ = synth.packed_collection(8, 50000, first_seed=1)
    codes, off = _keyphrases(1000)
    exp = _oracle_rows(oracle_mod, packed, ms, range(8), codes, off)
    for normalized in (True, False):
        if not normalized:
            exp = _oracle_rows(oracle_mod, packed, ms, range(8), codes, off, False)
        idx, out = _table_host(packed, ms, codes, off, normalized)
        assert idx.info()["doc_sorted"] and idx.stat("tables_fused") == 1
        for d in range(8):
            assert np.array_equal(_bits(out[d]), _bits(exp[d])), ("host", normalized, d)
        if normalized:
            for d in (0, 7):
                _check_arrays(idx, d, oracle_mod.OracleEASA(text=packed[d], m=ms[d]), ("host", d))
        idx.close()
        idx, out = _table_dev(packed, ms, codes, off, normalized)
        for d in range(8):
            assert np.array_equal(_bits(out[d]), _bits(exp[d])), ("dev", normalized, d)
        idx.close()


def test_pipelined_one_call_at_benchmark_size(oracle_mod):
    # 300 documents x 50 KB: large enough for the pipelined host build (runs of whole documents copied, indexed and
    # scored while the rest is in flight).  Sampled rows against the oracle, every row against the two-call path.
    import synth
    capi = _capi()
    packed, ms, _ = synth.packed_collection(300, 50000, first_seed=1001)
    codes, off = _keyphrases(1000)
    idx, out = _table_host(packed, ms, codes, off)
    assert idx.stat("pipelined") == 1 and idx.stat("pipeline_miss") == 0
    sample = [0, 1, 36, 37, 147, 148, 149, 295, 296, 299] + list(range(60, 300, 47))
    exp = _oracle_rows(oracle_mod, packed, ms, sample, codes, off)
    for d in sample:
        assert np.array_equal(_bits(out[d]), _bits(exp[d])), d
    for d in (36, 37, 299):
        _check_arrays(idx, d, oracle_mod.OracleEASA(text=packed[d], m=ms[d]), ("pipelined", d))
    again = idx.score_table(codes, off, True)     # the batched scorer kernels on the finished index
    assert np.array_equal(_bits(again), _bits(out))
    idx.close()
    idx, out_dev = _table_dev(packed, ms, codes, off, host_copy=False)
    assert np.array_equal(_bits(out_dev), _bits(out))
    idx.close()
    try:   # the host variant of the keyphrase preparation gives the same table (another visiting order, same values)
        capi.set_option("kp_prep_host", 1)
        idx, out_h = _table_host(packed, ms, codes, off)
        assert np.array_equal(_bits(out_h), _bits(out))
        idx.close()
    finally:
        capi.set_option("kp_prep_host", 0)


def _document_of(rng, target):
    """A strings collection whose packed form has exactly `target` code points."""
    words = ["".join(rng.choice(list("ABCDEFGHIJKLMNOPQRSTUVWXYZ"), size=int(rng.integers(3, 9)))) for _ in range(2000)]
    p = 1.0 / np.arange(1, len(words) + 1)
    p /= p.sum()
    strings, total = [], 0
    while total < target:
        s_ = "".join(rng.choice(words, size=3, p=p))
        if total + len(s_) + 1 > target:
            s_ = s_[: max(1, target - total - 1)]
        strings.append(s_)
        total += len(s_) + 1
    if total != target:
        strings[0] = strings[0][: len(strings[0]) - (total - target)] or "A"
    return strings


def test_documents_at_the_kernel_limits_are_indexed_and_scored_in_one_call(oracle_mod):
    # documents of 58 000 / 64 800 / 65 535 code points: the scorer phase of the per-document kernel stages text and
    # suffix array in what the sort left of the shared memory; the fused table phases stop fitting near 64.9 k
    from east.asts import utils as au
    rng = np.random.default_rng(5)
    codes, off = _keyphrases(300, extra=["QZX", "E", "A" * 40])
    for group in ((58000, 1200), (64800, 30000), (65535, 64800, 58000, 700)):
        cols = [_document_of(rng, t) for t in group]
        packed = [au.pack_strings_collection(c) for c in cols]
        assert [p.size for p in packed] == list(group)
        ms = [len(c) for c in cols]
        exp = _oracle_rows(oracle_mod, packed, ms, range(len(group)), codes, off)
        for fn in (_table_host, _table_dev):
            idx, out = fn(packed, ms, codes, off)
            assert idx.info()["doc_sorted"]
            for d in range(len(group)):
                assert np.array_equal(_bits(out[d]), _bits(exp[d])), (fn.__name__, group, d)
                _check_arrays(idx, d, oracle_mod.OracleEASA(text=packed[d], m=ms[d]), (fn.__name__, group, d))
            idx.close()


def test_keyphrase_preparation_on_the_device_equals_the_host_variant(oracle_mod):
    # duplicates, shared tails, one-symbol and long keyphrases, code points outside the alphabet and in the terminator
    # range: the table must not depend on where (or in which order) the distinct suffixes were found
    import synth
    capi = _capi()
    packed, ms, _ = synth.packed_collection(5, 4000, first_seed=31)
    extra = ["E", "E", "THE", "THETHE", "ATHE", "中文A", "A中", "਀", "ABਁC", "Q" * 300, "Q" * 299, "ABCDEFGHIJKLMNOPQRSTUVWXYZ" * 9,
             "ABCDEFGHIJKLMNOPQRSTUVWXYZ" * 9 + "A", "XYZABCDEFGHIJ", "WXYZABCDEFGHIJ", "ZABCDEFGHIJK"]
    for K in (1, 7, 1000, 4000):   # 4000: ~60 thousand suffixes, just below the one-kernel limit
        codes, off = _keyphrases(K, extra=extra if K > 1 else ())
        exp = _oracle_rows(oracle_mod, packed, ms, range(5), codes, off)
        # the device preparation in one CTA (few suffixes), as the chain of kernels (many), and the host variant
        for host_prep, small_max in ((0, 0), (0, 1), (1, 0)):
            try:
                capi.set_option("kp_prep_host", host_prep)
                capi.set_option("kp_small_max", small_max)
                idx, out = _table_host(packed, ms, codes, off)
                two = idx.score_table(codes, off, True)
                idx.close()
            finally:
                capi.set_option("kp_prep_host", 0)
                capi.set_option("kp_small_max", 0)
            for d in range(5):
                assert np.array_equal(_bits(out[d]), _bits(exp[d])), (K, host_prep, small_max, d)
            assert np.array_equal(_bits(two), _bits(out)), (K, host_prep, small_max)
    # per-suffix results (return_suffix_scores): every suffix is its own group, in suffix order
    from east.asts import base
    ast = base.AST.get_ast(["XABXAC", "HI"])
    s, per = ast.score("ABCIAB", return_suffix_scores=True)
    assert list(per.keys()) == ["ABCIAB", "BCIAB", "CIAB", "IAB", "AB", "B"]


def test_one_byte_text_entries_equal_the_uint32_entries(oracle_mod):
    # east_*_host_u8: the text crosses the host link as one byte per code point (0xFF = end of a string); the
    # per-document kernel byte-codes from it and restores the code points (terminators 0x0A00 + k) itself
    import synth
    from east.asts import utils as au
    capi = _capi()
    packed, ms, cols = synth.packed_collection(320, 10000, first_seed=501)
    packed8 = [au.pack_strings_collection_u8(c) for c in cols]
    assert all(p8 is not None and p8.size == p.size for p8, p in zip(packed8, packed))
    codes, off = _keyphrases(200, extra=["ABਁC", "਀", "E", "Zਁ"])   # code points in the terminator range: the generic
    text8, doc_off = _concat(packed8)                                 # walk reads the restored uint32 text
    text8 = np.ascontiguousarray(text8, dtype=np.uint8)

    def table8(t8, d_off, d_m):
        out = np.full((len(d_m), len(off) - 1), -1.0)
        return capi.DeviceIndex.build_host_and_score(t8, d_off, d_m, codes, off, out, True), out

    idx32, exp = _table_host(packed, ms, codes, off)
    idx8, out = table8(text8, doc_off, ms)
    assert idx8.stat("pipelined") == 1 and idx8.info()["doc_sorted"]
    assert np.array_equal(_bits(out), _bits(exp))
    for d in (0, 36, 37, 148, 319):
        o = oracle_mod.OracleEASA(text=packed[d], m=ms[d])
        _check_arrays(idx8, d, o, ("u8", d))
        assert np.array_equal(_bits(out[d]), _bits(o.score_many(codes, off, True)))
    assert np.array_equal(_bits(idx8.score_table(codes, off, True)), _bits(exp))   # the index that comes back is complete
    idx8.close(); idx32.close()
    # build only
    idx8 = capi.DeviceIndex.build_host_u8(text8, doc_off, ms)
    assert idx8.stat("pipelined") == 1
    assert np.array_equal(_bits(idx8.score_table(codes, off, True)), _bits(exp))
    idx8.close()
    # a batch too small for the pipelined build, and one with a document the per-document kernel cannot take: expanded
    # on the device, then the ordinary build
    t5, o5 = _concat(packed8[:5])
    idx8, out5 = table8(np.ascontiguousarray(t5, dtype=np.uint8), o5, ms[:5])
    assert idx8.stat("pipelined") == 0
    assert np.array_equal(_bits(out5), _bits(exp[:5]))
    idx8.close()
    big, big_m, big_cols = synth.packed_collection(1, 90000, first_seed=77)
    mix8 = packed8[:3] + [au.pack_strings_collection_u8(big_cols[0])]
    tm, om = _concat(mix8)
    idx8, outm = table8(np.ascontiguousarray(tm, dtype=np.uint8), om, ms[:3] + big_m)
    assert not idx8.info()["doc_sorted"]
    o = oracle_mod.OracleEASA(text=big[0], m=big_m[0])
    assert np.array_equal(_bits(outm[3]), _bits(o.score_many(codes, off, True)))
    _check_arrays(idx8, 3, o, "u8 big")
    idx8.close()
    # a symbol that first appears in a late run: the speculation fails, the batch is expanded and rebuilt
    late_col = ["0123456789 QUIZ", "ZEBRA9"]
    late8, late32 = au.pack_strings_collection_u8(late_col), au.pack_strings_collection(late_col)
    tl, ol = _concat(packed8 + [late8])
    idx8, outl = table8(np.ascontiguousarray(tl, dtype=np.uint8), ol, ms + [2])
    assert idx8.stat("pipeline_miss") == 1 and idx8.stat("pipelined") == 0
    assert np.array_equal(_bits(outl[:320]), _bits(exp))
    assert np.array_equal(_bits(outl[320]), _bits(oracle_mod.OracleEASA(text=late32, m=2).score_many(codes, off, True)))
    idx8.close()
    # malformed: a document that does not end with 0xFF / holds another number of string ends than doc_m says
    broken = text8.copy()
    broken[doc_off[200 + 1] - 1] = ord("A")
    with pytest.raises(ValueError):
        table8(broken, doc_off, ms)
    with pytest.raises(ValueError):
        table8(np.ascontiguousarray(t5, dtype=np.uint8), o5, [m_ + 1 for m_ in ms[:5]])
    # the Python layer takes the narrow form on its own when the text allows it
    assert au.pack_strings_collection_u8(["ÿ"]) is None and au.pack_strings_collection_u8(["Я"]) is None


def test_saved_index_gives_identical_arrays_and_scores(tmp_path, oracle_mod):
    # index persistence -- packed text, suffix array, LCP, child table, annotation and the scorer's
    # side tables of a device batch go to a file and come back bit-identical; the measure saves / loads a collection
    import synth
    from east import applications, relevance, utils
    capi = _capi()
    packed, ms, cols = synth.packed_collection(6, 3000, first_seed=900)
    codes, off = _keyphrases(40, extra=["QZX", "E"])
    for opt in (0, 1):   # per-document kernel / global sort
        try:
            capi.set_option("no_doc_sort", opt)
            idx = capi.DeviceIndex(packed, ms)
        finally:
            capi.set_option("no_doc_sort", 0)
        exp = idx.score_table(codes, off, True)
        path = str(tmp_path / ("idx%d.eastidx" % opt))
        idx.save(path)
        again = capi.DeviceIndex.load(path)
        assert again.n_docs == 6 and list(again.doc_m) == list(ms)
        for d in range(6):
            for which in range(7):
                assert np.array_equal(idx.array(d, which), again.array(d, which)), (opt, d, which)
            assert again.strings_collection(d) == cols[d]
        for normalized in (True, False):
            assert np.array_equal(_bits(again.score_table(codes, off, normalized)), _bits(idx.score_table(codes, off, normalized)))
        o = oracle_mod.OracleEASA(text=packed[3], m=ms[3])
        assert np.array_equal(_bits(exp[3]), _bits(o.score_many(codes, off, True)))
        idx.close(); again.close()
    with pytest.raises(ValueError):
        capi.DeviceIndex.load(str(tmp_path / "missing.eastidx"))
    bad = tmp_path / "bad.eastidx"
    bad.write_bytes(b"not an index")
    with pytest.raises(ValueError):
        capi.DeviceIndex.load(str(bad))
    # through the measure: a mixed collection (two device batches), saved and loaded
    docs = synth.documents(4, 3000, first_seed=300)
    docs.insert(1, synth.documents(1, 90000, first_seed=400)[0])
    texts = {"t%d" % i: d for i, d in enumerate(docs)}
    kps = synth.keyphrases(10)
    measure = relevance.ASTRelevanceMeasure("easa", True)
    table = applications.keyphrases_table(kps, texts, measure)
    measure.save_index(str(tmp_path / "collection"))
    loaded = relevance.ASTRelevanceMeasure.load_index(str(tmp_path / "collection"))
    prepared = [utils.prepare_text(k) for k in kps]
    t2 = loaded.relevance_table(prepared)
    for k, kp in enumerate(kps):
        for j, name in enumerate(texts):
            assert float(t2[j, k]).hex() == float(table[kp][name]).hex()
    assert loaded.relevance(prepared[0], text=1) == measure.relevance(prepared[0], text=1)
    assert loaded.asts[2].string == measure.asts[2].string
    assert np.array_equal(loaded.asts[1].suftab, measure.asts[1].suftab)


def test_synonym_expanded_scoring_is_the_best_variant(oracle_mod):
    # easa.py:27-34: with a synonimizer the score is the maximum over all substitutions of the query words by their
    # synonyms (always normalized); variants are generated on the host, each scored on the device
    import collections
    import itertools
    from east import utils
    from east.asts import base

    class Synonimizer(object):
        def get_synonyms(self):
            d = collections.defaultdict(list)
            d.update({"QUICK": ["FAST", "SPEEDY"], "FOX": ["VIXEN"], "LAZY": ["IDLE"]})
            return d

    strings = utils.text_to_strings_collection("the speedy brown vixen jumps over the idle dog while a fast fox sleeps")
    ast = base.AST.get_ast(strings)
    o = oracle_mod.OracleEASA(strings)
    for query in ("QUICK FOX", "LAZY DOG", "QUICK BROWN FOX JUMPS", "NOTHING HERE"):
        words = utils.tokenize(query)
        syn = Synonimizer().get_synonyms()
        variants = ["".join(w) for w in itertools.product(*[syn[x] + [x] for x in words])]
        from east.asts.utils import codepoints
        best = max(float(o.score_many(np.ascontiguousarray(codepoints(v), dtype=np.uint32), np.array([0, len(v)], dtype=np.int64), True)[0])
                   for v in variants)
        got = ast.score(query, synonimizer=Synonimizer())
        assert float(got).hex() == float(best).hex(), query
        assert float(got) >= float(ast.score(query))


def _host_packed(texts):
    from east import utils
    from east.asts import utils as au
    cols = [utils.text_to_strings_collection(t) for t in texts]
    return cols, [au.pack_strings_collection(c) for c in cols]


def test_device_preprocessing_equals_text_to_strings_collection(golden):
    # utils.text_to_strings_collection + packing on the device (csrc/tokenize.cu) for ASCII and
    # Cyrillic texts: the packed documents must equal the host path's code point for code point -- on the reference's
    # own preprocessing goldens, on hand-made edge cases and on random texts
    import synth
    capi = _capi()
    texts = [c["in"] for c in golden["prep"]["collections"]]
    packed, doc_off, doc_m = capi.texts_to_packed(texts)
    for d, c in enumerate(golden["prep"]["collections"]):
        seg = packed[doc_off[d]:doc_off[d + 1]]
        strs = "".join(chr(x) if x < 0x0A00 else "\n" for x in seg).split("\n")[:-1]
        assert strs == c["out"], (c["in"], strs)
        assert doc_m[d] == len(c["out"])
    edge = ["", " ", "a", "ab", "abc", "abc ", " abc", "ab cd ef", "123 4567 89", "12a 34' ''' ___ _1_", "x" * 5000, "7" * 5000 + "x",
            "7" * 5000, "ab " * 3000, "abc " * 3000, "one two", "one two three", "one two three four",
            "Tabs\tand\nnew\r\nlines\x00nul\x7fdel", "it's o'clock 'quoted' rock'n'roll", "UPPER lower MiXeD",
            "привет мир ёжик Ёлка ѐѝџ ЀЍЏ як", "mixed смесь text текст 123 ab аб абв", "ёё ё ёёё", "q" * 65000,
            synth.document(50000, 5), synth.document(777, 6).replace(" ", ", "), "a" * 511 + " " + "b" * 513 + " cc dd eee"]
    rng = np.random.default_rng(11)
    alphabet = list("abcXYZ019_' .,;-\n\t") + list("яЖё")
    for _ in range(300):
        n = int(rng.integers(0, 400)) if rng.random() < 0.8 else int(rng.integers(400, 6000))
        edge.append("".join(rng.choice(alphabet, size=n)))
    cols, exp = _host_packed(edge)
    packed, doc_off, doc_m = capi.texts_to_packed(edge)
    for d in range(len(edge)):
        got = packed[doc_off[d]:doc_off[d + 1]]
        assert np.array_equal(got, exp[d]), (d, edge[d][:60], got[:20], exp[d][:20])
        assert doc_m[d] == len(cols[d])
    # every code point the device accepts, against Python's own tables: between letters, doubled, next to digits, at both ends
    from east import relevance
    accepted = [c for c in list(range(0x80)) + list(range(0x80, 0xC0)) + list(range(0x400, 0x460)) + list(range(0x2000, 0x2070)) +
                [0x2116, 0xFEFF] if relevance._DEVICE_TEXT_RE.match(chr(c))]
    assert len(accepted) == 128 + 55 + 96 + 112 + 2
    per_cp = ["%sab%scd%s%sxyz 12%s3 %sэюя%sABC%s" % ((chr(c),) * 8) for c in accepted]
    cols_cp, exp_cp = _host_packed(per_cp)
    packed_cp, off_cp, m_cp = capi.texts_to_packed(per_cp)
    for d, c in enumerate(accepted):
        assert np.array_equal(packed_cp[off_cp[d]:off_cp[d + 1]], exp_cp[d]), (hex(c), cols_cp[d])
    russian = "«Положение о зачётах» – документ №5… утверждён 12.03.2014 г. (см. п. 3.1–3.4); «ёлки-палки», it's ok"
    cols_r, exp_r = _host_packed([russian])
    packed_r, off_r, m_r = capi.texts_to_packed([russian])
    assert np.array_equal(packed_r, exp_r[0]) and m_r[0] == len(cols_r[0])
    # bytes go in as they are
    p2, o2, m2 = capi.texts_to_packed([t.encode("utf-8") for t in edge[:30]])
    assert np.array_equal(p2, packed[: doc_off[30]]) and np.array_equal(o2, doc_off[:31])
    # what the device does not handle is refused, never approximated
    for bad in ("café au lait", "中文 text", "emoji \U0001F600 here", b"broken \xff\xfe utf8", b"cut \xd0", "Ѡ beyond the block",
                b"stray \x80 continuation", "x² squared", "µm", "½ cup", "ª", b"cut \xe2\x80", b"\xe2\x80\x93\x93 one continuation too many",
                "superscript \u2070", "\u20ac euro", b"overlong \xc0\x80", b"\xc2 alone"):
        with pytest.raises(capi.UnsupportedText):
            capi.texts_to_packed(["fine text here", bad])


def test_keyphrases_table_from_raw_texts_in_one_call(oracle_mod):
    # east_table_texts_host: raw texts in, scores out; and the public API takes that path on its own for ASCII / Cyrillic
    import synth
    from east import applications, relevance, utils
    capi = _capi()
    docs = synth.documents(40, 20000, first_seed=2100) + ["", "12 ab", "привет мир привет ёжик " * 50]
    codes, off = _keyphrases(300, extra=["ПРИВЕТ", "ЁЖИК МИР"])
    out = np.full((len(docs), len(off) - 1), -1.0)
    idx = capi.DeviceIndex.table_from_texts(docs, codes, off, out, True)
    cols, packed = _host_packed(docs)
    for d in (0, 17, 39, 40, 41, 42):
        o = oracle_mod.OracleEASA(text=packed[d], m=len(cols[d]))
        assert np.array_equal(_bits(out[d]), _bits(o.score_many(codes, off, True))), d
        _check_arrays(idx, d, o, ("texts", d))
        assert idx.strings_collection(d) == cols[d]
    idx.close()
    texts = {"t%d" % i: t for i, t in enumerate(docs)}
    kps = synth.keyphrases(25) + ["привет", "ёжик мир"]
    on_device = relevance.ASTRelevanceMeasure("easa", True)
    on_host = relevance.ASTRelevanceMeasure("easa", True, device_preprocessing=False)
    t_dev = applications.keyphrases_table(kps, texts, on_device)
    t_host = applications.keyphrases_table(kps, texts, on_host)
    assert on_device.asts[0]._strings_collection is None    # indexed from raw text: nothing was tokenised on the host
    for kp in kps:
        for name in texts:
            assert float(t_dev[kp][name]).hex() == float(t_host[kp][name]).hex(), (kp, name)
    assert on_device.asts[42].string == on_host.asts[42].string
    assert on_device.relevance(utils.prepare_text(kps[0]), text=3) == on_host.relevance(utils.prepare_text(kps[0]), text=3)
    # a collection the device does not handle falls back to the host preprocessing, silently and exactly
    texts["latin1"] = "café crème brûlée " * 20
    t_mixed = applications.keyphrases_table(kps, texts, relevance.ASTRelevanceMeasure("easa", True))
    for kp in kps:
        assert float(t_mixed[kp]["t5"]).hex() == float(t_host[kp]["t5"]).hex()


def test_reference_sample_corpus_from_raw_text(golden):
    # the reference's own sample corpus (doc/samples: 30 Russian texts x 17 keyphrases), from RAW text: the device
    # preprocessing must give the strings collections the reference produced, and applications.keyphrases_table /
    # keyphrases_graph -- which now take the raw-text engine call on their own -- the reference's tables and graphs
    import json
    import os
    from conftest import GOLDEN_DIR
    from east import applications, relevance
    capi = _capi()
    hse = golden["hse"]
    with open(os.path.join(GOLDEN_DIR, "hse_texts.json"), encoding="utf-8") as f:
        raw = json.load(f)
    names = [d["name"] for d in hse["docs"]]
    packed, doc_off, doc_m = capi.texts_to_packed([raw[n] for n in names])
    for j, d in enumerate(hse["docs"]):
        seg = packed[doc_off[j]:doc_off[j + 1]]
        strs = "".join(chr(x) if x < 0x0A00 else "\n" for x in seg).split("\n")[:-1]
        assert strs == d["strings"], d["name"]
    texts = {n: raw[n] for n in names}
    kps = hse["keyphrases"]
    for normalized, key in ((True, "table_norm"), (False, "table_denorm")):
        measure = relevance.ASTRelevanceMeasure("easa", normalized)
        table = applications.keyphrases_table(kps, texts, measure)
        assert measure.asts[0]._strings_collection is None     # the raw-text path was taken
        for kp in kps:
            for n in names:
                assert float(table[kp][n]).hex() == hse[key][kp][n], (kp, n)
    for g in hse["graphs"]:
        graph = applications.keyphrases_graph(kps, texts, referral_confidence=g["c"], relevance_threshold=g["r"],
                                              support_threshold=g["p"], similarity_measure=relevance.ASTRelevanceMeasure("easa", True))
        assert graph["nodes"] == g["nodes"]
        assert [(e["source"], e["target"], float(e["confidence"]).hex()) for e in graph["edges"]] == \
               [(e["source"], e["target"], e["confidence"]) for e in g["edges"]]


def test_sampled_alphabet_of_device_resident_batches_recovers(oracle_mod):
    # east_table_dev / east_build_dev take the alphabet of a batch of small documents from a PREFIX of the text (2 M code
    # points by default) and let the per-document kernel report a code point the table lacks; a miss redoes the batch with
    # the alphabet of the whole text.  Forced here with a tiny sample.
    import synth
    from east.asts import utils as au
    capi = _capi()
    packed, ms, _ = synth.packed_collection(40, 3000, first_seed=3100)
    codes, off = _keyphrases(60, extra=["QUIZ7", "E"])
    try:
        capi.set_option("alphabet_sample", 2000)
        capi.set_option("no_alphabet_guess", 1)   # this test is about the sample; the guess has its own below
        idx, out = _table_dev(packed, ms, codes, off)
        assert idx.stat("alphabet_miss") == 0 and idx.info()["doc_sorted"]
        exp = _oracle_rows(oracle_mod, packed, ms, (0, 39), codes, off)
        for d in (0, 39):
            assert np.array_equal(_bits(out[d]), _bits(exp[d]))
            _check_arrays(idx, d, oracle_mod.OracleEASA(text=packed[d], m=ms[d]), ("sampled", d))
        idx.close()
        # digits and a rare letter first appear in the last document, far behind the sample
        late = au.pack_strings_collection(["0123456789 QUIZ", "ZEBRA9 J"])
        packed2, ms2 = list(packed) + [late], list(ms) + [2]
        idx, out = _table_dev(packed2, ms2, codes, off)
        assert idx.stat("alphabet_miss") == 1 and idx.info()["doc_sorted"]
        exp = _oracle_rows(oracle_mod, packed2, ms2, (0, 40), codes, off)
        for d in (0, 40):
            assert np.array_equal(_bits(out[d]), _bits(exp[d])), d
            _check_arrays(idx, d, oracle_mod.OracleEASA(text=packed2[d], m=ms2[d]), ("miss", d))
        assert np.array_equal(_bits(idx.score_table(codes, off, True)), _bits(out))
        idx.close()
        # a text that collides with the terminator range behind the sample: the general path takes over
        weird = au.pack_strings_collection(["中文", "AB"])
        idx, out = _table_dev(list(packed) + [weird], list(ms) + [2], codes, off)
        assert not idx.info()["fast_path"]
        o = oracle_mod.OracleEASA(text=weird, m=2)
        assert np.array_equal(_bits(out[40]), _bits(o.score_many(codes, off, True)))
        idx.close()
    finally:
        capi.set_option("alphabet_sample", 0)
        capi.set_option("no_alphabet_guess", 0)


def test_alphabet_of_the_previous_batch_is_only_a_guess(oracle_mod):
    # A batch of small documents may start from the alphabet of the thread's previous batch instead of a scan and its host
    # round trip; the per-document kernel reports every code point the guess lacks and the batch is redone from a scan.
    # Hit, miss (new symbols anywhere), a narrower alphabet after a wider one, option off -- pipelined host build and
    # device-resident build; the tables never depend on it.
    import synth
    from east.asts import utils as au
    capi = _capi()
    codes, off = _keyphrases(200, extra=["QUIZ7", "E", "Я"])
    packed, ms, cols = synth.packed_collection(300, 30000, first_seed=4100)
    digits = au.pack_strings_collection(["0123456789 QUIZ7", "ZEBRA9 J"])
    cyr = au.pack_strings_collection(["ЯБЛОКО И ГРУША", "ABC"])
    with_digits = (list(packed[:200]) + [digits] + list(packed[200:]), list(ms[:200]) + [2] + list(ms[200:]))
    with_cyr = (list(packed) + [cyr], list(ms) + [2])

    def check(idx, out, batch, rows, what):
        exp = _oracle_rows(oracle_mod, batch[0], batch[1], rows, codes, off)
        for d in rows:
            assert np.array_equal(_bits(out[d]), _bits(exp[d])), (what, d)
        _check_arrays(idx, rows[-1], oracle_mod.OracleEASA(text=batch[0][rows[-1]], m=batch[1][rows[-1]]), what)

    try:
        capi.set_option("alphabet_sample", 100000)   # device-resident batches above this size sample / guess
        for fn in (_table_host, _table_dev):
            name = fn.__name__
            capi.set_option("no_alphabet_guess", 1)
            idx, ref = fn(packed, ms, codes, off)
            assert idx.stat("alphabet_guessed") == 0
            idx.close()
            capi.set_option("no_alphabet_guess", 0)
            idx, out = fn(packed, ms, codes, off)          # whatever the guess was (an earlier test's batch), the table stands
            assert np.array_equal(_bits(out), _bits(ref)), name
            idx.close()
            idx, out = fn(packed, ms, codes, off)          # now the guess is this batch's own alphabet: a hit
            assert idx.stat("alphabet_guessed") == 1 and idx.stat("alphabet_miss") == 0 and idx.stat("pipeline_miss") == 0, name
            assert np.array_equal(_bits(out), _bits(ref)), name
            check(idx, out, (packed, ms), (0, 299), (name, "hit"))
            idx.close()
            for batch, rows, what in ((with_digits, (0, 200, 300), "digits in the middle"), (with_cyr, (5, 300), "cyrillic at the end")):
                idx, out = fn(batch[0], batch[1], codes, off)
                assert idx.stat("alphabet_guessed") + idx.stat("pipeline_miss") + idx.stat("alphabet_miss") >= 1, (name, what)
                assert idx.stat("pipeline_miss") == 1 or idx.stat("alphabet_miss") == 1, (name, what)   # the guess lacked them
                check(idx, out, batch, rows, (name, what))
                idx.close()
                idx, out2 = fn(batch[0], batch[1], codes, off)   # the redone batch left its own alphabet behind: a hit
                assert idx.stat("alphabet_guessed") == 1 and idx.stat("alphabet_miss") == 0 and idx.stat("pipeline_miss") == 0, (name, what)
                assert np.array_equal(_bits(out2), _bits(out)), (name, what)
                idx.close()
            weird = au.pack_strings_collection(["中文", "AB"])   # collides with the terminator range: the general path
            idx, out = fn(list(packed) + [weird], list(ms) + [2], codes, off)
            assert not idx.info()["fast_path"], name
            assert np.array_equal(_bits(out[300]), _bits(oracle_mod.OracleEASA(text=weird, m=2).score_many(codes, off, True))), name
            assert np.array_equal(_bits(out[:300]), _bits(ref)), name
            idx.close()
            idx, out = fn(packed, ms, codes, off)          # narrower than the guess: codes the batch never uses cost nothing
            assert idx.stat("alphabet_miss") == 0 and idx.stat("pipeline_miss") == 0, name
            assert np.array_equal(_bits(out), _bits(ref)), name
            check(idx, out, (packed, ms), (17, 299), (name, "narrower"))
            idx.close()
    finally:
        capi.set_option("alphabet_sample", 0)
        capi.set_option("no_alphabet_guess", 0)
