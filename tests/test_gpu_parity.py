"""GPU: the CUDA path (through the C ABI / the east package) against the oracle and the goldens.

Bit-exact for suftab / lcptab / childtab_* / anntab; scores must be bit-exact too (the
acceptance floor of the task is 1e-9 relative, asserted separately).
"""
import numpy as np
import pytest

from conftest import ARRAY_NAMES

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9  # north_star tolerance for scores


def _capi():
    from east import _capi
    return _capi


@pytest.fixture(autouse=True, params=["doc_sort", "doc_sort_split_tables", "global_sort"])
def sa_path(request):
    """Every test runs three times: with the per-document shared-memory kernel (doc_sort.cu, the default
    for documents of < 64 Ki code points) producing suffix array AND tables; with that kernel producing the
    suffix array only and the batch-wide LCP / child / annotation kernels (tables.cu) after it; and with
    the global prefix-doubling sort (sa_build.cu)."""
    capi = _capi()
    capi.set_option("no_doc_sort", 1 if request.param == "global_sort" else 0)
    capi.set_option("no_fused_tables", 1 if request.param == "doc_sort_split_tables" else 0)
    yield request.param
    capi.set_option("no_doc_sort", 0)
    capi.set_option("no_fused_tables", 0)


def _build(strings_collections, device=0):
    from east.asts import utils
    packed = [utils.pack_strings_collection(c) for c in strings_collections]
    return _capi().DeviceIndex(packed, [len(c) for c in strings_collections], device=device)


def _check_arrays(idx, doc, oracle_ast, tag):
    capi = _capi()
    for which, name in zip((capi.SUFTAB, capi.LCPTAB, capi.CHILDTAB_UP, capi.CHILDTAB_DOWN,
                            capi.CHILDTAB_NEXT_L_INDEX, capi.ANNTAB), ARRAY_NAMES):
        got = idx.array(doc, which)
        exp = getattr(oracle_ast, name)
        assert np.array_equal(got, exp), (tag, name, np.nonzero(got != exp)[0][:5])


def test_golden_arrays_and_scores(golden):
    import east  # noqa: F401
    from east.asts import base
    for c in golden["cases"]:
        ast = base.AST.get_ast(c["strings"], "easa")
        if c.get("arrays"):
            for name in ARRAY_NAMES:
                got = getattr(ast, name)
                assert got.dtype == np.int64
                assert np.array_equal(got, golden["arrays"]["%s/%s" % (c["name"], name)]), (c["name"], name)
        for q in c["queries"]:
            if q.get("raises"):
                with pytest.raises(ZeroDivisionError):
                    ast.score(q["q"])
                continue
            for normalized, key, skey in ((True, "norm", "suffix_norm"), (False, "denorm", "suffix_denorm")):
                exp = float.fromhex(q[key])
                got = ast.score(q["q"], normalized=normalized)
                assert abs(got - exp) <= REL_TOL * abs(exp)
                assert float(got).hex() == q[key], (c["name"], q["q"], normalized)
                got2, per_suffix = ast.score(q["q"], normalized=normalized, return_suffix_scores=True)
                assert float(got2).hex() == q[key]
                assert {k: float(v).hex() for k, v in per_suffix.items()} == q[skey]


def test_readme_known_answer():
    import east  # noqa: F401
    from east.asts import base
    ast = base.AST.get_ast(["XABXAC", "HI"])
    assert ast.score("ABCI") == 0.1875
    assert ast.score("NOPE") == 0
    assert ast.string == "XABXAC਀HIਁ"
    assert ast.suftab.tolist() == [1, 4, 2, 5, 7, 8, 0, 3, 6, 9]
    assert ast.anntab.tolist() == [8, 2, 0, 0, 0, 0, 0, 2, 0, 0]


def test_reference_cross_engine_fixture():
    # tests/asts/test_base.py:13-24: the GPU engine must give the value all three reference engines share
    import east  # noqa: F401
    from east.asts import base
    ast = base.AST.get_ast(["abcd efg ops", "xyzq", "test"])
    assert float(ast.score("aqcb")).hex() == "0x1.99999999999a0p-5"
    assert float(ast.score("aqcb", normalized=False)).hex() == "0x1.99999999999a0p-5"
    assert float(ast.score("efgp")).hex() == "0x1.2888888888888p-2"
    assert ast.score("efgp", normalized=False) == 0.6875
    assert ast.score("mn4") == 0


@pytest.mark.parametrize("force_general", [0, 1])
def test_random_collections_vs_oracle(oracle_mod, force_general):
    capi = _capi()
    capi.set_option("force_general", force_general)
    try:
        rng = np.random.default_rng(11 + force_general)
        for trial in range(40):
            sigma = int(rng.choice([2, 3, 4, 7, 27]))
            alpha = list("ABCDEFGHIJKLMNOPQRSTUVWXYZ ")[:sigma]
            n_docs = int(rng.integers(1, 6))
            cols = []
            for _ in range(n_docs):
                m = int(rng.integers(1, 12))
                cols.append(["".join(rng.choice(alpha, size=int(rng.integers(1, 40)))) for _ in range(m)])
            idx = _build(cols)
            assert idx.info()["fast_path"] == (not force_general)
            oracles = [oracle_mod.OracleEASA(c) for c in cols]
            for d, o in enumerate(oracles):
                _check_arrays(idx, d, o, (trial, d))
            letters = [a for a in alpha if a != " "] or ["A"]
            queries = ["".join(rng.choice(letters, size=int(rng.integers(1, 12)))) for _ in range(9)]
            codes, off = capi.pack_keyphrases(queries)
            for normalized in (True, False):
                table = idx.score_table(codes, off, normalized)
                assert table.shape == (n_docs, len(queries))
                for d, o in enumerate(oracles):
                    exp = o.score_many(codes, off, normalized)
                    assert np.array_equal(table[d].view(np.uint64), exp.view(np.uint64)), (trial, d, normalized)
            idx.close()
    finally:
        capi.set_option("force_general", 0)


def test_zipf_documents_vs_oracle(oracle_mod, sa_path):
    import synth
    capi = _capi()
    packed, ms, cols = synth.packed_collection(6, 10000)
    idx = capi.DeviceIndex(packed, ms)
    info = idx.info()
    assert info["fast_path"] and info["rounds"] <= 6
    assert info["doc_sorted"] == sa_path.startswith("doc_sort")
    assert bool(idx.stat("tables_fused")) == (sa_path == "doc_sort")
    oracles = [oracle_mod.OracleEASA(text=p, m=m) for p, m in zip(packed, ms)]
    for d, o in enumerate(oracles):
        _check_arrays(idx, d, o, d)
    from east import utils
    kps = [utils.prepare_text(k) for k in synth.keyphrases(50)]
    codes, off = capi.pack_keyphrases(kps)
    for normalized in (True, False):
        table = idx.score_table(codes, off, normalized)
        for d, o in enumerate(oracles):
            exp = o.score_many(codes, off, normalized)
            assert np.allclose(table[d], exp, rtol=REL_TOL, atol=0)
            assert np.array_equal(table[d].view(np.uint64), exp.view(np.uint64))


def test_key_window_sizes_give_the_same_arrays(oracle_mod):
    # the round-0 window only changes how many doubling rounds follow, never the result
    import synth
    capi = _capi()
    packed, ms, _ = synth.packed_collection(2, 6000, first_seed=40)
    oracles = [oracle_mod.OracleEASA(text=p, m=m) for p, m in zip(packed, ms)]
    try:
        for kc in (1, 2, 3, 5, 8):
            capi.set_option("key_chars", kc)
            idx = capi.DeviceIndex(packed, ms)
            for d, o in enumerate(oracles):
                _check_arrays(idx, d, o, (kc, d))
            idx.close()
    finally:
        capi.set_option("key_chars", 0)


def test_deep_lcp_and_degenerate_inputs(oracle_mod, sa_path):
    rng = np.random.default_rng(3)
    s = "".join(rng.choice(list("AB"), size=700))
    cols = [[s] * 5,            # analysis/utils.py:5-9 worst case: identical strings
            ["A" * 300],         # one run
            ["AB" * 2500, "B" * 1500],  # groups of thousands of suffixes: doubling rounds fall back to the radix sort
            [" "],               # empty text (utils.py:76-78)
            ["A"], ["AB", "AB", "AB", "B", "A"]]
    idx = _build(cols)
    for d, c in enumerate(cols):
        _check_arrays(idx, d, oracle_mod.OracleEASA(c), d)
    if sa_path == "global_sort":
        assert idx.info()["rounds"] >= 5
    else:
        assert idx.info()["doc_sorted"]
    # [" "].score("AB") == 0
    assert idx.score_one(3, np.array([65, 66], dtype=np.uint32)) == 0.0


def test_unicode_and_terminator_range_collisions(oracle_mod):
    # CJK / Gurmukhi code points sort among or above the terminators 0x0A00+i: general path
    cols = [["中中", "文中文", "ਅਆਇ"], ["ਁਂ", "A"], ["\U0001F600\U0001F600!", "ok"]]
    idx = _build(cols)
    assert not idx.info()["fast_path"]
    for d, c in enumerate(cols):
        o = oracle_mod.OracleEASA(c)
        capi = _capi()
        for which, name in ((capi.SUFTAB, "suftab"), (capi.LCPTAB, "lcptab")):
            assert np.array_equal(idx.array(d, which), getattr(o, name)), (d, name)
    # Cyrillic stays on the fast path
    idx2 = _build([["ЖУК", "ЖУРНАЛ", "ЁЖ"]])
    assert idx2.info()["fast_path"]
    _check_arrays(idx2, 0, oracle_mod.OracleEASA(["ЖУК", "ЖУРНАЛ", "ЁЖ"]), "cyr")


def test_hse_table_and_graph(golden):
    from east import applications, relevance
    hse = golden["hse"]
    texts = {d["name"]: " ".join(d["strings"]) for d in hse["docs"]}
    # rebuild texts whose strings collection equals the golden one: join 3-word groups is lossy, so
    # feed the collections straight to the measure instead
    measure = relevance.ASTRelevanceMeasure("easa", normalized=True)
    from east import _capi
    from east.asts import utils as au
    cols = [d["strings"] for d in hse["docs"]]
    idx = _capi.DeviceIndex([au.pack_strings_collection(c) for c in cols], [len(c) for c in cols])
    kps = hse["keyphrases"]
    from east import utils
    codes, off = _capi.pack_keyphrases([utils.prepare_text(k) for k in kps])
    for normalized, key in ((True, "table_norm"), (False, "table_denorm")):
        table = idx.score_table(codes, off, normalized)
        for k, kp in enumerate(kps):
            for j, d in enumerate(hse["docs"]):
                assert float(table[j, k]).hex() == hse[key][kp][d["name"]], (kp, d["name"])
    # graph: co-occurrence counts on device, edges/confidences as in applications.py:111-147
    table = idx.score_table(codes, off, True)
    for g in hse["graphs"]:
        cooc = _capi.cooc_host(table, g["r"])
        B = table >= g["r"]
        assert np.array_equal(cooc, (B.T.astype(np.int64) @ B.astype(np.int64)).astype(np.int32))
        support = np.diag(cooc)
        assert [n["support"] for n in g["nodes"]] == [int(support[n["id"]]) for n in g["nodes"]]
        # nodes and EVERY edge (order, endpoints, confidence bit for bit) as the reference built them
        graph = applications.graph_from_cooccurrence(kps, kps, cooc, g["c"], g["r"], g["p"])
        assert graph["nodes"] == g["nodes"]
        assert [(e["source"], e["target"], float(e["confidence"]).hex()) for e in graph["edges"]] == \
               [(e["source"], e["target"], e["confidence"]) for e in g["edges"]]
    del measure
    # the same through the public API (applications.keyphrases_graph -> ASTRelevanceMeasure -> one-call engine entry):
    # the goldens hold the strings collections, so the preprocessing step hands them out for marker texts
    from east import utils as east_utils
    real = east_utils.text_to_strings_collection
    try:
        east_utils.text_to_strings_collection = lambda text: cols[int(text[2:])]
        marker_texts = {d["name"]: "@@%d" % j for j, d in enumerate(hse["docs"])}
        for g in hse["graphs"]:
            graph = applications.keyphrases_graph(kps, marker_texts, referral_confidence=g["c"], relevance_threshold=g["r"],
                                                  support_threshold=g["p"],
                                                  similarity_measure=relevance.ASTRelevanceMeasure("easa", normalized=True,
                                                                                                   device_preprocessing=False))
            assert graph["nodes"] == g["nodes"], (g["c"], g["r"], g["p"])
            assert [(e["source"], e["target"], float(e["confidence"]).hex()) for e in graph["edges"]] == \
                   [(e["source"], e["target"], e["confidence"]) for e in g["edges"]]
        tab = applications.keyphrases_table(kps, marker_texts, relevance.ASTRelevanceMeasure("easa", normalized=True,
                                                                                             device_preprocessing=False))
        for kp in kps:
            for d in hse["docs"]:
                assert float(tab[kp][d["name"]]).hex() == hse["table_norm"][kp][d["name"]]
    finally:
        east_utils.text_to_strings_collection = real


def test_applications_api_end_to_end(golden, oracle_mod):
    from east import applications, relevance, utils
    import synth
    docs = synth.documents(5, 4000, first_seed=100)
    texts = {"doc%d.txt" % i: d for i, d in enumerate(docs)}
    kps = synth.keyphrases(20) + [""]
    table = applications.keyphrases_table(kps, texts, relevance.ASTRelevanceMeasure("easa", normalized=True))
    assert "" not in table and set(table.keys()) == set(k for k in kps if k)
    for name, text in texts.items():
        o = oracle_mod.OracleEASA(utils.text_to_strings_collection(text))
        for kp in kps:
            if kp:
                exp = o.score(utils.prepare_text(kp), True)
                assert float(table[kp][name]) == float(exp), (kp, name)
    graph = applications.keyphrases_graph([k for k in kps if k], texts, referral_confidence=0.5,
                                          relevance_threshold=0.1, support_threshold=1)
    # the same graph from the table with the reference's set arithmetic (applications.py:111-147)
    import itertools
    keys = [k for k in kps if k]
    kt = {k: set(t for t in texts if table[k][t] >= 0.1) for k in keys}
    nodes = [{"id": i, "label": k, "support": len(kt[k])} for i, k in enumerate(keys) if len(kt[k]) >= 1]
    edges = []
    for a, b in itertools.permutations(range(len(nodes)), 2):
        conf = float(len(kt[nodes[a]["label"]] & kt[nodes[b]["label"]])) / max(len(kt[nodes[a]["label"]]), 1)
        if conf >= 0.5:
            edges.append({"source": nodes[a]["id"], "target": nodes[b]["id"], "confidence": conf})
    assert graph["nodes"] == nodes
    assert graph["edges"] == edges


def test_traversals_cover_all_intervals(oracle_mod):
    import east  # noqa: F401
    from east.asts import base
    ast = base.AST.get_ast(["XABXAC", "HI"])
    seen = []
    ast.traverse(lambda node: seen.append((node[0], node[1], node[2])), "depth-first|post-order")
    assert seen == [(1, 0, 1), (2, 6, 7), (0, 0, 9)]
    pre = []
    ast.traverse(lambda node: pre.append((node[1], node[2])), "depth-first|pre-order")
    assert pre[0] == (0, 9) and (0, 1) in pre and (6, 7) in pre and len(pre) == 1 + 8 + 4
    with pytest.raises(NotImplementedError):
        ast.traverse(lambda node: None, "breadth-first")


def test_error_paths():
    capi = _capi()
    idx = _build([["AB"]])
    with pytest.raises(ZeroDivisionError):
        idx.score_one(0, np.zeros(0, dtype=np.uint32))
    with pytest.raises(ZeroDivisionError):
        codes, off = capi.pack_keyphrases(["AB", " "])
        idx.score_table(codes, off)
    with pytest.raises(ValueError):
        idx.array(5, capi.SUFTAB)
    assert capi.launch_count() > 0


def test_determinism_two_builds_identical():
    import synth
    capi = _capi()
    packed, ms, _ = synth.packed_collection(3, 8000, first_seed=7)
    a, b = capi.DeviceIndex(packed, ms), capi.DeviceIndex(packed, ms)
    for d in range(3):
        for which in range(6):
            assert np.array_equal(a.array(d, which), b.array(d, which))


def _nccl_worker(rank, world, port, out_dir):
    import os
    import sys
    from conftest import PKG, ROOT
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import synth
        from east import distributed, utils
        docs = [synth.document(3000 + 1000 * (j % 4), 300 + j) for j in range(7)]
        kps = [utils.prepare_text(k) for k in synth.keyphrases(9)]
        full = distributed.relevance_table_sharded(docs, kps, True, device=rank)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), full.cpu().numpy())
        plain = distributed.relevance_table_sharded(docs, kps, True, device=rank, fused_gather=False)
        np.save(os.path.join(out_dir, "plain%d.npy" % rank), plain.cpu().numpy())
        # the batched scorer path (scratch too small for in-kernel scoring, many document tiles): the keyphrase sums
        # store the rows to the peers
        from east import _capi
        try:
            _capi.set_option("score_tmp_doubles", 150)
            tiled = distributed.relevance_table_sharded(docs, kps, True, device=rank)
        finally:
            _capi.set_option("score_tmp_doubles", 0)
        np.save(os.path.join(out_dir, "tiled%d.npy" % rank), tiled.cpu().numpy())
        # a ragged collection (one document the per-document kernel cannot take): several device batches per rank
        ragged = docs[:3] + [synth.document(90000, 999)] + docs[3:]
        rag = distributed.relevance_table_sharded(ragged, kps, True, device=rank)
        np.save(os.path.join(out_dir, "ragged%d.npy" % rank), rag.cpu().numpy())
        names = synth.keyphrases(9)
        graph = distributed.keyphrases_graph_sharded(names, {"t%d" % j: d for j, d in enumerate(docs)}, 0.5, 0.1, 1, device=rank)
        import json
        with open(os.path.join(out_dir, "graph%d.json" % rank), "w") as f:
            json.dump(graph, f)
    finally:
        dist.destroy_process_group()


def test_multi_gpu_sharded_table_is_bit_identical(tmp_path):
    """N > 1 (run with gpurun --gpus 2): documents sharded over ranks, one NCCL all-gather; the
    gathered table equals the single-GPU table bit for bit (pure concatenation)."""
    import socket
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import synth
    from east import utils
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    docs = [synth.document(3000 + 1000 * (j % 4), 300 + j) for j in range(7)]
    kps = [utils.prepare_text(k) for k in synth.keyphrases(9)]
    capi = _capi()
    cols = [utils.text_to_strings_collection(d) for d in docs]
    idx = _build(cols)
    codes, off = capi.pack_keyphrases(kps)
    expect = idx.score_table(codes, off, True)
    from east import applications, relevance
    names = synth.keyphrases(9)
    graph1 = applications.keyphrases_graph(names, {"t%d" % j: d for j, d in enumerate(docs)}, 0.5, 0.1, 1,
                                           similarity_measure=relevance.ASTRelevanceMeasure("easa", True))
    ragged = docs[:3] + [synth.document(90000, 999)] + docs[3:]
    big = _build([utils.text_to_strings_collection(ragged[3])])
    expect_ragged = np.concatenate([expect[:3], big.score_table(codes, off, True), expect[3:]])
    import json
    for r in range(2):
        got = np.load(str(tmp_path / ("rank%d.npy" % r)))
        assert np.array_equal(got.view(np.uint64), expect.view(np.uint64))
        got = np.load(str(tmp_path / ("plain%d.npy" % r)))
        assert np.array_equal(got.view(np.uint64), expect.view(np.uint64))
        got = np.load(str(tmp_path / ("tiled%d.npy" % r)))
        assert np.array_equal(got.view(np.uint64), expect.view(np.uint64))
        got = np.load(str(tmp_path / ("ragged%d.npy" % r)))
        assert np.array_equal(got.view(np.uint64), expect_ragged.view(np.uint64))
        with open(str(tmp_path / ("graph%d.json" % r))) as f:
            assert json.load(f) == json.loads(json.dumps(graph1))


@pytest.mark.parametrize("K,D", [(5, 3), (130, 257), (300, 1000), (257, 4100), (700, 900)])
def test_cooccurrence_tensor_core_counts_are_exact(K, D):
    """C = B B^T on tcgen05 (u8 x u8 -> s32 in TMEM) against numpy and against the AND+POPC kernel."""
    capi = _capi()
    rng = np.random.default_rng(K * 1000 + D)
    S = rng.random((D, K))
    S[rng.random((D, K)) < 0.3] = 0.25  # values exactly at the threshold count as present (>=)
    B = (S >= 0.25).astype(np.int64)
    expect = (B.T @ B).astype(np.int32)
    got = capi.cooc_host(S, 0.25)   # pipelined 128 x 256 tcgen05 kernel (cp.async ring, warp-specialised)
    assert np.array_equal(got, expect)
    try:
        for variant in (1, 2):      # AND+POPC cross-check, simple 128 x 128 tcgen05 kernel
            capi.set_option("cooc_variant", variant)
            assert np.array_equal(capi.cooc_host(S, 0.25), expect), variant
    finally:
        capi.set_option("cooc_variant", 0)


def test_traversal_callbacks_match_reference(golden):
    """traverse() pre-/post-order: same nodes, same order, same callback payload as easa.py:38-85."""
    import east  # noqa: F401
    from east.asts import base
    for item in golden["traversals"]:
        ast = base.AST.get_ast(item["strings"], "easa")
        pre, post = [], []
        ast.traverse(lambda node: pre.append([int(node[0]), int(node[1]), int(node[2]), node[3]]),
                     "depth-first|pre-order")
        ast.traverse(lambda node: post.append([int(node[0]), int(node[1]), int(node[2]),
                                               [[int(c[0]), int(c[1]), int(c[2])] for c in node[3]]]),
                     "depth-first|post-order")
        assert pre == item["pre"]
        assert post == item["post"]


def test_tree_engine_names_resolve_to_the_same_scores():
    # tests/asts/test_base.py:16-24: all engine names give equal scores
    import east  # noqa: F401
    from east.asts import base
    strings, queries = ["abcd efg ops", "xyzq", "test"], ["aqcb", "efgp", "mn4"]
    for normalized in (True, False):
        ref = [base.AST.get_ast(strings, "easa").score(q, normalized=normalized) for q in queries]
        for alg in ("ast_linear", "ast_naive"):
            assert [base.AST.get_ast(strings, alg).score(q, normalized=normalized) for q in queries] == ref


def test_cli_table_and_graph(tmp_path, golden):
    """The fixed thin CLI (east/main.py) on a directory of texts: csv/xml table and edges/gml graph."""
    import io
    from east import main as cli
    hse = golden["hse"]
    d = tmp_path / "texts"
    d.mkdir()
    for doc in hse["docs"][:6]:
        (d / doc["name"]).write_text(" ".join(doc["strings"]), encoding="utf-8")
    kp = tmp_path / "kp.txt"
    kp.write_text("\n".join(hse["keyphrases"][:5]) + "\n", encoding="utf-8")
    out = io.StringIO()
    assert cli.main(["-f", "csv", "keyphrases", "table", str(kp), str(d)], out=out) == 0
    lines = out.getvalue().strip().splitlines()
    assert len(lines) == 1 + 6 and lines[0].count(",") == 5
    out = io.StringIO()
    assert cli.main(["keyphrases", "table", str(kp), str(d)], out=out) == 0
    assert out.getvalue().startswith("<table>") and out.getvalue().count("<text name=") == 30
    out = io.StringIO()
    assert cli.main(["-f", "gml", "-r", "0.1", "-c", "0.5", "keyphrases", "graph", str(kp), str(d)], out=out) == 0
    assert out.getvalue().startswith("graph\n[") and "referral_confidence 0.50" in out.getvalue()
    out = io.StringIO()
    assert cli.main(["-s", "cosine", "keyphrases", "table", str(kp), str(d)], out=out) == 1
    assert cli.main(["keyphrases"], out=io.StringIO()) == 1


def test_score_table_tiles_over_documents(oracle_mod):
    """The per-suffix scratch is bounded: scoring in document tiles gives the identical table."""
    import synth
    from east import utils
    capi = _capi()
    packed, ms, _ = synth.packed_collection(9, 3000, first_seed=70)
    idx = capi.DeviceIndex(packed, ms)
    kps = [utils.prepare_text(k) for k in synth.keyphrases(40)]
    codes, off = capi.pack_keyphrases(kps)
    full = idx.score_table(codes, off, True)
    try:
        capi.set_option("score_tmp_doubles", int(off[-1]) * 2)  # two documents per tile
        tiled = idx.score_table(codes, off, True)
    finally:
        capi.set_option("score_tmp_doubles", 0)
    assert np.array_equal(full.view(np.uint64), tiled.view(np.uint64))
    exp = oracle_mod.OracleEASA(text=packed[8], m=ms[8]).score_many(codes, off, True)
    assert np.array_equal(tiled[8].view(np.uint64), exp.view(np.uint64))


def test_segmented_and_global_round0_sort_agree(oracle_mod):
    """Large documents take the per-document (segmented) radix sort; the global sort (document id in
    the sorted bits) must give the same index.  Mixed sizes: tiny documents get tiny tiles."""
    import synth
    from east import utils
    from east.asts import utils as au
    capi = _capi()
    sizes = [30000, 200, 12000, 60, 25000, 9000]
    cols = [utils.text_to_strings_collection(synth.document(sz, 500 + i)) for i, sz in enumerate(sizes)]
    packed = [au.pack_strings_collection(c) for c in cols]
    ms = [len(c) for c in cols]
    oracles = [oracle_mod.OracleEASA(text=p, m=m) for p, m in zip(packed, ms)]
    try:
        for global_sort in (0, 1):
            capi.set_option("global_sort", global_sort)
            idx = capi.DeviceIndex(packed, ms)
            for d, o in enumerate(oracles):
                _check_arrays(idx, d, o, (global_sort, d))
            idx.close()
    finally:
        capi.set_option("global_sort", 0)


def test_doc_sort_bucket_overflow_falls_back_to_the_global_sort(oracle_mod, sa_path):
    # one bucket of > 4096 suffixes (a run of 9000 equal symbols) is more than a group refines in shared
    # memory: the build must notice, redo the batch with the global sort and still be exact
    if not sa_path.startswith("doc_sort"):
        pytest.skip("per-document sort only")
    cols = [["A" * 9000, "AB"], ["XABXAC", "HI"]]
    idx = _build(cols)
    info = idx.info()
    assert info["doc_sort_overflow"] and not info["doc_sorted"]
    for d, c in enumerate(cols):
        _check_arrays(idx, d, oracle_mod.OracleEASA(c), d)
    # buckets of thousands of suffixes that stay below the limit: many refinement levels
    cols = [["A" * 3000, "BA" * 1200], ["C" * 1100], ["AB" * 40] * 60]
    idx = _build(cols)
    assert idx.info()["doc_sorted"]
    for d, c in enumerate(cols):
        _check_arrays(idx, d, oracle_mod.OracleEASA(c), d)


def test_doc_sort_alphabet_sizes(oracle_mod, sa_path):
    # bits per symbol 1..7 select different bucket / key geometries (G, WS) of the per-document sort
    if not sa_path.startswith("doc_sort"):
        pytest.skip("per-document sort only")
    rng = np.random.default_rng(5)
    pool = [chr(c) for c in range(0x21, 0x7f)] + [chr(c) for c in range(0x410, 0x450)]
    for sigma in (1, 2, 3, 6, 7, 8, 14, 15, 16, 30, 31, 32, 62, 63, 64, 120, 126):
        alpha = pool[:sigma]
        cols = []
        for _ in range(3):
            m = int(rng.integers(1, 30))
            cols.append(["".join(rng.choice(alpha, size=int(rng.integers(1, 60)))) for _ in range(m)])
        # repeated phrases: ties beyond the first key word
        cols.append(["".join(rng.choice(alpha, size=50)) * 3] * 4)
        idx = _build(cols)
        assert idx.info()["doc_sorted"], sigma
        for d, c in enumerate(cols):
            _check_arrays(idx, d, oracle_mod.OracleEASA(c), (sigma, d))
        idx.close()


def test_pipelined_host_build_matches_and_recovers(oracle_mod, sa_path):
    # east_build_host copies large batches chunk by chunk and sorts chunk c while chunk c+1 is in flight,
    # trusting the alphabet of chunk 0; a validating scan at the end confirms it or the build is redone
    if sa_path == "global_sort":
        pytest.skip("the pipelined build drives the per-document kernel")
    import synth
    capi = _capi()
    packed, ms, _ = synth.packed_collection(400, 1200, first_seed=70)
    try:
        capi.set_option("pipeline_chunk", 40000)   # 3 runs of 148, 148, 104 documents
        idx = capi.DeviceIndex(packed, ms)
        assert idx.stat("pipelined") == 1 and idx.stat("pipeline_miss") == 0 and idx.info()["doc_sorted"]
        for d in (0, 147, 148, 399):
            _check_arrays(idx, d, oracle_mod.OracleEASA(text=packed[d], m=ms[d]), ("pipelined", d))
        from east import utils
        kps = [utils.prepare_text(k) for k in synth.keyphrases(20)]
        codes, off = capi.pack_keyphrases(kps)
        table = idx.score_table(codes, off, True)
        capi.set_option("no_pipeline", 1)
        ref = capi.DeviceIndex(packed, ms)
        assert ref.stat("pipelined") == 0
        assert np.array_equal(table.view(np.uint64), ref.score_table(codes, off, True).view(np.uint64))
        capi.set_option("no_pipeline", 0)
        # a symbol that first appears in a late chunk: the speculation fails, the result must not change
        from east.asts import utils as au
        late = au.pack_strings_collection(["0123456789 QUIZ", "ZEBRA9"])
        packed2, ms2 = list(packed) + [late], list(ms) + [2]
        idx2 = capi.DeviceIndex(packed2, ms2)
        assert idx2.stat("pipeline_miss") == 1 and idx2.stat("pipelined") == 0
        for d in (0, 400):
            _check_arrays(idx2, d, oracle_mod.OracleEASA(text=packed2[d], m=ms2[d]), ("miss", d))
        # a broken terminator layout in a late chunk: general path, still exact
        weird = au.pack_strings_collection(["中文", "AB"])
        idx3 = capi.DeviceIndex(list(packed) + [weird], list(ms) + [2])
        assert idx3.stat("pipelined") == 0 and not idx3.info()["fast_path"]
        o = oracle_mod.OracleEASA(text=weird, m=2)
        assert np.array_equal(idx3.array(400, capi.SUFTAB), o.suftab)
    finally:
        capi.set_option("pipeline_chunk", 0)
        capi.set_option("no_pipeline", 0)


def test_build_and_score_in_one_call_matches_the_two_calls(oracle_mod, sa_path):
    # east_table_host scores every run of documents as soon as the per-document kernel has sorted it, while later
    # runs are still being copied; whatever happens to the speculation the table equals build + score
    import synth
    from east import utils
    from east.asts import utils as au
    capi = _capi()
    packed, ms, _ = synth.packed_collection(400, 1200, first_seed=170)
    kps = [utils.prepare_text(k) for k in synth.keyphrases(40)] + ["QQQQQ7", "中文A", "A中", "E", "TH"]
    codes, off = capi.pack_keyphrases(kps)

    def both(packed_docs, doc_m, normalized):
        doc_off = np.zeros(len(packed_docs) + 1, dtype=np.int64)
        np.cumsum([len(p) for p in packed_docs], out=doc_off[1:])
        text = np.ascontiguousarray(np.concatenate(packed_docs), dtype=np.uint32)
        out = np.full((len(packed_docs), len(kps)), -1.0)
        idx = capi.DeviceIndex.build_host_and_score(text, doc_off, doc_m, codes, off, out, normalized)
        try:
            capi.set_option("no_pipeline", 1)
            ref = capi.DeviceIndex(packed_docs, doc_m)
            exp = ref.score_table(codes, off, normalized)
        finally:
            capi.set_option("no_pipeline", 0)
        assert np.array_equal(out.view(np.uint64), exp.view(np.uint64))
        # the index that comes back is a complete one
        again = idx.score_table(codes, off, normalized)
        assert np.array_equal(again.view(np.uint64), exp.view(np.uint64))
        # the same with everything resident on the device (east_table_dev), and the two-call device entries
        import torch
        text_t = torch.from_numpy(text.view(np.int32)).cuda()
        kp_t = torch.from_numpy(codes.view(np.int32).copy()).cuda()
        out_t = torch.full((len(packed_docs), len(kps)), -1.0, dtype=torch.float64, device="cuda")
        for host_copy in (codes, None):
            out_t.fill_(-1.0)
            dev = capi.DeviceIndex.build_dev_and_score(text_t.data_ptr(), doc_off, doc_m, kp_t.data_ptr(), host_copy, off,
                                                       out_t.data_ptr(), normalized)
            assert np.array_equal(out_t.cpu().numpy().view(np.uint64), exp.view(np.uint64))
        out_t.fill_(-1.0)
        dev.score_table_dev(kp_t.data_ptr(), off, out_t.data_ptr(), normalized)
        assert np.array_equal(out_t.cpu().numpy().view(np.uint64), exp.view(np.uint64))
        if len(packed_docs) > 4:
            part = torch.full((3, len(kps)), -1.0, dtype=torch.float64, device="cuda")
            dev.score_range_dev(kp_t.data_ptr(), off, 2, 3, part.data_ptr(), normalized)
            assert np.array_equal(part.cpu().numpy().view(np.uint64), exp[2:5].view(np.uint64))
        dev.close()
        return idx, out

    try:
        capi.set_option("pipeline_chunk", 40000)   # 3 runs of 148, 148, 104 documents
        # the per-document kernel scores its document itself (default); the batched scorer after each run
        # (option, or a scratch budget too small for a run) must give the same table
        capi.set_option("no_fused_score", 1)
        both(packed, ms, True)
        capi.set_option("no_fused_score", 0)
        capi.set_option("score_tmp_doubles", 20000)
        both(packed, ms, True)
        capi.set_option("score_tmp_doubles", 0)
        for normalized in (True, False):
            idx, out = both(packed, ms, normalized)
            if sa_path != "global_sort":
                assert idx.stat("pipelined") == 1
            for d in (0, 148, 399):
                exp = oracle_mod.OracleEASA(text=packed[d], m=ms[d]).score_many(codes, off, normalized)
                assert np.array_equal(out[d].view(np.uint64), exp.view(np.uint64))
                _check_arrays(idx, d, oracle_mod.OracleEASA(text=packed[d], m=ms[d]), ("fused", d))
        # runs scored speculatively and then discarded: a new symbol, a broken layout, a bucket too large (late runs)
        late = au.pack_strings_collection(["0123456789 QUIZ", "ZEBRA9"])
        idx, _ = both(list(packed) + [late], list(ms) + [2], True)
        if sa_path != "global_sort":
            assert idx.stat("pipeline_miss") == 1 and idx.stat("pipelined") == 0
        weird = au.pack_strings_collection(["中文", "AB"])
        idx, _ = both(list(packed) + [weird], list(ms) + [2], True)
        assert idx.stat("pipelined") == 0
        runs = au.pack_strings_collection(["E" * 6000, "THE END"])
        idx, _ = both(list(packed) + [runs], list(ms) + [2], True)
        assert idx.stat("pipelined") == 0
        # a batch too small to be pipelined: plain build + score inside the call
        both(packed[:5], ms[:5], True)
    finally:
        capi.set_option("pipeline_chunk", 0)
        capi.set_option("no_pipeline", 0)
        capi.set_option("no_fused_score", 0)
        capi.set_option("score_tmp_doubles", 0)
    with pytest.raises(ZeroDivisionError):
        c2, o2 = capi.pack_keyphrases(["A", ""])
        capi.DeviceIndex.build_host_and_score(np.ascontiguousarray(packed[0], dtype=np.uint32), np.array([0, len(packed[0])]),
                                              [ms[0]], c2, o2, np.zeros((1, 2)))


def test_duplicate_query_suffixes_are_scored_once_and_identically(oracle_mod):
    # identical query suffixes (same tail of the same word, repeated keyphrases) are walked once; the table
    # must not depend on that (option score_no_dedup walks every suffix)
    import synth
    capi = _capi()
    packed, ms, _ = synth.packed_collection(3, 4000, first_seed=90)
    idx = capi.DeviceIndex(packed, ms)
    words = ["ALPHA", "BETA", "ALPHABETA", "TA", "A", "GAMMA"]
    rng = np.random.default_rng(2)
    queries = [" ".join(rng.choice(words, size=int(rng.integers(1, 4)))) for _ in range(60)] + ["ALPHA", "ALPHA", "A"]
    from east import utils
    queries += [utils.prepare_text(k) for k in synth.keyphrases(30)] * 2
    queries += ["QQQQQ7", "中文A", "A中", "ALPHA" * 8, "0", "ALPH", "ALPHAB"]   # absent symbols, code points >= 0x0A00, long
    codes, off = capi.pack_keyphrases(queries)
    for normalized in (True, False):
        table = idx.score_table(codes, off, normalized)
        try:
            capi.set_option("score_no_dedup", 1)
            plain = idx.score_table(codes, off, normalized)
        finally:
            capi.set_option("score_no_dedup", 0)
        assert np.array_equal(table.view(np.uint64), plain.view(np.uint64))
        for d in range(3):
            exp = oracle_mod.OracleEASA(text=packed[d], m=ms[d]).score_many(codes, off, normalized)
            assert np.array_equal(table[d].view(np.uint64), exp.view(np.uint64))


def test_terminator_layout_is_validated_by_the_document_kernel(oracle_mod, sa_path):
    # Batches of small documents start with an alphabet-only scan; the per-document kernel itself checks that
    # string k ends with 0x0A00 + k.  A text whose terminator COUNT is right but whose layout is not (swapped,
    # repeated, missing at the end) must be caught there and take the general path: code point order, no
    # terminator semantics -- exactly what the reference's DC3 does with such a string.
    capi = _capi()
    T = 0x0A00
    good = np.array([65, 66, T, 67, T + 1], dtype=np.uint32)
    for bad in ([65, T + 1, 66, T],            # swapped
                [65, T, 66, T],                # repeated
                [65, T, T + 1, 66],            # does not end with its last terminator
                [T + 1, 65, 66, T]):           # starts with one
        bad = np.array(bad, dtype=np.uint32)
        idx = capi.DeviceIndex([good, bad, good], [2, 2, 2])
        info = idx.info()
        assert not info["fast_path"] and not info["doc_sorted"]
        for d, text in enumerate((good, bad, good)):
            o = oracle_mod.OracleEASA(text=text, m=2)
            assert np.array_equal(idx.array(d, capi.SUFTAB), o.suftab), (bad.tolist(), d)
            assert np.array_equal(idx.array(d, capi.LCPTAB), o.lcptab), (bad.tolist(), d)
        idx.close()


def test_document_sizes_around_the_kernel_limits(oracle_mod, sa_path):
    # per-document kernel: 16-bit positions (n <= 65535), the fused LCP / child / annotation phases need the
    # 16-bit LCP copy + pyramid to fit the scratch (n up to ~64.9 k), chunks of ceil(n / 1024) <= 64 ranks per
    # thread in the stack walk; 65 536 code points and more take the global sort
    import synth
    capi = _capi()
    rng = np.random.default_rng(17)
    for target in (1023, 1025, 32768, 49000, 58000, 64800, 65535, 65536, 70000):
        words = ["".join(rng.choice(list("ABCDEFGHIJKLMNOPQRSTUVWXYZ"), size=int(rng.integers(3, 9)))) for _ in range(400)]
        strings, total = [], 0
        while total < target:
            s_ = "".join(rng.choice(words, size=3))
            if total + len(s_) + 1 > target:
                s_ = s_[: max(1, target - total - 1)]
            strings.append(s_)
            total += len(s_) + 1
        if total != target:      # the last string was cut to 1 symbol but still overshoots by one: drop a symbol elsewhere
            strings[0] = strings[0][: len(strings[0]) - (total - target)] or "A"
        from east.asts import utils as au
        packed = au.pack_strings_collection(strings)
        idx = capi.DeviceIndex([packed], [len(strings)])
        info = idx.info()
        if sa_path.startswith("doc_sort"):
            assert info["doc_sorted"] == (packed.size <= 65535), (target, packed.size)
        _check_arrays(idx, 0, oracle_mod.OracleEASA(text=packed, m=len(strings)), target)
        idx.close()
    # thousands of tiny documents in one batch
    tiny = [au.pack_strings_collection(["".join(rng.choice(list("AB "), size=int(rng.integers(1, 6)))) for _ in range(int(rng.integers(1, 4)))])
            for _ in range(3000)]
    tiny_m = [int((p >= 0x0A00).sum()) for p in tiny]
    idx = capi.DeviceIndex(tiny, tiny_m)
    for d in (0, 1, 1499, 2999):
        _check_arrays(idx, d, oracle_mod.OracleEASA(text=tiny[d], m=tiny_m[d]), ("tiny", d))


def test_collections_mixing_small_and_large_documents(oracle_mod, sa_path):
    # east.relevance splits a collection into device batches: documents the per-document kernel takes, and larger ones
    from east import applications, relevance, utils
    import synth
    docs = synth.documents(5, 3000, first_seed=300)
    docs.insert(2, synth.documents(1, 90000, first_seed=400)[0])   # ~82 k code points: global sort
    texts = {"t%d" % i: d for i, d in enumerate(docs)}
    kps = synth.keyphrases(12)
    measure = relevance.ASTRelevanceMeasure("easa", True)
    table = applications.keyphrases_table(kps, texts, measure)
    assert len(measure._batches) == 2
    assert measure._batches[0][0].info()["doc_sorted"] == sa_path.startswith("doc_sort")
    assert not measure._batches[1][0].info()["doc_sorted"]
    names = list(texts.keys())
    for i, name in enumerate(names):
        o = oracle_mod.OracleEASA(utils.text_to_strings_collection(texts[name]))
        for kp in kps:
            q = utils.prepare_text(kp).replace(" ", "")
            from east.asts.utils import codepoints
            codes = np.ascontiguousarray(codepoints(q), dtype=np.uint32)
            exp = o.score_many(codes, np.array([0, len(codes)], dtype=np.int64), True)[0]
            assert float(table[kp][name]).hex() == float(exp).hex(), (name, kp)
