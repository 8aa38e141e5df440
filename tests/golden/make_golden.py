#!/usr/bin/env python3
"""Generate the committed golden vectors from the REFERENCE ITSELF.

Runs the py3-patched copy of the reference (oracle/_ref, produced by oracle/make_ref.py from
/root/reference; mechanical patch only) and records inputs and outputs of the EASA hot path:

  golden.json        strings collections, queries, scores (float.hex, exact), exceptions,
                     keyphrase table / graph of the HSE sample corpus
  golden_arrays.npz  suftab / lcptab / childtab_* / anntab of every case that lists "arrays"

Nothing here runs on the GPU box; tests read only the two output files.

Usage:  python oracle/make_ref.py && python tests/golden/make_golden.py
"""
import json
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
warnings.filterwarnings("ignore")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from east import applications, relevance, utils  # noqa: E402  (the reference)
from east.asts import base  # noqa: E402
import synth  # noqa: E402

ARRAYS = ["suftab", "lcptab", "childtab_up", "childtab_down", "childtab_next_l_index", "anntab"]


def fhex(x):
    return float(x).hex()


def case(name, strings, queries, arrays_out, with_arrays=True, algorithms=("easa",)):
    ast = base.AST.get_ast(strings, "easa")
    rec = {"name": name, "strings": strings, "queries": []}
    if with_arrays:
        rec["arrays"] = True
        for a in ARRAYS:
            arrays_out["%s/%s" % (name, a)] = np.asarray(getattr(ast, a), dtype=np.int32)
    others = [base.AST.get_ast(strings, alg) for alg in algorithms if alg != "easa"]
    for q in queries:
        item = {"q": q}
        for normalized in (True, False):
            try:
                score, suffix_scores = ast.score(q, normalized=normalized, return_suffix_scores=True)
                for o in others:  # the reference's own test: engines agree exactly (tests/asts/test_base.py)
                    assert o.score(q, normalized=normalized) == score, (name, q)
                item["norm" if normalized else "denorm"] = fhex(score)
                item["suffix_norm" if normalized else "suffix_denorm"] = {k: fhex(v) for k, v in suffix_scores.items()}
            except ZeroDivisionError:
                item["raises"] = "ZeroDivisionError"
        rec["queries"].append(item)
    return rec


def main():
    random.seed(20240229)
    arrays = {}
    cases = []
    all_algs = ("easa", "ast_linear", "ast_naive")
    # README.rst:147-152 known answer
    cases.append(case("readme", ["XABXAC", "HI"], ["ABCI", "NOPE", "XABXAC", "A B C I", "", " "], arrays,
                      algorithms=all_algs))
    # tests/asts/test_base.py:13-14 fixture
    cases.append(case("test_base", ["abcd efg ops", "xyzq", "test"], ["aqcb", "efgp", "mn4", "abcd efg ops"],
                      arrays, algorithms=all_algs))
    # doc/samples/texts/test.txt x doc/samples/keyphrases/test.txt
    cases.append(case("sample_test", utils.text_to_strings_collection("XABXAC"), ["ABC", "ORC", "NONE"], arrays,
                      algorithms=all_algs))
    # degenerate collection produced for an empty text (utils.py:76-78)
    cases.append(case("blank", [" "], ["AB", "A"], arrays))
    cases.append(case("single_char", ["A"], ["A", "AA", "B"], arrays, algorithms=all_algs))
    cases.append(case("repeats", ["AAAAAAAA", "AAAA", "AAAAAAAA"], ["A", "AAAA", "AAAAAAAAAAAA", "BA"], arrays,
                      algorithms=all_algs))
    # analysis/utils.py:5-9 "worst case": m identical strings (deep LCP)
    deep = ["".join(random.choice("AB") for _ in range(96))] * 6
    cases.append(case("deep_lcp", deep, [deep[0][:20], deep[0][40:70], "ABAB", "BBBBBB"], arrays))
    # random small alphabets
    for sigma in (2, 4, 7, 27):
        alpha = "ABCDEFGHIJKLMNOPQRSTUVWXYZ "[:sigma]
        for t in range(3):
            m = random.randint(1, 9)
            strings = ["".join(random.choice(alpha) for _ in range(random.randint(1, 14))) for _ in range(m)]
            letters = alpha.strip() or "A"
            queries = ["".join(random.choice(letters) for _ in range(random.randint(1, 9))) for _ in range(6)]
            cases.append(case("rand_s%d_%d" % (sigma, t), strings, queries, arrays,
                              algorithms=all_algs if t == 0 else ("easa",)))
    # Cyrillic + Latin + digits + apostrophes
    cases.append(case("cyrillic", utils.text_to_strings_collection(
        "Съешь же ещё этих мягких французских булок, да выпей чаю. Don't panic: 42 isn't the answer'"),
        ["БУЛОК", "ЧАЮ ДА", "DON'T", "ФРАНЦУЗ", "XYZ"], arrays))

    # synthetic Zipf document of the benchmark generator (10 KB, seed 1) with 12 keyphrases
    doc = synth.document(10000, 1)
    col = utils.text_to_strings_collection(doc)
    kps = [utils.prepare_text(k) for k in synth.keyphrases(12)]
    cases.append(case("zipf10k", col, kps, arrays))

    # HSE sample corpus: doc/samples/keyphrases/HSE.txt x doc/samples/texts/HSE rules/*.txt
    hse = {"docs": [], "keyphrases": []}
    sdir = os.path.join(REF, "samples")
    with open(os.path.join(sdir, "keyphrases", "HSE.txt"), "rb") as f:
        keyphrases = [line.decode("utf-8") for line in f.read().splitlines()]
    hse["keyphrases"] = keyphrases
    tdir = os.path.join(sdir, "texts", "HSE rules")
    texts = {}
    for fn in sorted(os.listdir(tdir)):
        if fn.endswith(".txt"):
            with open(os.path.join(tdir, fn), "rb") as f:
                texts[fn] = f.read().decode("utf-8")
    names = list(texts.keys())
    for j, fn in enumerate(names):
        col = utils.text_to_strings_collection(texts[fn])
        rec = {"name": fn, "strings": col}
        if j < 4:
            ast = base.AST.get_ast(col, "easa")
            rec["arrays"] = True
            for a in ARRAYS:
                arrays["hse%d/%s" % (j, a)] = np.asarray(getattr(ast, a), dtype=np.int32)
        hse["docs"].append(rec)
    for normalized in (True, False):
        table = applications.keyphrases_table(keyphrases, texts, relevance.ASTRelevanceMeasure("easa", normalized))
        hse["table_norm" if normalized else "table_denorm"] = {
            kp: {fn: fhex(v) for fn, v in row.items()} for kp, row in table.items()}
    graphs = []
    for c, r, p in ((0.6, 0.25, 1), (0.3, 0.2, 3), (0.9, 0.3, 1)):
        g = applications.keyphrases_graph(keyphrases, texts, referral_confidence=c, relevance_threshold=r,
                                          support_threshold=p,
                                          similarity_measure=relevance.ASTRelevanceMeasure("easa", True))
        graphs.append({"c": c, "r": r, "p": p, "nodes": g["nodes"],
                       "edges": [{"source": e["source"], "target": e["target"], "confidence": fhex(e["confidence"])}
                                 for e in g["edges"]]})
    hse["graphs"] = graphs

    # traversal callbacks (easa.py:38-85): pre-order nodes (l, i, j, char) and post-order nodes
    # (l, i, j, [children as (l, i, j)])
    traversals = []
    for strings in (["XABXAC", "HI"], ["abcd efg ops", "xyzq", "test"], ["AAAAAAAA", "AAAA", "AAAAAAAA"],
                    utils.text_to_strings_collection(doc[:600])):
        ast = base.AST.get_ast(strings, "easa")
        pre, post = [], []
        ast.traverse(lambda node: pre.append([int(node[0]), int(node[1]), int(node[2]), node[3]]),
                     "depth-first|pre-order")
        ast.traverse(lambda node: post.append([int(node[0]), int(node[1]), int(node[2]),
                                               [[int(c[0]), int(c[1]), int(c[2])] for c in node[3]]]),
                     "depth-first|post-order")
        traversals.append({"strings": strings, "pre": pre, "post": post})

    # host preprocessing known answers (tests/test_utils.py:10-13 and utils.py:49-79 behaviour)
    prep = {
        "tokenize": {"in": "Well, what a sunny day!", "out": utils.tokenize("Well, what a sunny day!")},
        "collections": [{"in": t, "out": utils.text_to_strings_collection(t)} for t in [
            "Well, what a sunny day!", "", "a bb 12 345", "one two three four five six seven",
            "It's 2024: don't stop_me now... ёлка Ёж", doc[:300]]],
    }
    errors = {}
    for label, fn in (("empty_collection", lambda: base.AST.get_ast([])),
                      ("unknown_algorithm", lambda: base.AST.get_ast(["A"], "nope"))):
        try:
            fn()
            errors[label] = None
        except Exception as e:  # noqa: BLE001
            errors[label] = type(e).__name__

    with open(os.path.join(HERE, "golden.json"), "w", encoding="utf-8") as f:
        json.dump({"cases": cases, "hse": hse, "prep": prep, "errors": errors, "traversals": traversals,
                   "source": "py3-patched reference (oracle/make_ref.py), EAST 0.3.8"}, f, ensure_ascii=False, separators=(",", ":"))
    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **arrays)
    print("cases:", len(cases), "arrays:", len(arrays))


if __name__ == "__main__":
    main()
