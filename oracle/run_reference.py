#!/usr/bin/env python3
"""Time the CPU implementation of the hot path on this box's host cores.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and `--impl reference`).

kind = "reference": the reference's own code (py3-patched copy in oracle/_ref, see make_ref.py):
                    east.applications.keyphrases_table(keyphrases, texts,
                    east.relevance.ASTRelevanceMeasure("easa", normalized=True)) -- the functions the
                    (broken) CLI `east keyphrases table -a easa` would reach.
kind = "port":      the C restatement oracle/east_oracle.c, used only when oracle/_ref is absent.

Workload: the synthetic Zipf documents / keyphrases of synth.py, same seeds as bench.py.
procs == 1 runs the table in-process; procs > 1 runs one task per document (build its AST, score
all K keyphrases) on a multiprocessing pool -- "one process per core across documents".
Prints one JSON object.
"""
import argparse
import json
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
warnings.filterwarnings("ignore")
sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != HERE]  # `oracle` must resolve to the package


def have_reference():
    return os.path.isdir(os.path.join(REF, "east"))


def _ref_doc_task(args):
    text, prepared_kps = args
    from east import utils
    from east.asts import base
    t0 = time.perf_counter()
    ast = base.AST.get_ast(utils.text_to_strings_collection(text), "easa")
    t1 = time.perf_counter()
    scores = [ast.score(kp, normalized=True) for kp in prepared_kps]
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, float(sum(scores))


def _port_doc_task(args):
    text, prepared_kps = args
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "ast-text-analysis_b200"))
    from oracle import oracle
    from east import utils  # host preprocessing of the product package (pure Python, no GPU involved)
    t0 = time.perf_counter()
    ast = oracle.OracleEASA(utils.text_to_strings_collection(text))
    t1 = time.perf_counter()
    scores = [ast.score(kp, True) for kp in prepared_kps]
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, float(sum(scores))


def run(kind, n_docs, doc_bytes, n_kps, procs, first_seed=1):
    sys.path.insert(0, ROOT)
    import synth
    docs = synth.documents(n_docs, doc_bytes, first_seed)
    kps = synth.keyphrases(n_kps)
    if kind == "reference":
        sys.path.insert(0, REF)
        from east import applications, relevance, utils
        prepared = [utils.prepare_text(k) for k in kps]
        task = _ref_doc_task
    else:
        prepared = [k.upper() for k in kps]
        task = _port_doc_task
    t0 = time.perf_counter()
    if procs <= 1 and kind == "reference":
        texts = {"doc%05d.txt" % i: d for i, d in enumerate(docs)}
        class TimedMeasure(relevance.ASTRelevanceMeasure):  # timing wrapper only
            index_s = 0.0

            def set_text_collection(self, *a, **kw):
                t = time.perf_counter()
                super(TimedMeasure, self).set_text_collection(*a, **kw)
                TimedMeasure.index_s = time.perf_counter() - t

        measure = TimedMeasure("easa", normalized=True)
        ti = time.perf_counter()
        # keyphrases_table = set_text_collection (index) + the K x D loop (applications.py:35-52)
        table = applications.keyphrases_table(kps, texts, measure)
        total = time.perf_counter() - ti
        index_s = TimedMeasure.index_s
        score_s = max(total - index_s, 1e-9)
        checksum = float(sum(sum(float(v) for v in row.values()) for row in table.values()))
        n_scores = sum(len(r) for r in table.values())
    else:
        work = [(d, prepared) for d in docs]
        if procs <= 1:
            res = [task(w) for w in work]
        else:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(task, work, chunksize=1)
        total = time.perf_counter() - t0
        index_s = sum(r[0] for r in res)   # CPU-seconds, summed over workers
        score_s = sum(r[1] for r in res)
        checksum = float(sum(r[2] for r in res))
        n_scores = n_docs * n_kps
    return {"kind": kind, "docs": n_docs, "doc_bytes": doc_bytes, "keyphrases": n_kps, "procs": procs,
            "wall_s": total, "index_cpu_s": index_s, "score_cpu_s": score_s, "scores": n_scores,
            "scores_per_s": n_scores / total, "build_MB_per_s": n_docs * doc_bytes / 1e6 / max(index_s / max(procs, 1), 1e-9),
            "checksum": checksum, "cores_available": os.cpu_count()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--docs", type=int, default=8)
    ap.add_argument("--doc-bytes", type=int, default=50000)
    ap.add_argument("--kps", type=int, default=200)
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--first-seed", type=int, default=1)
    a = ap.parse_args()
    kind = a.kind
    if kind == "auto":
        kind = "reference" if have_reference() else "port"
    print(json.dumps(run(kind, a.docs, a.doc_bytes, a.kps, a.procs, a.first_seed)))


if __name__ == "__main__":
    main()
