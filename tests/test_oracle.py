"""CPU: the C restatement in oracle/ against the golden vectors captured from the reference."""
import numpy as np
import pytest

from conftest import ARRAY_NAMES


def _cases(golden):
    for c in golden["cases"]:
        yield c["name"], c
    for j, d in enumerate(golden["hse"]["docs"]):
        yield "hse%d" % j, d


def test_oracle_arrays_match_reference(golden, oracle_mod):
    checked = 0
    for key, c in _cases(golden):
        if not c.get("arrays"):
            continue
        o = oracle_mod.OracleEASA(c["strings"])
        for a in ARRAY_NAMES:
            ref = golden["arrays"]["%s/%s" % (key, a)]
            assert np.array_equal(getattr(o, a), ref), (key, a)
            checked += 1
    assert checked >= 6 * 20


def test_oracle_scores_bit_exact(golden, oracle_mod):
    checked = 0
    for c in golden["cases"]:
        o = oracle_mod.OracleEASA(c["strings"])
        for q in c["queries"]:
            if q.get("raises"):
                with pytest.raises(ZeroDivisionError):
                    o.score(q["q"])
                continue
            for normalized, key, skey in ((True, "norm", "suffix_norm"), (False, "denorm", "suffix_denorm")):
                score, per_suffix = o.score(q["q"], normalized, return_suffix_scores=True)
                assert float(score).hex() == q[key], (c["name"], q["q"], normalized)
                assert {k: float(v).hex() for k, v in per_suffix.items()} == q[skey]
                # SA-interval narrowing is the same function (formulation of the CUDA scorer)
                qc = oracle_mod.codepoints(q["q"].replace(" ", ""))
                if len(o.text) <= 400:
                    alt = oracle_mod.interval_score(o.text, o.m, o.suftab, qc, normalized)
                    assert float(alt).hex() == q[key]
                checked += 1
    assert checked > 150


def test_oracle_known_answers(oracle_mod):
    # README.rst:149-152
    o = oracle_mod.OracleEASA(["XABXAC", "HI"])
    assert o.score("ABCI") == 0.1875
    assert o.score("NOPE") == 0
    # (0.05 + 1) - 1 is NOT 0.05
    o = oracle_mod.OracleEASA(["abcd efg ops", "xyzq", "test"])
    assert float(o.score("aqcb")).hex() == "0x1.99999999999a0p-5"
    assert float(o.score("efgp", normalized=False)) == 0.6875


def test_oracle_hse_table(golden, oracle_mod):
    hse = golden["hse"]
    asts = {d["name"]: oracle_mod.OracleEASA(d["strings"]) for d in hse["docs"]}
    for normalized, key in ((True, "table_norm"), (False, "table_denorm")):
        for kp, row in hse[key].items():
            for fn, hexv in row.items():
                assert float(asts[fn].score(kp.upper(), normalized)).hex() == hexv, (kp, fn)


def test_oracle_root_annotation_and_leaf_counts(oracle_mod):
    rng = np.random.default_rng(5)
    for _ in range(50):
        m = int(rng.integers(1, 8))
        strings = ["".join(rng.choice(list("ABC"), size=int(rng.integers(1, 10)))) for _ in range(m)]
        o = oracle_mod.OracleEASA(strings)
        assert o.anntab[0] == o.n - o.m
        assert sorted(o.suftab.tolist()) == list(range(o.n))


def test_bench_row_checker_subprocess(oracle_mod):
    # bench.py checks rows of its table in a subprocess (oracle/check_rows.py): right rows pass, a flipped bit is caught
    import importlib.util
    import os
    import synth
    from conftest import ROOT
    from east import _capi, utils
    from east.asts import utils as au
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cols = [utils.text_to_strings_collection(d) for d in synth.documents(3, 2000, first_seed=11)]
    packed = [au.pack_strings_collection(c) for c in cols]
    codes, off = _capi.pack_keyphrases([utils.prepare_text(k) for k in synth.keyphrases(15)])
    rows = [oracle_mod.OracleEASA(text=p, m=len(c)).score_many(codes, off, True) for p, c in zip(packed, cols)]
    sample = [(p, len(c), r) for p, c, r in zip(packed, cols, rows)]
    assert bench.oracle_check(sample, codes, off, True) == {"rows": 3, "mismatching_rows": 0}
    rows[2] = rows[2].copy()
    rows[2][4] = np.nextafter(rows[2][4], 1.0)
    sample[2] = (packed[2], len(cols[2]), rows[2])
    assert bench.oracle_check(sample, codes, off, True)["mismatching_rows"] == 1
