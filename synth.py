"""Synthetic Zipf word text of SURVEY.md section 8(d): the workload of bench.py and of the
size-independent parity tests.

  vocabulary : V = 50 000 words, word i = lowercase a-z string, length uniform in [3, 10],
               numpy.random.default_rng(0)
  documents  : words drawn with p(rank) ~ 1/rank, default_rng(doc seed), joined by single
               spaces, truncated at the target byte size; doc j of a collection has seed j + 1
  keyphrases : 1-3 words (uniform) from the same Zipf, default_rng(7), joined by spaces
"""
import numpy as np

V = 50000
_vocab_cache = {}


def vocabulary(v=V):
    if v not in _vocab_cache:
        rng = np.random.default_rng(0)
        lengths = rng.integers(3, 11, size=v)
        letters = rng.integers(0, 26, size=int(lengths.sum()))
        chars = (letters + ord("a")).astype(np.uint8).tobytes().decode("ascii")
        offs = np.concatenate([[0], np.cumsum(lengths)])
        words = [chars[offs[i]:offs[i + 1]] for i in range(v)]
        ranks = np.arange(1, v + 1, dtype=np.float64)
        cdf = np.cumsum(1.0 / ranks)
        cdf /= cdf[-1]
        _vocab_cache[v] = (words, cdf, lengths)
    return _vocab_cache[v]


def document(n_bytes, seed, v=V):
    """One synthetic .txt document of exactly n_bytes ASCII bytes (last word may be cut)."""
    words, cdf, lengths = vocabulary(v)
    rng = np.random.default_rng(seed)
    # mean word length 6.5 + 1 space; oversample, then truncate
    count = int(n_bytes / 5.0) + 16
    ids = np.searchsorted(cdf, rng.random(count), side="right")
    text = " ".join(words[i] for i in ids)
    while len(text) < n_bytes:
        ids = np.searchsorted(cdf, rng.random(count), side="right")
        text += " " + " ".join(words[i] for i in ids)
    return text[:n_bytes]


def documents(n_docs, n_bytes, first_seed=1, v=V):
    return [document(n_bytes, first_seed + j, v) for j in range(n_docs)]


def keyphrases(k, seed=7, v=V):
    words, cdf, _ = vocabulary(v)
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(k):
        nw = int(rng.integers(1, 4))
        ids = np.searchsorted(cdf, rng.random(nw), side="right")
        out.append(" ".join(words[i] for i in ids))
    return out


def packed_collection(n_docs, n_bytes, first_seed=1):
    """Documents run through the host preprocessing of the product package and packed:
    returns (list of uint32 arrays, list of m, list of strings collections)."""
    from east import utils
    from east.asts import utils as asts_utils
    cols = [utils.text_to_strings_collection(d) for d in documents(n_docs, n_bytes, first_seed)]
    packed = [asts_utils.pack_strings_collection(c) for c in cols]
    return packed, [len(c) for c in cols], cols
