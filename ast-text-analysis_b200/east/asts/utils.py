# -*- coding: utf-8 -*-
"""Helpers of the AST engines (mirror of east/asts/utils.py:6-40) plus the host-side packer."""
import numpy as np

from east import consts


def index(array, key, start=0):
    """Position of the first element equal to key at or after start (east/asts/utils.py:6-11)."""
    i = start
    while array[i] != key:
        i += 1
    return i


def match_strings(str1, str2):
    """Length of the common prefix (east/asts/utils.py:14-22)."""
    limit = min(len(str1), len(str2))
    i = 0
    while i < limit and str1[i] == str2[i]:
        i += 1
    return i


def make_unique_endings(strings_collection):
    """String i gets the terminator chr(0x0A00 + i) appended (east/asts/utils.py:25-40)."""
    base = consts.String.UNICODE_SPECIAL_SYMBOLS_START
    return [s + chr(base + i) for i, s in enumerate(strings_collection)]


def codepoints(s):
    """unicode string -> numpy uint32 code points"""
    if not s:
        return np.zeros(0, dtype=np.uint32)
    return np.frombuffer(s.encode("utf-32-le", errors="surrogatepass"), dtype=np.uint32)


def pack_strings_collection(strings_collection):
    """The packed document the device engine consumes: make_unique_endings + "".join
    (east/asts/utils.py:25-40, east/asts/easa.py:19) as uint32 code points.  Unlike chr(),
    uint32 terminators are not capped at 0x110000 - 0x0A00 strings."""
    m = len(strings_collection)
    lengths = np.fromiter((len(s) for s in strings_collection), dtype=np.int64, count=m)
    chars = codepoints("".join(strings_collection))
    if chars.size != int(lengths.sum()):
        raise ValueError("strings collection contains lone surrogates that do not round-trip")
    n = int(chars.size) + m
    term_pos = np.cumsum(lengths + 1) - 1
    out = np.empty(n, dtype=np.uint32)
    is_char = np.ones(n, dtype=bool)
    is_char[term_pos] = False
    out[is_char] = chars
    out[term_pos] = consts.String.UNICODE_SPECIAL_SYMBOLS_START + np.arange(m, dtype=np.uint32)
    return out


def pack_strings_collection_u8(strings_collection):
    """The same packed document as ONE BYTE per code point, for collections whose code points are all below 0xFF (ASCII,
    Latin-1): every string followed by the byte 0xFF, which stands for its terminator 0x0A00 + i (the engine restores the
    code points on the device).  A quarter of the bytes of pack_strings_collection() over the host link.  Returns None
    when a code point does not fit."""
    try:
        data = ("\xff".join(strings_collection) + "\xff").encode("latin-1")
    except UnicodeEncodeError:
        return None
    if data.count(b"\xff") != len(strings_collection):
        return None   # a string contains U+00FF itself
    return np.frombuffer(data, dtype=np.uint8)
