#!/usr/bin/env python3
"""Materialise a runnable copy of the UNMODIFIED-ALGORITHM reference under oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path may import from oracle/.

The reference (EAST 0.3.8, /root/reference) is pure Python 2 and hard-imports nltk
and testtools, none of which exist in this image.  This script copies the `east`
package and the reference's own `tests/` into oracle/_ref/ (git-ignored, never
committed) and applies a purely mechanical py2->py3 patch: xrange/unichr/np.int
renames, `/` -> `//` on the integer divisions of the DC3 code, metaclass syntax and
a few iterator/list fixes.  No algorithmic line is touched.  Two shim packages
(nltk, testtools) satisfy import-time dependencies that the EASA path never calls.

The resulting package is used
  * by tests/golden/make_golden.py to generate the committed golden vectors,
  * by tests (when oracle/_ref exists) to validate the C restatement in oracle/,
  * by bench.py --impl reference / cpu_baseline as the timed CPU reference.

Usage: python oracle/make_ref.py [--src /root/reference] [--dst oracle/_ref]
"""
import argparse
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _sub(text, pattern, repl, count=0, must=True, flags=0):
    new, n = re.subn(pattern, repl, text, count=count, flags=flags)
    if must and n == 0:
        raise RuntimeError("patch pattern did not match: %r" % pattern)
    return new


def patch_common(src):
    src = re.sub(r"\bxrange\b", "range", src)
    src = re.sub(r"\bunichr\b", "chr", src)
    src = re.sub(r"\bnp\.int\b(?!\d)", "np.int64", src)
    return src


def patch_easa(src):
    # integer divisions of the skew (DC3) code: py2 `/` on ints is floor division
    src = _sub(src, r"n0 = \(n \+ 2\) / 3", "n0 = (n + 2) // 3")
    src = _sub(src, r"n1 = \(n \+ 1\) / 3", "n1 = (n + 1) // 3")
    src = _sub(src, r"n2 = n / 3", "n2 = n // 3")
    src = _sub(src, r"SA12\[i\] / 3", "SA12[i] // 3")
    src = _sub(src, r"j/3", "j//3")
    src = _sub(src, r"suffix_result /= matched_chars", "suffix_result /= matched_chars")
    return src


def patch_asts_utils(src):
    # the "\U%08x".decode("unicode-escape") trick == chr(0x0A00 + i)
    src = _sub(
        src,
        r"hex_code = hex\(.*?\n\s*hex_code = .*?\n\s*res\.append\(strings_collection\[i\] \+ hex_code\.decode\(\"unicode-escape\"\)\)",
        "res.append(strings_collection[i] + chr(consts.String.UNICODE_SPECIAL_SYMBOLS_START + i))",
        flags=re.S,
    )
    return src


def patch_metaclass(src, clsname, base):
    src = _sub(src, r"class %s\(%s\):\n\s*__metaclass__ = abc\.ABCMeta" % (clsname, re.escape(base)),
               "class %s(%s, metaclass=abc.ABCMeta):" % (clsname, base if base != "object" else "object"))
    return src


def patch_ast_linear(src):
    src = _sub(src, r"for (\w+) in root\.children\.keys\(\):", r"for \1 in list(root.children.keys()):", must=False)
    src = _sub(src, r"for (\w+) in root\.children:", r"for \1 in list(root.children):", must=False)
    return src


def patch_exceptions(src):
    src = _sub(src, r"raise exc_info\[0\], exc_info\[1\], exc_info\[2\]", "raise exc_info[1]")
    src = src.replace("return unicode(self)", "return str(self)")
    return src


def patch_utils(src):
    src = src.replace("itertools.imap", "map")
    src = _sub(src, r"text = unicode\(text\.decode\('utf-8', errors='replace'\)\)",
               "text = text.decode('utf-8', errors='replace') if isinstance(text, bytes) else text")
    src = _sub(src, r"strings_collection = filter\((.*?), strings_collection\)",
               r"strings_collection = list(filter(\1, strings_collection))")
    return src


def patch_applications(src):
    src = _sub(src, r"text_titles = texts\.keys\(\)", "text_titles = list(texts.keys())")
    src = _sub(src, r"text_collection = texts\.values\(\)", "text_collection = list(texts.values())")
    return src


NLTK_INIT = '"""Import-time shim: the EASA path never calls nltk."""\n'
NLTK_CORPUS = (
    "class _Stopwords(object):\n"
    "    def words(self, language):\n"
    "        return []\n\n"
    "stopwords = _Stopwords()\n"
)
NLTK_STEM = (
    "class _Snowball(object):\n"
    "    class SnowballStemmer(object):\n"
    "        def __init__(self, language):\n"
    "            self.language = language\n"
    "        def stem(self, token):\n"
    "            return token\n\n"
    "snowball = _Snowball()\n"
)
TESTTOOLS = "import unittest\n\nTestCase = unittest.TestCase\n"


def build(src_root, dst_root):
    if not os.path.isdir(os.path.join(src_root, "east")):
        raise SystemExit("reference not found at %s" % src_root)
    if os.path.isdir(dst_root):
        shutil.rmtree(dst_root)
    os.makedirs(dst_root)
    # the synonyms sub-package (external Tomita binary, py2-only lambda syntax) is off the path
    shutil.copytree(os.path.join(src_root, "east"), os.path.join(dst_root, "east"),
                    ignore=shutil.ignore_patterns("synonyms", "main.py", "*.pyc", "__pycache__"))
    shutil.copytree(os.path.join(src_root, "tests"), os.path.join(dst_root, "tests"),
                    ignore=shutil.ignore_patterns("*.pyc", "__pycache__"))
    samples = os.path.join(src_root, "doc", "samples")
    if os.path.isdir(samples):
        shutil.copytree(samples, os.path.join(dst_root, "samples"))

    for dirpath, _, files in os.walk(dst_root):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            path = os.path.join(dirpath, fn)
            with open(path, encoding="utf-8") as f:
                src = f.read()
            rel = os.path.relpath(path, dst_root).replace(os.sep, "/")
            src = patch_common(src)
            if rel == "east/asts/easa.py":
                src = patch_easa(src)
            elif rel == "east/asts/utils.py":
                src = patch_asts_utils(src)
            elif rel == "east/asts/base.py":
                src = patch_metaclass(src, "AST", "object")
            elif rel == "east/asts/ast.py":
                src = patch_metaclass(src, "AnnotatedSuffixTree", "base.AST")
            elif rel == "east/asts/ast_linear.py":
                src = patch_ast_linear(src)
            elif rel == "east/exceptions.py":
                src = patch_exceptions(src)
            elif rel == "east/utils.py":
                src = patch_utils(src)
            elif rel == "east/applications.py":
                src = patch_applications(src)
            with open(path, "w", encoding="utf-8") as f:
                f.write(src)

    shims = {
        "nltk/__init__.py": NLTK_INIT,
        "nltk/corpus/__init__.py": NLTK_CORPUS,
        "nltk/stem/__init__.py": NLTK_STEM,
        "testtools/__init__.py": TESTTOOLS,
    }
    for rel, body in shims.items():
        path = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(body)
    return dst_root


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=os.path.join(HERE, "_ref"))
    args = ap.parse_args()
    dst = build(args.src, args.dst)
    print("reference materialised at", dst)


if __name__ == "__main__":
    sys.exit(main())
