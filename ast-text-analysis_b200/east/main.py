# -*- coding: utf-8 -*-
"""Thin command line front end: `east [options] keyphrases table|graph <keyphrases.txt> <dir|file>`.

Mirror of east/main.py:15-152 with the same getopt string ("s:a:w:v:l:f:c:r:p:dy") and defaults,
minus its three defects at the surveyed commit: the `-s` value is compared
case-insensitively (the default "AST" never matched the test `== "ast"`), the similarity measure
is actually passed to keyphrases_table (the reference used an undefined name), and format_table
receives its argument.  Only the AST measure is available (the cosine measure and the Tomita
synonym extractor are outside the accelerated path); asking for them is an error, not a fallback.

    python -m east.main -a easa -f csv keyphrases table kp.txt texts/
"""
import getopt
import os
import sys

from east import applications
from east import consts
from east import formatting
from east import relevance

USAGE = ("Invalid syntax: EAST should be called as:\n\n"
         "    east [options] <command> <subcommand> args\n\n"
         "Commands available: keyphrases.\n"
         "Subcommands available: table/graph.")


def read_keyphrases(path):
    with open(path, "rb") as f:
        return f.read().decode("utf-8", errors="replace").splitlines()


def read_texts(path):
    """A directory of *.txt files (one text per file, named by the file without '.txt') or a single
    file (one text per line, named "0", "1", ...) -- east/main.py:66-89."""
    path = os.path.abspath(path)
    if os.path.isdir(path):
        files = [os.path.join(path, name) for name in os.listdir(path) if name.endswith(".txt")]
    else:
        files = [path]
    texts = {}
    if len(files) == 1:
        with open(files[0], "rb") as f:
            for i, line in enumerate(f.read().decode("utf-8", errors="replace").splitlines()):
                texts[str(i)] = line
    else:
        for filename in files:
            with open(filename, "rb") as f:
                texts[os.path.basename(filename)[:-4]] = f.read().decode("utf-8", errors="replace")
    return texts


def main(argv=None, out=None):
    out = out or sys.stdout
    argv = sys.argv[1:] if argv is None else argv
    try:
        opts, args = getopt.getopt(argv, "s:a:w:v:l:f:c:r:p:dy")
    except getopt.GetoptError as e:
        out.write("%s\n" % e)
        return 1
    opts = dict(opts)
    opts.setdefault("-l", consts.Language.ENGLISH)
    opts.setdefault("-s", consts.RelevanceMeasure.AST)
    opts.setdefault("-a", consts.ASTAlgorithm.EASA)
    opts.setdefault("-c", "0.6")
    opts.setdefault("-r", "0.25")
    opts.setdefault("-p", "1")

    if len(args) < 2:
        out.write(USAGE + "\n")
        return 1
    command, subcommand = args[0], args[1]
    if command != "keyphrases":
        out.write("Invalid command: '%s'. Please use one of: 'keyphrases'.\n" % command)
        return 1
    if len(args) < 4:
        out.write('Invalid syntax. For keyphrases analysis, EAST should be called as:\n\n'
                  '    east [options] keyphrases <subcommand> "path/to/keyphrases.txt" "path/to/texts/dir"\n')
        return 1
    if subcommand not in ("table", "graph"):
        out.write("Invalid subcommand: '%s'. Please use one of: 'table', 'graph'.\n" % subcommand)
        return 1
    if opts["-s"].lower() != "ast":
        out.write("Only the AST relevance measure (-s ast) is available in this build.\n")
        return 1
    if "-y" in opts:
        out.write("Synonym extraction (-y) needs the external Tomita parser and is not available.\n")
        return 1

    keyphrases = read_keyphrases(os.path.abspath(args[2]))
    texts = read_texts(args[3])
    measure = relevance.ASTRelevanceMeasure(opts["-a"], normalized="-d" not in opts)

    try:
        if subcommand == "table":
            table = applications.keyphrases_table(keyphrases, texts, measure, None, opts["-l"])
            res = formatting.format_table(table, opts.get("-f", "xml").lower())
        else:
            graph = applications.keyphrases_graph([k for k in keyphrases if k], texts, float(opts["-c"]),
                                                  float(opts["-r"]), float(opts["-p"]), measure, None, opts["-l"])
            res = formatting.format_graph(graph, opts.get("-f", "edges").lower())
    except Exception as e:  # noqa: BLE001  (the reference prints the error and returns 1)
        out.write("%s\n" % e)
        return 1
    out.write(res if res.endswith("\n") else res + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
