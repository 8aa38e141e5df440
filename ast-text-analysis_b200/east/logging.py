# -*- coding: utf-8 -*-
"""Progress reporting: `progress(message, step, total)` / `clear()` -- the two calls the reference's callers
make (east/logging.py:8-17).  One carriage-return line on a terminal, silence when the output is a pipe or file."""
import sys

_WIDTH = 80
_shown = [False]


def _interactive():
    out = sys.stdout
    return bool(getattr(out, "isatty", None)) and out.isatty()


def _emit(line):
    sys.stdout.write("\r" + line)
    sys.stdout.flush()


def progress(message, step, total):
    if _interactive():
        _shown[0] = True
        _emit("{}: {:d}/{:d}".format(message, int(step), int(total)))


def clear():
    if _interactive() and _shown[0]:
        _emit(" " * _WIDTH + "\r")
        _shown[0] = False
