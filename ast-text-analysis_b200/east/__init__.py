# -*- coding: utf-8 -*-
"""east -- host-side mirror of EAST's public API with the "easa" engine on B200.

Same module layout as the reference (east/__init__.py:1-3): importing the package imports
every engine module under east.asts so that AST.get_ast() can find it in the subclass
registry (east/asts/base.py:13-18).
"""
from east import utils

utils.import_modules_from_package("east.asts")
