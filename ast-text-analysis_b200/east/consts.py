# -*- coding: utf-8 -*-
"""Constants of the EAST API (mirror of east/consts.py:1-74: same names, same values)."""


class _Enum(object):
    """Read-only namespace; iterating yields the values (east/utils.py:12-29 mixins)."""

    def __setattr__(self, key, value):
        raise TypeError("%s is immutable" % type(self).__name__)

    def __iter__(self):
        for name in dir(type(self)):
            if not name.startswith("_"):
                yield getattr(self, name)


class _TraversalOrder(_Enum):
    DEPTH_FIRST_PRE_ORDER = "depth-first|pre-order"
    DEPTH_FIRST_POST_ORDER = "depth-first|post-order"
    BREADTH_FIRST = "breadth-first"


class _String(_Enum):
    # first code point used for the per-string unique terminators (consts.py:23-24)
    UNICODE_SPECIAL_SYMBOLS_START = 0x0A00


class _RelevanceMeasure(_Enum):
    AST = "AST"
    COSINE = "cosine"


class _ASTAlgorithm(_Enum):
    AST_LINEAR = "ast_linear"
    AST_NAIVE = "ast_naive"
    EASA = "easa"


class _TermWeighting(_Enum):
    TF = "tf"
    TF_IDF = "tf-idf"


class _VectorSpace(_Enum):
    WORDS = "words"
    STEMS = "stems"
    LEMMATA = "lemmata"


class _Language(_Enum):
    DANISH = "danish"
    DUTCH = "dutch"
    ENGLISH = "english"
    FINNISH = "finnish"
    FRENCH = "french"
    GERMAN = "german"
    HUNGARIAN = "hungarian"
    ITALIAN = "italian"
    NORWEGIAN = "norwegian"
    PORTUGUESE = "portuguese"
    ROMANIAN = "romanian"
    RUSSIAN = "russian"
    SPANISH = "spanish"
    SWEDISH = "swedish"


TraversalOrder = _TraversalOrder()
String = _String()
RelevanceMeasure = _RelevanceMeasure()
ASTAlgorithm = _ASTAlgorithm()
TermWeighting = _TermWeighting()
VectorSpace = _VectorSpace()
Language = _Language()
