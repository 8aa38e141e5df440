# -*- coding: utf-8 -*-
"""Relevance measures (mirror of east/relevance.py:16-53; the cosine measure is out of scope).

ASTRelevanceMeasure keeps the reference's interface -- set_text_collection(texts, language)
and relevance(keyphrase, text=j) -- but indexes the whole collection as ONE device batch and
adds relevance_table(), which scores every (keyphrase, text) pair in one call; that is what
east.applications.keyphrases_table uses instead of K x D single-pair calls.
"""
import numpy as np

from east import _capi
from east import consts
from east import logging
from east import utils
from east.asts import base
from east.asts import easa
from east.asts import utils as asts_utils


class RelevanceMeasure(object):

    def set_text_collection(self, texts, language=consts.Language.ENGLISH):
        raise NotImplementedError()

    def relevance(self, keyphrase, text, synonimizer=None):
        raise NotImplementedError()


class ASTRelevanceMeasure(RelevanceMeasure):

    def __init__(self, ast_algorithm=consts.ASTAlgorithm.EASA, normalized=True, device=0):
        super(ASTRelevanceMeasure, self).__init__()
        self.ast_algorithm = ast_algorithm
        self.normalized = normalized
        self.device = device
        self.asts = []
        self._index = None

    def set_text_collection(self, texts, language=consts.Language.ENGLISH):
        """relevance.py:34-49: one AST per text; here all of them in one batched build."""
        self.texts = texts
        self.language = language
        self.asts = []
        self._index = None
        total_texts = len(texts)
        if self.ast_algorithm not in tuple(consts.ASTAlgorithm):
            # other registered engines (none ship in this package) go through the registry
            for i in range(total_texts):
                self.asts.append(base.AST.get_ast(utils.text_to_strings_collection(texts[i]),
                                                  self.ast_algorithm))
                logging.progress("Indexing texts with ASTs", i + 1, total_texts)
            logging.clear()
            return
        collections = [utils.text_to_strings_collection(text) for text in texts]
        if not collections:
            return
        packed = [asts_utils.pack_strings_collection(c) for c in collections]
        self._index = _capi.DeviceIndex(packed, [len(c) for c in collections], device=self.device)
        self.asts = [easa.EnhancedAnnotatedSuffixArray(c, _index=self._index, _doc=j)
                     for j, c in enumerate(collections)]

    def relevance(self, keyphrase, text, synonimizer=None):
        return self.asts[text].score(keyphrase, normalized=self.normalized, synonimizer=synonimizer)

    def relevance_table(self, prepared_keyphrases):
        """Scores of every prepared keyphrase against every text: float64 [n_texts, K]."""
        if self._index is None:
            return np.array([[ast.score(kp, normalized=self.normalized) for kp in prepared_keyphrases]
                             for ast in self.asts], dtype=np.float64)
        codes, off = _capi.pack_keyphrases(prepared_keyphrases)
        return self._index.score_table(codes, off, normalized=self.normalized)
