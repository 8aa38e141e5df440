# -*- coding: utf-8 -*-
"""Abstract AST and the algorithm registry (mirror of east/asts/base.py:10-46).

AST.get_ast(strings_collection, ast_algorithm) instantiates the first concrete subclass whose
__algorithm__ matches; engines register simply by being imported (east/__init__.py).
"""
import abc
import inspect

from east import consts
from east import exceptions
from east import utils


class AST(object, metaclass=abc.ABCMeta):

    @staticmethod
    def get_ast(strings_collection, ast_algorithm="easa"):
        for ast_cls in utils.itersubclasses(AST):
            if not inspect.isabstract(ast_cls) and ast_algorithm == getattr(ast_cls, "__algorithm__", None):
                return ast_cls(strings_collection)
        raise exceptions.NoSuchASTAlgorithm(name=ast_algorithm)

    def __init__(self, strings_collection):
        if not strings_collection:
            raise exceptions.EmptyStringsCollectionException()

    @abc.abstractmethod
    def score(self, query, normalized=True, synonimizer=None, return_suffix_scores=False):
        """Matching score of the query against the annotated suffix structure."""

    def traverse(self, callback, order=consts.TraversalOrder.DEPTH_FIRST_PRE_ORDER):
        if order == consts.TraversalOrder.DEPTH_FIRST_PRE_ORDER:
            self.traverse_depth_first_pre_order(callback)
        elif order == consts.TraversalOrder.DEPTH_FIRST_POST_ORDER:
            self.traverse_depth_first_post_order(callback)
        elif order == consts.TraversalOrder.BREADTH_FIRST:
            self.traverse_breadth_first(callback)

    @abc.abstractmethod
    def traverse_depth_first_pre_order(self, callback):
        """Visit the nodes depth first, parents before children."""

    @abc.abstractmethod
    def traverse_depth_first_post_order(self, callback):
        """Visit the nodes depth first, children before parents."""

    @abc.abstractmethod
    def traverse_breadth_first(self, callback):
        """Visit the nodes level by level."""
