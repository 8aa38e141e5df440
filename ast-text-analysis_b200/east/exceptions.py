# -*- coding: utf-8 -*-
"""Exception types of the EAST API (mirror of east/exceptions.py:4-64)."""


class EastException(Exception):
    """Base class; subclasses define msg_fmt, formatted with the constructor's kwargs."""

    msg_fmt = "An unknown exception occurred."

    def __init__(self, message=None, **kwargs):
        self.kwargs = kwargs
        if not message:
            try:
                message = self.msg_fmt % kwargs
            except (KeyError, TypeError):
                message = self.msg_fmt
        super(EastException, self).__init__(message)

    def format_message(self):
        return str(self)


class NotFoundException(EastException):
    msg_fmt = "Not found."


class NoSuchASTAlgorithm(NotFoundException):
    msg_fmt = "There is no AST construction algorithm with name `%(name)s`."


class EmptyStringsCollectionException(EastException):
    msg_fmt = "The input strings collection is empty."


class DeviceError(EastException):
    """Raised when the B200 engine is unavailable or a CUDA call fails (no CPU fallback)."""

    msg_fmt = "B200 engine failure: %(reason)s"
