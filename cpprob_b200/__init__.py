"""cpprob_b200 — B200-native sequential importance sampling behind CPProb's API.

The product is the CUDA engine in cpprob_b200/csrc (C ABI: include/cpprob_sis.h) and the C++14 host
API in include/cpprob.  This package only carries the build recipe and the ctypes plumbing used by
the tests and bench.py.
"""
from . import capi  # noqa: F401
from .capi import Engine, SisError  # noqa: F401

__all__ = ["capi", "Engine", "SisError"]
