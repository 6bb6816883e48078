"""Top-level alias so that ``import VMATlibrary`` picks up the B200 implementation.

The reference's file of this name is an abandoned stub (never imported, cannot run); the
names its driver scripts actually use are provided by vmatwpencode_b200.VMATlibrary.
"""
import sys as _sys

from vmatwpencode_b200 import VMATlibrary as _impl

_sys.modules[__name__] = _impl
