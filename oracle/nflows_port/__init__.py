"""CPU restatement of the subset of nflows 0.14 used by sbi (test infrastructure).

Module layout mirrors nflows so that ``oracle.ref_shim`` can alias it as ``nflows``.
"""
from . import utils, distributions, transforms, flows, nn  # noqa: F401
