"""TEST INFRASTRUCTURE ONLY -- import shim for the real reference at /root/reference.

Used only in the build container (the GPU box has no /root/reference) by
oracle/make_golden.py and tests marked `needs_reference` to validate the oracle
restatement and to generate tests/golden/*.npz. Never imported by the product.

Recipe follows SURVEY.md Appendix A: stub the absent optional deps (h5py, cirq,
pubchempy, deprecation) and register a synthetic `openfermion` package whose
__path__ points at the read-only source tree, skipping the heavy top-level
__init__.
"""
import os
import sys
import types
from unittest import mock

REFERENCE_SRC = '/root/reference/src/openfermion'


def available():
    return os.path.isdir(REFERENCE_SRC)


class _Any(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return mock.MagicMock(name=f'{self.__name__}.{k}')


def load():
    """Returns the reference `openfermion` package (linalg imported first)."""
    if 'openfermion' in sys.modules and getattr(sys.modules['openfermion'], '_ofb_shim', False):
        return sys.modules['openfermion']
    if not available():
        raise RuntimeError('reference tree not present: ' + REFERENCE_SRC)
    for name in ('h5py', 'cirq', 'pubchempy', 'deprecation'):
        if name not in sys.modules:
            sys.modules[name] = _Any(name)
    sys.modules['deprecation'].deprecated = lambda *a, **k: (lambda f: f)
    pkg = types.ModuleType('openfermion')
    pkg.__path__ = [REFERENCE_SRC]
    pkg._ofb_shim = True
    sys.modules['openfermion'] = pkg
    import openfermion.linalg  # noqa: F401  (must be the first import)
    import openfermion.ops.operators  # noqa: F401
    import openfermion.transforms.opconversions  # noqa: F401
    import openfermion.hamiltonians.hubbard  # noqa: F401
    import openfermion.testing.testing_utils  # noqa: F401
    return pkg
