"""``geot`` -- the reference package's import name, served by ``geot_b200`` (the B200-native build).

``import geot; geot.index_scatter(...)`` call sites of the reference (``/root/reference/geot/__init__.py:4-9``)
work unchanged: every public name of the reference package resolves to the ``geot_b200`` operator of the same
name, and ``torch.ops.geot.*`` are the operators ``geot_b200/_C.so`` registers.  This package holds no code of
its own.  (Do not install it next to the reference's own ``geot``: both register the ``geot::`` op namespace.)
"""
import geot_b200 as _impl
from geot_b200 import (index_scatter, gather_scatter, gather_weight_scatter, mh_spmm, mh_spmm_transposed,  # noqa: F401
                       csr_gws, coo_to_csr, sddmm_coo_impl, pattern_transform, format_preprocess)

# `import geot.csr_gws`, `import geot.index_scatter` ... (reference call sites: test/compile/test_csr_gws.py:6) resolve to
# the geot_b200 module of the same name; as in the reference, the attribute `geot.csr_gws` stays the operator.
import importlib as _importlib
import sys as _sys
for _name in ("index_scatter", "gather_scatter", "gather_weight_scatter", "mh_spmm", "csr_gws", "format_preprocess"):
    _sys.modules[__name__ + "." + _name] = _importlib.import_module("geot_b200." + _name)
del _importlib, _sys, _name

__version__ = _impl.__version__
__all__ = ["index_scatter", "gather_scatter", "gather_weight_scatter", "mh_spmm", "mh_spmm_transposed", "csr_gws",
           "coo_to_csr"]
