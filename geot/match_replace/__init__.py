"""``geot.match_replace`` -- the FX rewriter entry point under the reference's module path
(``/root/reference/geot/match_replace/__init__.py``), served by ``geot_b200.match_replace``."""
from geot_b200.match_replace import pattern_transform  # noqa: F401
from geot_b200 import coo_to_csr  # noqa: F401
