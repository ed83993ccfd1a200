"""Importable alias of the package directory ``lns-latent-neural-pde-solver_b200/`` (a hyphenated directory name
cannot be imported directly).  ``import lns_b200.ops`` resolves to ``lns-latent-neural-pde-solver_b200/ops.py``."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "lns-latent-neural-pde-solver_b200")
__path__.append(_PKG_DIR)
PKG_DIR = _PKG_DIR
REPO_ROOT = _os.path.dirname(_PKG_DIR)

from .version import __version__  # noqa: E402,F401
