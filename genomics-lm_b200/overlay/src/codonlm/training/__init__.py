"""Overlay of `src.codonlm.training`: the reference's trainer modules, with `objectives` served by codonlm_b200."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
