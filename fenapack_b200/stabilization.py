"""Streamline-diffusion stabilisation parameter -- drop-in for fenapack/stabilization.py.

Assembly-side helper (host): the reference JIT-compiles a DOLFIN ``Expression`` whose value in
a cell with diameter ``h`` is (stabilization.py:64-67)

    PE    = |w| h rho / (2 nu)                     (mesh Peclet number)
    delta = h (1 - 1/PE) / (2 |w|)   if PE > 1,   0 otherwise

and which multiplies ``inner(dot(grad(u), w), dot(grad(v), w))*dx`` in the stabilised 00-block
``a_pc`` handed to the velocity AMG (demo_navier-stokes-pcd.py:123-125; on the device this is
operator ``FNP_MAT_P00``).  Here the returned object is a DG0 evaluator fed with per-cell
arrays by the host assembler (``eval_cells``); it carries no DOLFIN dependency.
"""
from __future__ import annotations

import numpy as np

__all__ = ["StabilizationParameterSD", "streamline_diffusion_parameter"]


def streamline_diffusion_parameter(h, wind_norm, viscosity, density=1.0):
    """``delta`` per cell from cell diameters ``h``, wind norms ``|w|`` and (scalar or per-cell)
    viscosity / density -- the arithmetic of ``StabilizationParameterSD::eval``
    (reference stabilization.py:40-67)."""
    h = np.asarray(h, dtype=np.float64)
    wn = np.asarray(wind_norm, dtype=np.float64)
    pe = 0.5 * wn * h * np.asarray(density, dtype=np.float64) / np.asarray(viscosity, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(pe > 1.0, 0.5 * h * (1.0 - 1.0 / pe) / wn, 0.0)


class _CellwiseSD(object):
    """DG0 stand-in for the compiled Expression: ``wind`` is a callable ``cells -> [ncells, d]``
    (the wind at the evaluation point of each cell) or such an array; ``viscosity`` / ``density``
    scalars, arrays or callables of the same kind."""

    def __init__(self, wind, viscosity, density):
        self.wind, self.viscosity, self.density = wind, viscosity, density

    @staticmethod
    def _values(f, cells):
        return f(cells) if callable(f) else f

    def eval_cells(self, h, cells=None):
        w = np.asarray(self._values(self.wind, cells), dtype=np.float64)
        return streamline_diffusion_parameter(h, np.linalg.norm(w, axis=-1), self._values(self.viscosity, cells),
                                              self._values(self.density, cells))


def StabilizationParameterSD(wind, viscosity, density=None):
    """Same signature as the reference (stabilization.py:84-118): ``wind`` a vector field,
    ``viscosity`` and optional ``density`` scalar fields."""
    if density is None:
        density = 1.0
    return _CellwiseSD(wind, viscosity, density)
