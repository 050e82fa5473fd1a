# -*- coding: utf-8 -*-
"""
Stokes / polarised amplitude conversion (reference: tricolour/stokes.py).

``stokes_corr_map`` is host dictionary logic; the per-sample arithmetic of
``polarised_intensity`` / ``unpolarised_intensity`` (complex128 evaluation of
a*(s1*v1 + s2*v2), |.|, accumulate, sqrt, narrow to complex64) runs on the GPU.
"""
import ctypes

import numpy as np

from . import _cabi
from ._cabi import check, ptr, context_for

# Enumeration of stokes, linear and circular correlations used in Measurement
# Set 2.0 as per Stokes.h in casacore (tricolour/stokes.py:12-25)
STOKES_TYPES = {
    'I': 1, 'Q': 2, 'U': 3, 'V': 4,
    'RR': 5, 'RL': 6, 'LR': 7, 'LL': 8,
    'XX': 9, 'XY': 10, 'YX': 11, 'YY': 12,
}

# (corr1, corr2, a, s1, s2): stokes = a*(s1*corr1 + s2*corr2)  (stokes.py:29-39)
stokes_deps = {
    'I': [('XX', 'YY', 0.5 + 0.0j, 1, 1), ('RR', 'LL', 0.5 + 0.0j, 1, 1)],
    'Q': [('XX', 'YY', 0.5 + 0.0j, 1, -1), ('RL', 'LR', 0.5 + 0.0j, 1, 1)],
    'U': [('XY', 'YX', 0.5 + 0.0j, 1, 1), ('RL', 'LR', 0.0 - 0.5j, 1, -1)],
    'V': [('XY', 'YX', 0.0 - 0.5j, 1, -1), ('RR', 'LL', 0.5 + 0.0j, 1, -1)],
}
stokes_deps = {k: [(STOKES_TYPES[c1], STOKES_TYPES[c2], a, s1, s2)
                   for (c1, c2, a, s1, s2) in deps]
               for k, deps in stokes_deps.items()}


def stokes_corr_map(corr_types):
    """
    Map describing how to combine visibility correlations into stokes
    parameters: ``{stokes: (c1, c2, a, s1, s2)}`` with
    ``stokes = a*(s1*vis[:,:,c1] + s2*vis[:,:,c2])`` (stokes.py:42-76).
    A later dependency overwrites an earlier one if both are available.
    """
    corr_type_set = set(corr_types)
    corr_maps = {}
    for stokes, deps in stokes_deps.items():
        for (corr1, corr2, alpha, sign1, sign2) in deps:
            if len(corr_type_set.intersection((corr1, corr2))) == 2:
                c1 = corr_types.index(corr1)
                c2 = corr_types.index(corr2)
                corr_maps[stokes] = (c1, c2, alpha, sign1, sign2)
    return corr_maps


def _terms(stokes):
    idx = np.array([[t[0], t[1]] for t in stokes], np.int32).reshape(-1, 2)
    coef = np.array([[complex(t[2]).real, complex(t[2]).imag, float(t[3]), float(t[4])]
                     for t in stokes], np.float64).reshape(-1, 4)
    return np.ascontiguousarray(idx), np.ascontiguousarray(coef)


def _hp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _vis_in(vis):
    if _cabi.is_device_array(vis):
        import torch
        if vis.dtype != torch.complex64:
            raise TypeError("device visibilities must be complex64")
        return vis.contiguous(), None
    v = np.asarray(vis)
    if not np.iscomplexobj(v):
        raise TypeError("visibilities must be complex")
    back = None if v.dtype == np.complex64 else v.dtype
    return np.ascontiguousarray(v, dtype=np.complex64), back


def _out_like(vis, shape):
    if _cabi.is_device_array(vis):
        import torch
        return torch.empty(shape, dtype=torch.complex64, device=vis.device)
    return np.empty(shape, np.complex64)


def polarised_intensity(vis, stokes_pol):
    r"""
    :math:`\sqrt{Q^2 + U^2 + V^2}` from visibilities of shape
    :code:`(row, chan, corr)` and tuples :code:`(c1,c2,a,s1,s2)` (see
    :func:`stokes_corr_map`).  Returns shape :code:`(row, chan, 1)` in the
    dtype of ``vis`` with a zero imaginary part (stokes.py:157-209).
    """
    v, back = _vis_in(vis)
    nrow, nchan, ncorr = (int(s) for s in v.shape)
    idx, coef = _terms(stokes_pol)
    out = _out_like(v, (nrow, nchan, 1))
    ctx, space = context_for(v)
    check(_cabi.load().tc_polarised_intensity(ctx.handle, ptr(v), nrow * nchan, ncorr, _hp(idx),
                                              _hp(coef), idx.shape[0], ptr(out), space))
    return out if back is None else out.astype(back)


def unpolarised_intensity(vis, stokes_unpol, stokes_pol):
    r"""
    :math:`I - \sqrt{Q^2 + U^2 + V^2}` (stokes.py:79-154).  ``stokes_unpol`` must
    hold exactly one entry and ``stokes_pol`` at least one.
    """
    if not len(stokes_unpol) == 1:
        raise ValueError("There should be exactly one entry "
                         "for unpolarised stokes (stokes_unpol)")
    if not len(stokes_pol) > 0:
        raise ValueError("No entries for polarised stokes (stokes_pol)")
    v, back = _vis_in(vis)
    nrow, nchan, ncorr = (int(s) for s in v.shape)
    ui, uc = _terms(stokes_unpol)
    pi, pc = _terms(stokes_pol)
    out = _out_like(v, (nrow, nchan, 1))
    ctx, space = context_for(v)
    check(_cabi.load().tc_unpolarised_intensity(ctx.handle, ptr(v), nrow * nchan, ncorr, _hp(ui),
                                                _hp(uc), ui.shape[0], _hp(pi), _hp(pc),
                                                pi.shape[0], ptr(out), space))
    return out if back is None else out.astype(back)
