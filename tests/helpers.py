"""Shared test helpers: load golden cases, rebuild their tables, error norms."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from exptool_b200 import synthetic as S      # noqa: E402
from oracle import oracle_np as O            # noqa: E402

GOLDEN = os.path.join(HERE, 'golden')


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    meta = json.loads(str(d['meta']))
    return d, meta


def relerr(a, b):
    """max|a-b| / max|b| -- the norm of BASELINE.md's parity gates."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    if den == 0.0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / den)


def eof_tables(meta):
    """(params, tables, geo) for a golden case -- in-memory, no file round trip."""
    p, T = S.make_eof_tables(meta['eof_params'], kind=meta['kind'], seed=meta['seed'])
    XMIN, XMAX, dX, YMIN, YMAX, dY = O.eof_set_table_params(RMAX=p['rmax'], RMIN=p['rmin'], ASCALE=p['ascale'],
                                                            HSCALE=p['hscale'], NUMX=p['numx'], NUMY=p['numy'],
                                                            CMAP=p['cmap'])
    g = dict(XMIN=XMIN, dX=dX, YMIN=YMIN, dY=dY, numx=p['numx'], numy=p['numy'], mmax=p['mmax'],
             norder=p['norder'], ascale=p['ascale'], hscale=p['hscale'], cmap=p['cmap'])
    return p, T, g


def sl_tables(meta, seed_offset=0):
    """(params, ev, ef, xi, p0, d0) for a golden case."""
    p, ev, ef = S.make_sl_tables(meta['sl_params'], kind=meta['kind'], seed=meta['seed'] + seed_offset)
    R1, D1, P1 = hernquist_model_columns(p['scale'])
    xi, r, p0, d0 = O.sl_init_table(R1, D1, P1, p['numr'], p['rmin'], p['rmax'], p['cmap'], p['scale'])
    return p, ev, ef, xi, p0, d0


def hernquist_model_columns(a, tmpdir=None):
    """Write + read the text model file exactly as the readers do (skip 5 lines)."""
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        f = S.write_hernquist_model(os.path.join(tmp, 'm'), a=a)
        A = np.genfromtxt(f, comments='!', skip_header=5)
    return A[:, 0], A[:, 1], A[:, 3]


def eof_geo_args(g):
    return (g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'])


def build_field(meta, d, oracle=True):
    """FrozenField for the field_* golden cases (oracle container)."""
    pe, T, g = eof_tables(meta)
    ps, ev, ef, xi, p0, d0 = sl_tables(meta, seed_offset=1)
    F = O.FrozenField(cos=d['cos'], sin=d['sin'], potC=T['potC'], rforceC=T['rforceC'], zforceC=T['zforceC'],
                      potS=T['potS'], rforceS=T['rforceS'], zforceS=T['zforceS'],
                      XMIN=g['XMIN'], dX=g['dX'], YMIN=g['YMIN'], dY=g['dY'], numx=g['numx'], numy=g['numy'],
                      mmax=g['mmax'], norder=g['norder'], ascale=g['ascale'], hscale=g['hscale'], cmapdisk=g['cmap'],
                      halofac=meta['halofac'], expcoef=d['coef'], xihalo=xi, p0halo=p0, d0halo=d0,
                      cmaphalo=ps['cmap'], scalehalo=ps['scale'], lmaxhalo=ps['lmax'], nmaxhalo=ps['nmax'],
                      evtablehalo=ev, eftablehalo=ef)
    return F
