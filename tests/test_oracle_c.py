"""The C restatement (oracle/bfe_oracle.c) agrees with the NumPy oracle, which is pinned to the reference."""
import numpy as np
import pytest

from helpers import load_golden, relerr, eof_tables, sl_tables, eof_geo_args, O, S
from oracle import oracle_c as OC


@pytest.mark.parametrize('name', ['eof_small_random_cmap1', 'eof_small_random_cmap0', 'eof_std_smooth'])
def test_c_eof_matches_golden(name):
    d, meta = load_golden(name)
    p, T, g = eof_tables(meta)
    c, s = OC.eof_accumulate(d['x'], d['y'], d['z'], d['m'], T['potC'], T['potS'], g)
    assert relerr(c, d['cos']) < 1e-12 and relerr(s, d['sin']) < 1e-12
    nf = meta['nforce']
    out = OC.eof_force(d['x'][:nf], d['y'][:nf], d['z'][:nf], d['cos'], d['sin'], T, g)
    win = OC.eof_force(d['x'][:nf], d['y'][:nf], d['z'][:nf], d['cos'], d['sin'], T, g, m1=1, m2=2)
    for i in range(6):
        assert relerr(out[i], d['full'][i]) < 1e-12, i
        assert relerr(win[i], d['win12'][i]) < 1e-12, i


@pytest.mark.parametrize('name', ['sl_small_random_cmap1', 'sl_small_random_cmap0', 'sl_std_l4', 'sl_std_l6'])
def test_c_sl_matches_golden(name):
    d, meta = load_golden(name)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    c = OC.sl_accumulate(d['x'], d['y'], d['z'], d['m'], p['lmax'], p['nmax'], ev, ef, xi, p0, p['cmap'], p['scale'])
    assert relerr(c, d['coef']) < 1e-12
    c2 = OC.sl_accumulate(d['x'], d['y'], d['z'], d['m'], p['lmax'], p['nmax'], ev, ef, xi, p0, p['cmap'], p['scale'],
                          no_odd=True)
    assert relerr(c2, d['coef_noodd']) < 1e-12


def test_c_matches_numpy_oracle_20k():
    meta = dict(eof_params={}, kind='smooth', seed=0)
    p, T, g = eof_tables(meta)
    x, y, z, m = S.exponential_disc(20000, 99)
    c, s = OC.eof_accumulate(x, y, z, m, T['potC'], T['potS'], g)
    co, so = O.eof_accumulate(x, y, z, m, T['potC'], T['potS'], g['mmax'], g['norder'], *eof_geo_args(g), g['ascale'],
                              g['hscale'], g['cmap'])
    assert relerr(c, co) < 1e-12 and relerr(s, so) < 1e-12
    assert OC.threads() >= 1
