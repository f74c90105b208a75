"""
Ingest path: PSP reader / writer, BarTransform, centring and
Fields.total_coefficients against golden vectors from the unmodified reference (tests/golden/make_ingest_golden.py).
"""
import os
import sys

import numpy as np
import pytest

from helpers import load_golden, relerr, S, O, GOLDEN

sys.path.insert(0, GOLDEN)
from exptool_b200.io import psp_io, particle


def _components(snap, float32):
    if float32:
        return [dict(name='star', data=snap['star']), dict(name='dark', data=snap['dark'])]
    dark = dict(snap['dark']); dark['id'] = np.arange(len(dark['m'])) + 1
    return [dict(name='star', data=snap['star']),
            dict(name='dark', data=dark, indexing=True, extra={'force': {'id': 'sphereSL'}})]


def _write(tmp_path, meta):
    snap = S.barred_snapshot(meta['seed'], meta['nd'], meta['nh'])
    f = psp_io.write_psp(str(tmp_path / 'OUT.run.00001'), 0.125, _components(snap, meta.get('float32', False)),
                         float32=meta.get('float32', False))
    return snap, f


def test_psp_reader_round_trip_f64_with_indexing(tmp_path):
    d, meta = load_golden('ingest_small')
    snap, f = _write(tmp_path, meta)
    hdr = psp_io.Input(f)
    assert set(hdr.header.keys()) == {'star', 'dark'} and hdr.time == 0.125
    assert hdr.header['dark']['parameters']['indexing'] is True and hdr.header['dark']['force']['id'] == 'sphereSL'
    D = particle.Input(f, comp='star'); H = particle.Input(f, comp='dark')
    for k in ('m', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'potE'):
        assert np.array_equal(D.data[k], snap['star'][k]) and np.array_equal(H.data[k], snap['dark'][k])
    assert np.array_equal(H.data['id'], np.arange(meta['nh']) + 1) and 'id' not in D.data
    L = particle.Input(f, comp='dark', legacy=True)
    assert np.array_equal(L.xpos, snap['dark']['x']) and np.array_equal(L.id, H.data['id'])
    with pytest.raises(IOError):
        psp_io.Input(f, comp='gas')


def test_psp_reader_float32(tmp_path):
    d, meta = load_golden('ingest_f32')
    snap, f = _write(tmp_path, meta)
    D = psp_io.Input(f, comp='star'); H = psp_io.Input(f, comp='dark')
    assert D.data['x'].dtype == np.float32
    assert np.array_equal(D.data['x'][:64], d['star_x']) and np.array_equal(H.data['potE'][:64], d['dark_potE'])


def test_oracle_transforms_match_reference():
    d, meta = load_golden('ingest_small')
    snap = S.barred_snapshot(meta['seed'], meta['nd'], meta['nh'])
    st, dk = snap['star'], snap['dark']
    ang = -1. * O.bar_fourier_angle(st['x'], st['y'], maxr=1.)
    assert abs(ang - float(d['bar_angle'])) < 1e-13
    tx, ty = O.bar_rotate(st['x'], st['y'], ang)
    hx, hy = O.bar_rotate(dk['x'], dk['y'], ang)
    assert relerr(tx[:256], d['star_tx']) < 1e-14 and relerr(hy[:256], d['dark_ty']) < 1e-14
    cd = O.inner_center(tx, ty, st['z'], tx, ty, st['z'], st['m'])
    ch = O.inner_center(tx, ty, st['z'], hx, hy, dk['z'], dk['m'])
    assert relerr(cd, d['cen_disk']) < 1e-12 and relerr(ch, d['cen_halo']) < 1e-12


@pytest.mark.gpu
def test_bar_transform_and_total_coefficients_gpu(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU tests selected but CUDA is not available')
    from exptool_b200.analysis import pattern
    from exptool_b200.basis import potential
    from exptool_b200 import ops
    d, meta = load_golden('ingest_small')
    snap, f = _write(tmp_path, meta)
    D = particle.Input(f, comp='star'); H = particle.Input(f, comp='dark')
    DT = pattern.BarTransform(D)
    assert abs(DT.bar_angle - float(d['bar_angle'])) < 1e-12
    HT = pattern.BarTransform(H, bar_angle=DT.bar_angle)
    assert relerr(DT.data['x'][:256], d['star_tx']) < 1e-13 and relerr(DT.data['y'][:256], d['star_ty']) < 1e-13
    assert relerr(DT.data['vx'][:256], d['star_tvx']) < 1e-13 and relerr(DT.data['vy'][:256], d['star_tvy']) < 1e-13
    assert relerr(HT.data['x'][:256], d['dark_tx']) < 1e-13 and DT.time == 0.125 and DT.comp == 'star'
    # inner centre of mass: radix select + sums vs the reference's argsort
    cd = ops.inner_center_of_mass(DT.data['x'], DT.data['y'], DT.data['z'], DT.data['m'], 10000)
    assert relerr(cd, d['cen_disk']) < 1e-11
    small = ops.inner_center_of_mass(DT.data['x'][:777], DT.data['y'][:777], DT.data['z'][:777], DT.data['m'][:777], 10000)
    want = O.inner_center(DT.data['x'][:777], DT.data['y'][:777], DT.data['z'][:777], DT.data['x'][:777],
                          DT.data['y'][:777], DT.data['z'][:777], DT.data['m'][:777])
    assert relerr(small, want) < 1e-11                     # ncenter > n: all particles
    for k in (1, 2, 500):
        got = ops.inner_center_of_mass(D.data['x'], D.data['y'], D.data['z'], D.data['m'], k)
        want = O.inner_center(D.data['x'], D.data['y'], D.data['z'], D.data['x'], D.data['y'], D.data['z'], D.data['m'], k)
        assert relerr(got, want) < 1e-11, k
    # Fields.total_coefficients: ingest -> bar frame -> centring -> both accumulations
    eof_file, sl_file, model_file = S.write_fixture_files(str(tmp_path), eof_params=meta['eof_params'],
                                                          sl_params=meta['sl_params'], kind=meta['kind'], seed=meta['seed'])
    F = potential.Fields(f, eof_file, sl_file, model_file, transform=True, centering=True, mutual_center=False, verbose=0)
    F.total_coefficients()
    assert abs(F.bar_angle - float(d['bar_angle'])) < 1e-12 and F.time == 0.125 and F.halofac == 1.0
    assert relerr([F.xcen_disk, F.ycen_disk, F.zcen_disk], d['cen_disk']) < 1e-11
    assert relerr([F.xcen_halo, F.ycen_halo, F.zcen_halo], d['cen_halo']) < 1e-11
    assert relerr(F.EOF.cos, d['cos']) < 1e-10 and relerr(F.EOF.sin, d['sin']) < 1e-10
    assert relerr(F.SL.expcoef, d['coef']) < 1e-10
    assert F.EOF.nbodies == meta['nd'] and F.SL.nbodies == meta['nh']
    Fm = potential.Fields(f, eof_file, sl_file, model_file, transform=False, centering=True, mutual_center=True, verbose=0)
    Fm.total_coefficients()
    want = O.inner_center(D.data['x'], D.data['y'], D.data['z'], D.data['x'], D.data['y'], D.data['z'], D.data['m'])
    assert relerr([Fm.xcen_disk, Fm.ycen_disk, Fm.zcen_disk], want) < 1e-11
    assert (Fm.xcen_halo, Fm.ycen_halo, Fm.zcen_halo) == (Fm.xcen_disk, Fm.ycen_disk, Fm.zcen_disk)
    F.prep_tables()                                        # and the field is usable downstream
    out = F.return_forces_cart(0.01, 0.002, 0.0005)
    assert len(out) == 8 and all(np.isfinite(float(v)) for v in out)


def test_spl_split_file_reader(tmp_path):
    """SPL split snapshots (spl_io.py:25-270): master file + per-process subfiles; round trip, particle.Input dispatch,
    and -- where the reference is mounted -- the reference's own reader on the same files."""
    import contextlib
    import io
    from exptool_b200.io import spl_io
    d, meta = load_golden('ingest_small')
    snap = S.barred_snapshot(meta['seed'], meta['nd'], meta['nh'])
    comps = _components(snap, False)          # 'dark' carries ids (indexing) and an extra header entry
    f = spl_io.write_spl(str(tmp_path / 'SPL.run.00001'), 0.25, comps, nprocs=3, float32=True)
    with contextlib.redirect_stdout(io.StringIO()) as out:
        hdr = spl_io.Input(f)
    assert set(hdr.header.keys()) == {'star', 'dark'} and hdr.time == 0.25 and 'Found 2 components' in out.getvalue()
    D = particle.Input(f, comp='star'); H = particle.Input(f, comp='dark')
    assert D.style == 'SPL' and len(spl_io.Input(f, comp='dark').subfiles) == 3
    for k in ('m', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'potE'):
        assert D.data[k].dtype == np.float32
        assert np.array_equal(D.data[k], snap['star'][k].astype(np.float32))
        assert np.array_equal(H.data[k], snap['dark'][k].astype(np.float32))
    assert np.array_equal(H.data['id'], np.arange(meta['nh']) + 1)
    with pytest.raises(IOError):
        spl_io.Input(f, comp='gas')
    if os.path.isdir('/root/reference/exptool'):
        from oracle import refshim
        import importlib
        refshim.load()
        ref_spl = importlib.import_module('exptool.io.spl_io')
        with contextlib.redirect_stdout(io.StringIO()):
            R = ref_spl.Input(f, comp='dark')
        assert R.subfiles == spl_io.Input(f, comp='dark').subfiles
        for k in R.data.keys():
            assert np.array_equal(R.data[k], H.data[k]), k
