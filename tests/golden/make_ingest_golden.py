"""
make_ingest_golden.py -- golden vectors for the ingest path from the UNMODIFIED
reference (build container only):   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_ingest_golden.py

A synthetic two-component PSP file is written by exptool_b200.io.psp_io.write_psp and read back by the reference's
psp_io.Input (which validates writer and format); pattern.BarTransform runs on both components; the centring
lines of Fields.total_coefficients (potential.py:158-200, restated verbatim here because total_coefficients itself
cannot run at HEAD: eof.compute_coefficients raises on NumPy >= 2.2) run on the transformed
sets; eof.accumulate and spheresl.compute_coefficients_solitary give the coefficients of the transformed, centred sets.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim                      # noqa: E402
from exptool_b200 import synthetic as S         # noqa: E402
from exptool_b200.io import psp_io as my_psp    # noqa: E402

R = refshim.load()
eof, spheresl = R['eof'], R['spheresl']
with contextlib.redirect_stdout(io.StringIO()):
    import exptool.io.psp_io as ref_psp
    import exptool.analysis.pattern as ref_pattern


def snapshot_components(snap, float32):
    """star without particle indices, dark with them (both PSP record layouts) -- float32 files: no indices"""
    if float32:
        return [dict(name='star', data=snap['star']), dict(name='dark', data=snap['dark'])]
    dark = dict(snap['dark']); dark['id'] = np.arange(len(dark['m'])) + 1
    return [dict(name='star', data=snap['star']),
            dict(name='dark', data=dark, indexing=True, extra={'force': {'id': 'sphereSL'}})]


def run(name, seed, nd, nh, eparams, sparams, float32=False):
    snap = S.barred_snapshot(seed, nd, nh)
    with tempfile.TemporaryDirectory() as tmp:
        comps = snapshot_components(snap, float32)
        f = my_psp.write_psp(os.path.join(tmp, 'OUT.run.00001'), 0.125, comps, float32=float32)
        D = ref_psp.Input(f, 'star'); H = ref_psp.Input(f, 'dark')
        for k in ('m', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'potE'):
            want = snap['star'][k].astype(np.float32 if float32 else np.float64)
            assert np.array_equal(D.data[k], want), k
            assert np.array_equal(H.data[k], snap['dark'][k].astype(np.float32 if float32 else np.float64)), k
        assert D.time == 0.125 and set(D.header.keys()) == {'star', 'dark'}
        if float32:
            np.savez_compressed(os.path.join(HERE, name + '.npz'),
                                meta=json.dumps(dict(seed=seed, nd=nd, nh=nh, float32=True)),
                                star_x=D.data['x'][:64], dark_potE=H.data['potE'][:64])
            print('wrote', name)
            return
        DT = ref_pattern.BarTransform(D)
        HT = ref_pattern.BarTransform(H, bar_angle=DT.bar_angle)
        # potential.py:152-200, centering=True, mutual_center=False (verbatim arithmetic)
        ncenter = 10000
        rrank = (DT.data['x'] * DT.data['x'] + DT.data['y'] * DT.data['y'] + DT.data['z'] * DT.data['z']) ** 0.5
        cp = rrank.argsort()[0:ncenter]
        cd = [np.sum(DT.data[k][cp] * DT.data['m'][cp]) / np.sum(DT.data['m'][cp]) for k in 'xyz']
        ch = [np.sum(HT.data[k][cp] * HT.data['m'][cp]) / np.sum(HT.data['m'][cp]) for k in 'xyz']
        xd, yd, zd = [DT.data[k] - c for k, c in zip('xyz', cd)]
        xh, yh, zh = [HT.data[k] - c for k, c in zip('xyz', ch)]
        eof_file, sl_file, model_file = S.write_fixture_files(tmp, eof_params=eparams, sl_params=sparams, kind='smooth', seed=seed)
        with contextlib.redirect_stdout(io.StringIO()):
            potC, rfC, zfC, dC, potS, rfS, zfS, dS = eof.parse_eof(eof_file)
        rmin, rmax, numx, numy, mmax, norder, ascale, hscale, cmap, dens = eof.eof_params(eof_file)
        XMIN, XMAX, dX, YMIN, YMAX, dY = eof.set_table_params(RMAX=rmax, RMIN=rmin, ASCALE=ascale, HSCALE=hscale,
                                                              NUMX=numx, NUMY=numy, CMAP=cmap)
        cosd, sind = eof.accumulate(S.ParticleSet(xd, yd, zd, DT.data['m']), potC, potS, mmax, norder, XMIN, dX, YMIN, dY,
                                    numx, numy, ascale, hscale, cmap)
        with contextlib.redirect_stdout(io.StringIO()):
            Hs = R['particle'].holder(); Hs.xpos, Hs.ypos, Hs.zpos, Hs.mass = xh, yh, zh, HT.data['m']
            coef = spheresl.compute_coefficients_solitary(Hs, sl_file, model_file)
    np.savez_compressed(os.path.join(HERE, name + '.npz'),
                        meta=json.dumps(dict(seed=seed, nd=nd, nh=nh, eof_params=eparams, sl_params=sparams, kind='smooth')),
                        bar_angle=np.array(DT.bar_angle), cen_disk=np.array(cd), cen_halo=np.array(ch),
                        star_tx=DT.data['x'][:256], star_ty=DT.data['y'][:256], star_tvx=DT.data['vx'][:256],
                        star_tvy=DT.data['vy'][:256], dark_tx=HT.data['x'][:256], dark_ty=HT.data['y'][:256],
                        cos=cosd, sin=sind, coef=np.asarray(coef, dtype=np.float64))
    print('wrote', name, 'bar_angle', DT.bar_angle, 'centre', cd, ch)


if __name__ == '__main__':
    run('ingest_small', 41, 24000, 24000, dict(mmax=4, numx=48, numy=32, nmax=8, norder=6), dict(lmax=4, nmax=6, numr=400))
    run('ingest_f32', 42, 500, 400, None, None, float32=True)
