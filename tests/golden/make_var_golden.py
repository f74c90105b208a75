"""
make_var_golden.py -- golden vectors for the VAR (jackknife partitions) branch of eof.accumulate (eof.py:554-574,
617-637), produced by the UNMODIFIED reference with NumPy's global generator seeded (np.random.seed(SEED)) right
before the call, on the particles and tables of the eof_small_random_cmap1 case, both container branches.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_var_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as G                          # noqa: E402
from exptool_b200 import synthetic as S          # noqa: E402

SEED, NVAR = 77, 4


def main():
    d = np.load(os.path.join(HERE, 'eof_small_random_cmap1.npz'))
    meta = json.loads(str(d['meta']))
    with tempfile.TemporaryDirectory() as tmp:
        f, tabs, g = G.eof_setup(tmp, meta['eof_params'], meta['kind'], meta['seed'])
        potC, potS = tabs[0], tabs[4]
        a = (potC, potS, g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'],
             g['ascale'], g['hscale'], g['cmap'])
        H = G.particle.holder(); H.xpos, H.ypos, H.zpos, H.mass = d['x'], d['y'], d['z'], d['m']
        np.random.seed(SEED)
        ch, sh, c2h, s2h = G.eof.accumulate(H, *a, VAR=NVAR)
        P = S.ParticleSet(d['x'], d['y'], d['z'], d['m'])
        np.random.seed(SEED)
        cd, sd, c2d, s2d = G.eof.accumulate(P, *a, VAR=NVAR)
    np.savez_compressed(os.path.join(HERE, 'eof_var.npz'), meta=json.dumps(dict(case='eof_small_random_cmap1', seed=SEED, nvar=NVAR)),
                        cos=ch, sin=sh, cos2_holder=c2h, sin2_holder=s2h, cos2_data=c2d, sin2_data=s2d)
    print('wrote eof_var', c2h.shape, float(np.max(np.abs(c2h - c2d))))


if __name__ == '__main__':
    main()
