"""
make_outcoef_golden.py -- pin exptool_b200.io.outcoef to the reference's OWN fixture.

The reference ships one EXP coefficient file with its tests (exptool/tests/outcoef.star.run0.dat, 1.1 MB,
YAML-style cylinder records).  This script (build container only, needs /root/reference) reads it with the
UNMODIFIED reference reader (exptool/io/outcoef.py OutCoef) and freezes
  * outcoef_star_head.dat : the first 6 records of the fixture, byte for byte (2.7 kB of data, not source);
  * outcoef_star.npz      : T and coefs of those 6 records as the reference reader returns them, plus T of all
                            records and the SHA-256 of the full coefficient array, for the live comparison.
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_outcoef_golden.py
"""
import contextlib
import hashlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim                      # noqa: E402

FIXTURE = '/root/reference/exptool/tests/outcoef.star.run0.dat'
refshim.load()
with contextlib.redirect_stdout(io.StringIO()):
    import exptool.io.outcoef as ref_outcoef
    full = ref_outcoef.OutCoef(FIXTURE)
nrec = 6
raw = open(FIXTURE, 'rb').read()
pos = 0
for _ in range(nrec):
    magic, ln = np.frombuffer(raw[pos:pos + 8], dtype='<u4')
    mmax, nmax = full.coefs.shape[2] - 1, full.coefs.shape[3]
    pos += 8 + int(ln) + 8 * (2 * mmax + 1) * nmax
head = os.path.join(HERE, 'outcoef_star_head.dat')
with open(head, 'wb') as f:
    f.write(raw[:pos])
with contextlib.redirect_stdout(io.StringIO()):
    h = ref_outcoef.OutCoef(head)
assert h.coefs.shape[0] == nrec and np.array_equal(h.coefs, full.coefs[:nrec])
np.savez_compressed(os.path.join(HERE, 'outcoef_star.npz'), T_head=h.T, coefs_head=h.coefs, T_full=full.T,
                    shape_full=np.array(full.coefs.shape),
                    sha_full=np.array(hashlib.sha256(np.ascontiguousarray(full.coefs).tobytes()).hexdigest()))
print('wrote', head, os.path.getsize(head), 'bytes;', full.coefs.shape, 'records in the fixture')
