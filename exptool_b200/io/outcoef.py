"""
outcoef.py -- EXP `outcoef.*` coefficient time-series files (exptool/io/outcoef.py), reader AND writer.

SURVEY.md section 8(f) rank 4: the consumers of a coefficient time series (BASELINE.json configs[4]) read EXP's
own wire format, so a series accumulated on the GPU (exptool_b200.parallel.accumulate_series) can be handed to
them unchanged.  Host-side NumPy, like the reference; records are parsed with bulk reads instead of the
reference's per-row `np.fromfile` calls.

Formats (Appendix B.6; little-endian):
  Cylinder, YAML  : per record  u4 magic 202004387, u4 len, YAML{time, mmax, nmax, ...}; then for m = 0..mmax
                    f8[nmax] cosine row and, if m > 0, f8[nmax] sine row                 (outcoef.py:339-411)
  SphereSL, YAML  : per record  u4 magic 202004386, u4 len, YAML{time, lmax, nmax, ...}; then nmax rows of
                    f8[(lmax+1)^2]  -> coefs[t, :, n]                                      (outcoef.py:264-336)
  Cylinder, old   : per record  f8 time, u4 mmax, u4 nmax, then the same rows              (outcoef.py:135-193)
  SphereSL, old   : per record  S64 id ('Sphere SL' inside), f8 time, f8 scale, u4 nmax, u4 lmax, then nmax
                    rows of f8[(lmax+1)^2]                                                 (outcoef.py:196-260)

`OutCoef(filename)` mirrors the reference class: attributes `.basis` ('Cylinder' / 'SphereSL'), `.T` (times),
`.coefs` ([t, 2, m, n] cylinder: 0 cosine / 1 sine, m = 0 sine row zero; [t, (lmax+1)^2, n] sphere).
"""
import os

import numpy as np

CYL_MAGIC = 202004387
SPH_MAGIC = 202004386


class EOF_Object(object):
    """eof object carrying one cylinder record (outcoef.py:27-37)"""
    time = None
    dump = None
    comp = None
    nbodies = None
    mmax = None
    nmax = None
    eof_file = None
    cos = None
    sin = None


def _yaml_load(raw):
    import yaml
    return yaml.safe_load(raw.decode('utf-8', 'replace').rstrip('\x00'))


class OutCoef(object):
    """python reader for outcoef files from exp (outcoef.py:41-133)"""

    def __init__(self, filename, verbose=1):
        self.coeffile = filename
        with open(filename, 'rb') as f:
            head = f.read(64)
        cmagic = int(np.frombuffer(head[:4], dtype='<u4')[0]) if len(head) >= 4 else 0
        if cmagic == SPH_MAGIC:
            self.basis = 'SphereSL'
            if verbose:
                print('OutCoef: reading SphereSL coefficients . . .')
            self.read_binary_sl_coefficients()
        elif cmagic == CYL_MAGIC:
            self.basis = 'Cylinder'
            if verbose:
                print('OutCoef: reading Cylinder coefficients . . .')
            self.read_binary_eof_coefficients()
        elif b'Sphere SL' in head:
            self.basis = 'SphereSL'
            if verbose:
                print('OutCoef: reading OLD SphereSL coefficients . . .')
                print('CAUION. These coefficients have a different normalisation scheme.')
            self.read_binary_sl_coefficients_old()
        else:
            self.basis = 'Cylinder'
            if verbose:
                print('OutCoef: reading OLD Cylinder coefficients . . .')
            self.read_binary_eof_coefficients_old()

    # -- cylinder -----------------------------------------------------------
    @staticmethod
    def _cyl_rows_to_array(rows, mmax, nmax):
        """(2*mmax+1, nmax) rows [cos0, cos1, sin1, cos2, sin2, ...] -> [2, mmax+1, nmax]"""
        out = np.zeros((2, mmax + 1, nmax))
        out[0, 0] = rows[0]
        if mmax > 0:
            out[0, 1:] = rows[1::2]
            out[1, 1:] = rows[2::2]
        return out

    def read_binary_eof_coefficients(self):
        '''outcoef.py:339-411 -> self.T, self.coefs[t, cos/sin, m, n]'''
        buf = np.fromfile(self.coeffile, dtype=np.uint8)
        pos, times, recs = 0, [], []
        while pos + 8 <= buf.size:
            magic, ln = np.frombuffer(buf[pos:pos + 8].tobytes(), dtype='<u4')
            if int(magic) != CYL_MAGIC:
                raise ValueError('outcoef: bad cylinder record magic at byte %d' % pos)
            D = _yaml_load(buf[pos + 8:pos + 8 + int(ln)].tobytes())
            pos += 8 + int(ln)
            mmax, nmax = int(D['mmax']), int(D['nmax'])
            nrow = 2 * mmax + 1
            rows = np.frombuffer(buf[pos:pos + 8 * nrow * nmax].tobytes(), dtype='<f8').reshape(nrow, nmax)
            pos += 8 * nrow * nmax
            times.append(float(D['time']))
            recs.append(self._cyl_rows_to_array(rows, mmax, nmax))
        self.T = np.array(times)
        self.coefs = np.array(recs) if recs else np.zeros((0, 2, 0, 0))

    def read_binary_eof_coefficients_old(self):
        '''outcoef.py:135-193'''
        size = os.path.getsize(self.coeffile)
        with open(self.coeffile, 'rb') as f:
            f.read(8)
            mmax, nmax = [int(v) for v in np.frombuffer(f.read(8), dtype='<u4')]
        nrow = 2 * mmax + 1
        reclen = 8 * nrow * nmax + 4 * 2 + 8
        n_outputs = int(size / reclen)
        rec = np.dtype([('time', '<f8'), ('mn', '<u4', (2,)), ('rows', '<f8', (nrow, nmax))])
        a = np.fromfile(self.coeffile, dtype=rec, count=n_outputs)
        self.T = a['time'].astype(np.float64)
        self.coefs = np.array([self._cyl_rows_to_array(r, mmax, nmax) for r in a['rows']]) if n_outputs else \
            np.zeros((0, 2, mmax + 1, nmax))

    # -- sphere -------------------------------------------------------------
    def read_binary_sl_coefficients(self):
        '''outcoef.py:264-336 -> self.T, self.coefs[t, (lmax+1)^2, n]'''
        buf = np.fromfile(self.coeffile, dtype=np.uint8)
        pos, times, recs = 0, [], []
        while pos + 8 <= buf.size:
            magic, ln = np.frombuffer(buf[pos:pos + 8].tobytes(), dtype='<u4')
            if int(magic) != SPH_MAGIC:
                raise ValueError('outcoef: bad sphere record magic at byte %d' % pos)
            D = _yaml_load(buf[pos + 8:pos + 8 + int(ln)].tobytes())
            pos += 8 + int(ln)
            lmax, nmax = int(D['lmax']), int(D['nmax'])
            nl = lmax * (lmax + 2) + 1
            rows = np.frombuffer(buf[pos:pos + 8 * nl * nmax].tobytes(), dtype='<f8').reshape(nmax, nl)
            pos += 8 * nl * nmax
            times.append(float(D['time']))
            recs.append(rows.T.copy())
        self.T = np.array(times)
        self.coefs = np.array(recs) if recs else np.zeros((0, 0, 0))

    def read_binary_sl_coefficients_old(self):
        '''outcoef.py:196-260'''
        size = os.path.getsize(self.coeffile)
        with open(self.coeffile, 'rb') as f:
            f.read(64 + 16)
            nmax, lmax = [int(v) for v in np.frombuffer(f.read(8), dtype='<u4')]
        nl = lmax * (lmax + 2) + 1
        reclen = 8 * nl * nmax + 4 * 2 + 8 * 2 + 64
        n_outputs = int(size / reclen)
        rec = np.dtype([('id', 'S64'), ('ts', '<f8', (2,)), ('nl', '<u4', (2,)), ('rows', '<f8', (nmax, nl))])
        a = np.fromfile(self.coeffile, dtype=rec, count=n_outputs)
        self.T = a['ts'][:, 0].astype(np.float64)
        self.coefs = np.transpose(a['rows'], (0, 2, 1)).copy()

    # -- repackaging (outcoef.py:413-520) ------------------------------------
    def _repackage_cylindrical_coefficients(self):
        numt, _1, numm, numn = self.coefs.shape
        self.C = dict()
        for m in range(0, numm):
            self.C[m] = dict()
            for ic, c in enumerate(['cos', 'sin']):
                self.C[m][c] = dict()
                for n in range(0, numn):
                    self.C[m][c][n] = self.coefs[:, ic, m, n]

    def _repackage_cylindrical_coefficients_compatibility(self):
        numt, _1, numm, numn = self.coefs.shape
        EOF_Dict = dict()
        for tt in range(0, numt):
            o = EOF_Object()
            o.time = self.T[tt]
            o.mmax = numm - 1
            o.nmax = numn
            o.filename = '[redacted]'
            o.comp = 'star'
            o.nbodies = 0.
            o.eof_file = '[redacted]'
            o.cos = self.coefs[tt, 0].copy()
            o.sin = self.coefs[tt, 1].copy()
            EOF_Dict[o.time] = o
        return EOF_Dict

    def _repackage_spherical_coefficients(self):
        numt, numl2, numn = self.coefs.shape
        numl = int(np.sqrt(numl2))
        self.C = dict()
        for l in range(0, numl):
            self.C[l] = dict()
            for m in range(0, l + 1):
                self.C[l][m] = dict()
                names = ['cos', 'sin'] if m > 0 else ['cos']
                for ip, pname in enumerate(names):
                    self.C[l][m][pname] = dict()
                    k = l * l + (0 if m == 0 else 2 * m - 1 + ip)
                    for n in range(0, numn):
                        self.C[l][m][pname][n] = self.coefs[:, k, n]


# ---------------------------------------------------------------------------
# writers: a GPU-accumulated series in EXP's own wire format
# ---------------------------------------------------------------------------
def _yaml_header(fields):
    return '\n'.join('%s: %s' % (k, repr(float(v)) if isinstance(v, float) else v) for k, v in fields).encode()


def write_cylinder_outcoef(filename, times, cos, sin, append=False):
    """times [S]; cos, sin [S, mmax+1, nmax] (eof.accumulate layout) -> YAML-style `outcoef` cylinder records."""
    cos = np.asarray(cos, dtype=np.float64); sin = np.asarray(sin, dtype=np.float64)
    S, M, N = cos.shape
    with open(filename, 'ab' if append else 'wb') as f:
        for t in range(S):
            hdr = _yaml_header([('time', float(times[t])), ('mmax', M - 1), ('nmax', N)])
            np.array([CYL_MAGIC, len(hdr)], dtype='<u4').tofile(f)
            f.write(hdr)
            for m in range(M):
                cos[t, m].astype('<f8').tofile(f)
                if m > 0:
                    sin[t, m].astype('<f8').tofile(f)


def write_sphere_outcoef(filename, times, expcoef, append=False):
    """times [S]; expcoef [S, (lmax+1)^2, nmax] (spheresl layout) -> YAML-style `outcoef` sphere records."""
    expcoef = np.asarray(expcoef, dtype=np.float64)
    S, K, N = expcoef.shape
    lmax = int(round(np.sqrt(K))) - 1
    if (lmax + 1) ** 2 != K:
        raise ValueError('expcoef rows are not (lmax+1)^2')
    with open(filename, 'ab' if append else 'wb') as f:
        for t in range(S):
            hdr = _yaml_header([('time', float(times[t])), ('lmax', lmax), ('nmax', N)])
            np.array([SPH_MAGIC, len(hdr)], dtype='<u4').tofile(f)
            f.write(hdr)
            for n in range(N):
                expcoef[t, :, n].astype('<f8').tofile(f)
