"""
spl_io.py -- reader (and writer, for synthetic snapshots) of EXP's split phase-space "SPL." files
(exptool/io/spl_io.py:25-270).

A master file  SPL.<runtag>.<filenum>  holds the primary header (f8 time, u4 nbodies_tot, u4 ncomp) and, per
component, seven u4 (magic 2915019716, pad, nprocs, nbodies, nint_attr, nfloat_attr, infostringlen), the YAML info
string and nprocs names of 1024 bytes each (NUL padded).  Each named subfile, in the master's directory, is a u4 body
count followed by that many PSP records ([i8 id] m x y z vx vy vz potE [i4 attrs] [float attrs]).  As in the
reference, the seven-word component header is read whatever the precision: only files that carry the float32 magic
word are laid out that way.  Host-side NumPy; one structured bulk read per subfile.
"""
import os

import numpy as np

from . import psp_io

FLOAT_MAGIC = psp_io.FLOAT_MAGIC
PBUF_SZ = 1024


class Input(psp_io.Input):
    """spl_io.Input (spl_io.py:25-118): .header, .time, .filename, .comp, .subfiles, .data{...} concatenated over subfiles"""

    def __init__(self, filename, comp=None, verbose=0):
        self.verbose = verbose
        self.filename = filename
        try:
            self.f = open(self.filename, 'rb')
        except Exception:
            raise IOError('Failed to open "{}"'.format(filename))
        self.primary_header = dict()
        self.component_map = dict()
        self.header = dict()
        self._nprocs = dict()
        self._read_primary_header()
        self.comp = comp
        if comp is None:
            self._summarise_primary_header()
            self.f.close()
            return
        if comp not in self.header:
            self.f.close()
            raise IOError('The specified component does not exist.')
        self.indir = filename.split('SPL')[0]
        self._make_spl_file_list(comp)
        self.f.close()
        parts = []
        for name in self.subfiles:
            path = self.indir + name
            with open(path, 'rb') as sf:
                nbodies, = np.fromfile(sf, dtype=np.uint32, count=1)
            parts.append(self._read_component_data(path, int(nbodies), 4))
        self.data = {k: np.concatenate([p[k] for p in parts]) for k in parts[0].keys()}

    def _summarise_primary_header(self):
        comps = list(self.header.keys())
        print("Found {} components.".format(len(comps)))
        for n, c in enumerate(comps):
            print("Component {}: {}".format(n, c))

    def _read_out_component_header(self):
        import yaml
        data_start = self.f.tell()
        _1, _2, nprocs, nbodies, nint_attr, nfloat_attr, infostringlen = np.fromfile(self.f, dtype=np.uint32, count=7)
        head = self.f.read(int(infostringlen))
        head_dict = yaml.safe_load(head.decode().rstrip('\x00'))
        head_dict['nint_attr'] = int(nint_attr)
        head_dict['nfloat_attr'] = int(nfloat_attr)
        head_dict['nbodies'] = int(nbodies)
        names_at = 4 * 7 + int(infostringlen) + data_start
        self.component_map[head_dict['name']] = names_at
        self._nprocs[head_dict['name']] = int(nprocs)
        self.nprocs = int(nprocs)
        self.header[head_dict['name']] = head_dict
        try:
            self.indexing = head_dict['parameters']['indexing']
        except Exception:
            self.indexing = head_dict['indexing'] == 'true'
        return names_at + int(nprocs) * PBUF_SZ

    def _make_spl_file_list(self, comp):
        self.f.seek(self.component_map[comp])
        self.subfiles = []
        for _ in range(self._nprocs[comp]):
            buf = self.f.read(PBUF_SZ)
            self.subfiles.append(buf.split(b'\x00')[0].decode())


def write_spl(filename, time, components, nprocs=2, float32=True):
    """
    Write an SPL master file plus its subfiles (synthetic snapshots for tests / benchmarks).  `filename` must contain
    'SPL' (e.g. <dir>/SPL.run.00001); component c, piece p goes to <filename>_<c>-<p>.  components as in
    psp_io.write_psp.  The particles of a component are split into `nprocs` consecutive blocks.
    """
    import yaml
    fl = '<f4' if float32 else '<f8'
    base = os.path.basename(filename)
    indir = filename.split('SPL')[0]
    ntot = sum(len(c['data']['m']) for c in components)
    with open(filename, 'wb') as f:
        np.array([time], dtype='<f8').tofile(f)
        np.array([ntot, len(components)], dtype='<u4').tofile(f)
        for ci, c in enumerate(components):
            indexing = bool(c.get('indexing', False))
            hd = {'name': c['name'], 'parameters': {'indexing': indexing}}
            hd.update(c.get('extra', {}))
            info = yaml.safe_dump(hd).encode()
            n = len(c['data']['m'])
            np.array([FLOAT_MAGIC if float32 else 0, 0, nprocs, n, 0, 0, len(info)], dtype='<u4').tofile(f)
            f.write(info)
            names = (['id'] if indexing else []) + ['m', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'potE']
            dtype = np.dtype({'names': names, 'formats': (['<i8'] if indexing else []) + [fl] * 8})
            bounds = np.linspace(0, n, nprocs + 1).astype(int)
            for p in range(nprocs):
                sub = '{}_{}-{}'.format(base, ci, p)
                f.write(sub.encode().ljust(PBUF_SZ, b'\x00'))
                lo, hi = bounds[p], bounds[p + 1]
                rec = np.zeros(hi - lo, dtype=dtype)
                for k in names:
                    rec[k] = np.asarray(c['data'][k])[lo:hi]
                with open(os.path.join(indir, sub), 'wb') as sf:
                    np.array([hi - lo], dtype='<u4').tofile(sf)
                    rec.tofile(sf)
    return filename
