"""
psp_io.py -- reader (and writer, for synthetic snapshots) of EXP's PSP "OUT." phase-space files with YAML
component headers (exptool/io/psp_io.py:36-230).

Layout (little-endian): f8 time, u4 nbodies_tot, u4 ncomp; per component: [u4 magic 2915019716 + u4 pad when the
particle fields are float32], u4 nbodies, u4 nint_attr, u4 nfloat_attr, u4 infostringlen, the YAML info string
(`name`, `parameters: {indexing: ...}`, ...), then nbodies records of [i8 id if indexing] m x y z vx vy vz potE
[i4 int attrs] [float attrs].  Host-side NumPy, as in the reference; one structured bulk read per component.
"""
import numpy as np

FLOAT_MAGIC = 2915019716


class Input:
    """psp_io.Input (psp_io.py:36-100): .header, .time, .filename, .comp, .data{'x','y','z','vx','vy','vz','m','potE'[,'id']}"""

    def __init__(self, filename, comp=None, verbose=0):
        self.verbose = verbose
        self.filename = filename
        try:
            self.f = open(self.filename, 'rb')
        except Exception:
            raise IOError('Failed to open "{}"'.format(filename))
        self.primary_header = dict()
        self.comp_map = dict()
        self.header = dict()
        self._read_primary_header()
        self.comp = comp
        if comp is not None:
            if comp not in self.header:
                self.f.close()
                raise IOError('The specified component does not exist.')
            self.data = self._read_component_data(self.filename, self.header[comp]['nbodies'],
                                                  int(self.header[comp]['data_start']))
        self.f.close()

    def _check_magic_number(self):
        self.f.seek(16)
        cmagic = np.fromfile(self.f, dtype=np.uint32, count=1)
        if cmagic.size and cmagic[0] == FLOAT_MAGIC:
            self._float_len, self._float_str = 4, 'f'
        else:
            self._float_len, self._float_str = 8, 'd'

    def _read_primary_header(self):
        self._check_magic_number()
        self.f.seek(0)
        self.time, = np.fromfile(self.f, dtype='<f8', count=1)
        self._nbodies_tot, self._ncomp = np.fromfile(self.f, dtype=np.uint32, count=2)
        data_start = 16
        for _ in range(0, self._ncomp):
            self.f.seek(data_start)
            data_start = self._read_out_component_header()

    def _read_out_component_header(self):
        import yaml
        if self._float_len == 4:
            _1, _2, nbodies, nint_attr, nfloat_attr, infostringlen = np.fromfile(self.f, dtype=np.uint32, count=6)
        else:
            nbodies, nint_attr, nfloat_attr, infostringlen = np.fromfile(self.f, dtype=np.uint32, count=4)
        head = self.f.read(int(infostringlen))
        head_dict = yaml.safe_load(head.decode().rstrip('\x00'))
        comp_data_pos = self.f.tell()
        nfields = 8
        comp_length = int(nbodies) * (8 * int(head_dict['parameters']['indexing']) + self._float_len * nfields +
                                      4 * int(nint_attr) + self._float_len * int(nfloat_attr))
        head_dict['nint_attr'] = int(nint_attr)
        head_dict['nfloat_attr'] = int(nfloat_attr)
        head_dict['nbodies'] = int(nbodies)
        head_dict['data_start'] = comp_data_pos
        head_dict['data_end'] = comp_data_pos + comp_length
        self.header[head_dict['name']] = head_dict
        try:
            self.indexing = head_dict['parameters']['indexing']
        except Exception:
            self.indexing = head_dict['indexing'] == 'true'
        return comp_data_pos + comp_length

    def _component_dtype(self, comp):
        h = self.header[comp]
        fl = '<f4' if self._float_len == 4 else '<f8'
        fields, names = [], []
        if h['parameters']['indexing']:
            fields.append('<i8'); names.append('id')
        fields += [fl] * 8; names += ['m', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'potE']
        fields += ['<i4'] * h['nint_attr']; names += ['i_attr{}'.format(i) for i in range(h['nint_attr'])]
        fields += [fl] * h['nfloat_attr']; names += ['f_attr{}'.format(i) for i in range(h['nfloat_attr'])]
        return np.dtype({'names': names, 'formats': fields}), names

    def _read_component_data(self, filename, nbodies, offset):
        dtype, names = self._component_dtype(self.comp)
        out = np.memmap(filename, dtype=dtype, shape=(int(nbodies),), offset=offset, mode='r')
        tbl = dict()
        for name in names:
            tbl[name] = np.array(out[name], copy=True)
        del out
        return tbl


def write_psp(filename, time, components, float32=False):
    """
    Write a PSP "OUT." file.  components: list of dicts {name, data{'m','x','y','z','vx','vy','vz','potE'[,'id']},
    indexing (bool, default False), extra (dict of further YAML header entries)}.  For synthetic snapshots
    (tests, benchmarks); the reference reader reads these files back unchanged.
    """
    import yaml
    fl = '<f4' if float32 else '<f8'
    ntot = sum(len(c['data']['m']) for c in components)
    with open(filename, 'wb') as f:
        np.array([time], dtype='<f8').tofile(f)
        np.array([ntot, len(components)], dtype='<u4').tofile(f)
        for c in components:
            indexing = bool(c.get('indexing', False))
            hd = {'name': c['name'], 'parameters': {'indexing': indexing}}
            hd.update(c.get('extra', {}))
            info = yaml.safe_dump(hd).encode()
            n = len(c['data']['m'])
            if float32:
                np.array([FLOAT_MAGIC, 0], dtype='<u4').tofile(f)
            np.array([n, 0, 0, len(info)], dtype='<u4').tofile(f)
            f.write(info)
            names = (['id'] if indexing else []) + ['m', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'potE']
            rec = np.zeros(n, dtype=np.dtype({'names': names, 'formats': (['<i8'] if indexing else []) + [fl] * 8}))
            for k in names:
                rec[k] = c['data'][k]
            rec.tofile(f)
    return filename
