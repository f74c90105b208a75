"""
potential.py -- drop-in for exptool.basis.potential.Fields on the hot path
(potential.py:49-497): hold the disc (EOF) and halo (SL) tables and coefficients
of one snapshot and return the combined force at points.

The device state is two table handles (ops.EOFTables / ops.SLTables) holding the
coefficient contraction for the current truncation (set_field_parameters).

Not mirrored: PSP snapshot ingest inside total_coefficients (pass in-memory
particle sets instead), centering / BarTransform, rotation curves, resonance
helpers, EnergyKappa.
"""
import numpy as np

from . import eof
from . import spheresl
from ..utils import halo_methods
from .. import ops


class Fields():
    '''
    class to accumulate particles from a snapshot and return the field quantities
    (exptool.basis.potential.Fields, potential.py:49)
    '''

    def __init__(self, infile, eof_file, sph_file, model_file, nhalo=1000000, transform=False, no_odd=False,
                 centering=False, mutual_center=False, verbose=1):
        self.filename = infile
        self.eof_file = eof_file
        self.sph_file = sph_file
        self.model_file = model_file
        self.nhalo = nhalo
        self.transform = transform
        self.no_odd = no_odd
        self.centering = centering
        self.mutual_center = mutual_center
        self.verbose = verbose
        self.halofac = 1.0
        self.time = 0.0
        self._E = None
        self._H = None
        self._contract_key = None

    # -- coefficients ---------------------------------------------------------
    def total_coefficients(self, disc=None, halo=None, halofac=None):
        '''
        Fields.total_coefficients (potential.py:98-252).  The reference reads the disc and
        halo components from a PSP file; snapshot ingest is outside the path, so the
        particle sets are passed in (`holder`, `.data` container or (x,y,z,m) tuple).
        halofac = N_total_halo / N_used_halo (potential.py:143); 1 if all halo particles are given.
        '''
        if disc is None or halo is None:
            raise NotImplementedError('Fields.total_coefficients: PSP snapshot ingest is outside the B200 hot path; '
                                      'pass disc= and halo= particle sets')
        if self.transform or self.centering:
            raise NotImplementedError('Fields.total_coefficients: transform/centering are outside the B200 hot path')
        self.xcen_disk = self.ycen_disk = self.zcen_disk = 0.
        self.xcen_halo = self.ycen_halo = self.zcen_halo = 0.
        if halofac is not None:
            self.halofac = float(halofac)
        self.EOF = eof.compute_coefficients(disc, self.eof_file, verbose=self.verbose, no_odd=self.no_odd)
        self.SL = spheresl.compute_coefficients(halo, self.sph_file, self.model_file, verbose=self.verbose,
                                                no_odd=self.no_odd)
        self._contract_key = None

    # -- tables ---------------------------------------------------------------
    def prep_tables(self):
        '''Fields.prep_tables (potential.py:257-294): read the cache files, keep host copies
        under the reference's attribute names, and build the device handles.'''
        try:
            x = self.EOF.eof_file
        except AttributeError:
            print('potential.Fields.prep_tables: must first call total_coefficients.')
            raise
        self.potC, self.rforceC, self.zforceC, self.densC, \
            self.potS, self.rforceS, self.zforceS, self.densS = eof.parse_eof(self.EOF.eof_file)
        self.rmindisk, self.rmaxdisk, self.numx, self.numy, self.mmax, self.norder, self.ascale, self.hscale, \
            self.cmapdisk, self.densdisk = eof.eof_params(self.EOF.eof_file)
        self.XMIN, self.XMAX, self.dX, self.YMIN, self.YMAX, self.dY = eof.set_table_params(
            RMAX=self.rmaxdisk, RMIN=self.rmindisk, ASCALE=self.ascale, HSCALE=self.hscale,
            NUMX=self.numx, NUMY=self.numy, CMAP=self.cmapdisk)
        self.disk_use_m = self.mmax
        self.disk_use_n = self.norder
        self.lmaxhalo, self.nmaxhalo, self.numrhalo, self.cmaphalo, self.rminhalo, self.rmaxhalo, self.scalehalo, \
            self.ltablehalo, self.evtablehalo, self.eftablehalo = halo_methods.read_cached_table(self.SL.sph_file)
        self.xihalo, self.rarrhalo, self.p0halo, self.d0halo = halo_methods.init_table(
            self.SL.model_file, self.numrhalo, self.rminhalo, self.rmaxhalo, cmap=self.cmaphalo, scale=self.scalehalo)
        self.halo_use_l = self.lmaxhalo
        self.halo_use_n = self.nmaxhalo
        self._E = self._H = None           # device handles are (re)built on first use
        self._contract_key = None

    def _build_device(self):
        self._E = ops.EOFTables(self.potC, self.potS, self.mmax, self.norder, self.XMIN, self.dX, self.YMIN, self.dY,
                                self.numx, self.numy, self.ascale, self.hscale, self.cmapdisk,
                                rforceC=self.rforceC, zforceC=self.zforceC, rforceS=self.rforceS, zforceS=self.zforceS)
        self._H = ops.SLTables(self.lmaxhalo, self.nmaxhalo, self.numrhalo, self.cmaphalo, self.scalehalo,
                               self.evtablehalo, self.eftablehalo, self.xihalo, self.p0halo, self.d0halo)
        self._contract_key = None

    def set_field_parameters(self, no_odd=False, halo_l=-1, halo_n=-1, disk_m=-1, disk_n=-1):
        '''Fields.set_field_parameters (potential.py:365-377)'''
        self.no_odd = no_odd
        if halo_l > -1: self.halo_use_l = halo_l
        if halo_n > -1: self.halo_use_n = halo_n
        if disk_m > -1: self.disk_use_m = disk_m
        if disk_n > -1: self.disk_use_n = disk_n

    def reset_field_parameters(self):
        '''Fields.reset_field_parameters (potential.py:380-386)'''
        self.no_odd = False
        self.halo_use_l = self.lmaxhalo
        self.halo_use_n = self.nmaxhalo
        self.disk_use_m = self.mmax
        self.disk_use_n = self.norder

    # -- device state ---------------------------------------------------------
    def device_handles(self):
        """(ops.EOFTables, ops.SLTables) holding the contraction for the current parameters."""
        if self._E is None or self._H is None:
            if not hasattr(self, 'potC') or not hasattr(self, 'xihalo'):
                raise RuntimeError('potential.Fields: must first call total_coefficients and prep_tables.')
            self._build_device()
        try:
            x = self.no_odd
        except AttributeError:
            self.set_field_parameters()
        cosc = np.asarray(self.EOF.cos, dtype=np.float64)
        sinc = np.asarray(self.EOF.sin, dtype=np.float64)
        coef = np.asarray(self.SL.expcoef, dtype=np.float64)
        key = (int(self.disk_use_m), int(self.disk_use_n), int(self.halo_use_l), int(self.halo_use_n),
               bool(self.no_odd), float(self.halofac), eof._fingerprint(cosc), eof._fingerprint(sinc),
               eof._fingerprint(coef))
        if key != self._contract_key:
            # eof.force_eval slices to MMAX, NMAX (eof.py:784-793); spheresl.force_eval to lmax, nmax (1138-1148)
            self._E.contract(cosc, sinc, m1=0, m2=self.disk_use_m, nuse=self.disk_use_n, no_odd=self.no_odd)
            self._H.contract(self.halofac * coef, l1=0, l2=self.halo_use_l, nuse=self.halo_use_n, no_odd=self.no_odd)
            self._contract_key = key
        return self._E, self._H

    # -- forces ---------------------------------------------------------------
    def _eval(self, fn, xval, yval, zval, rotpos):
        scalar = np.ndim(xval) == 0
        E, H = self.device_handles()
        out = ops.to_host(fn(E, H, np.atleast_1d(np.asarray(xval, dtype=np.float64)),
                             np.atleast_1d(np.asarray(yval, dtype=np.float64)),
                             np.atleast_1d(np.asarray(zval, dtype=np.float64)), rotpos=float(rotpos)))
        if scalar:
            return tuple(np.float64(v[0]) for v in out)
        return tuple(out)

    def return_forces_cart(self, xval, yval, zval, rotpos=0.0):
        '''
        Fields.return_forces_cart (potential.py:445-497):
        fxdisk, fxhalo, fydisk, fyhalo, fzdisk, fzhalo, diskp, halop+halop0.
        Scalars (as the reference is called) or equal-length arrays.
        '''
        return self._eval(ops.field_force_cart, xval, yval, zval, rotpos)

    def return_forces_cyl(self, xval, yval, zval, rotpos=0.0):
        '''
        Fields.return_forces_cyl (potential.py:389-440):
        diskfr, frhalo, diskfp, -halofp, diskfz, fzhalo, -diskp, halop+halop0.
        '''
        return self._eval(ops.field_force_cyl, xval, yval, zval, rotpos)


    # -- frozen-field file (potential.py:738-958), byte-compatible ---------------
    def save_field(self, filename=''):
        '''
        Fields.save_field (potential.py:738-958): everything needed to evaluate the field, in the
        reference's byte layout.  Geometry scalars are stored as float32 exactly
        as the reference does, so a restored field differs from the original at ~1e-6 relative.
        '''
        if filename == '':
            print('potential.Fields.save_field: No filename specified.')
        with open(filename, 'wb') as f:
            for sname in (self.filename, self.eof_file, self.sph_file, self.model_file):
                np.array([sname], dtype='S100').tofile(f)
            for v in (self.nhalo, self.transform, self.no_odd, self.centering, self.mutual_center, self.verbose):
                np.array([v], dtype='i4').tofile(f)
            np.array([self.time], dtype='f4').tofile(f)
            for v in (self.numx, self.numy, self.mmax, self.norder, self.cmapdisk, self.densdisk):
                np.array([v], dtype='i4').tofile(f)
            for v in (self.rmindisk, self.rmaxdisk, self.ascale, self.hscale, self.XMIN, self.dX, self.YMIN, self.dY,
                      getattr(self, 'xcen_disk', 0.), getattr(self, 'ycen_disk', 0.), getattr(self, 'zcen_disk', 0.)):
                np.array([v], dtype='f4').tofile(f)
            np.array(np.asarray(self.EOF.cos).reshape(-1, ), dtype='f8').tofile(f)
            np.array(np.asarray(self.EOF.sin).reshape(-1, ), dtype='f8').tofile(f)
            for t in (self.potC, self.rforceC, self.zforceC, self.densC, self.potS, self.rforceS, self.zforceS, self.densS):
                np.array(t.reshape(-1, ), dtype='f8').tofile(f)
            for v in (self.halofac, self.rminhalo, self.rmaxhalo, self.scalehalo, getattr(self, 'xcen_halo', 0.),
                      getattr(self, 'ycen_halo', 0.), getattr(self, 'zcen_halo', 0.)):
                np.array([v], dtype='f4').tofile(f)
            for v in (self.numrhalo, self.cmaphalo, self.lmaxhalo, self.nmaxhalo):
                np.array([v], dtype='i4').tofile(f)
            ltable = getattr(self, 'ltablehalo', getattr(self, 'ltable', np.arange(self.lmaxhalo + 1.)))
            for t in (self.xihalo, self.p0halo, self.d0halo, ltable, self.evtablehalo, self.eftablehalo,
                      np.asarray(self.SL.expcoef, dtype=np.float64)):
                np.array(np.asarray(t).reshape(-1, ), dtype='f8').tofile(f)


def restore_field(filename=''):
    '''
    potential.restore_field (potential.py:968-1040): rebuild a Fields instance from a frozen-field file.
    Attribute names and types follow the reference (strings come back as bytes, scalars as NumPy types).
    '''
    with open(filename, 'rb') as f:
        [infile, eof_file, sph_file, model_file] = np.fromfile(f, dtype='S100', count=4)
        [nhalo, transform, no_odd, centering, mutual_center, verbose] = np.fromfile(f, dtype='i4', count=6)
        [time] = np.fromfile(f, dtype='f4', count=1)
        F = Fields(infile, eof_file, sph_file, model_file, nhalo=nhalo, transform=transform, no_odd=no_odd,
                   centering=centering, mutual_center=mutual_center, verbose=verbose)
        F.time = time
        [F.numx, F.numy, F.mmax, F.norder, F.cmapdisk, F.densdisk] = np.fromfile(f, dtype='i4', count=6)
        [F.rmindisk, F.rmaxdisk, F.ascale, F.hscale, F.XMIN, F.dX, F.YMIN, F.dY, F.xcen_disk, F.ycen_disk,
         F.zcen_disk] = np.fromfile(f, dtype='f4', count=11)
        F.EOF = eof.EOF_Object()
        nc = (F.mmax + 1) * F.norder
        F.EOF.cos = np.fromfile(f, dtype='f8', count=nc).reshape([(F.mmax + 1), F.norder])
        F.EOF.sin = np.fromfile(f, dtype='f8', count=nc).reshape([(F.mmax + 1), F.norder])
        shape = [(F.mmax + 1), F.norder, (F.numx + 1), (F.numy + 1)]
        nt = int(np.prod(shape))
        for name in ('potC', 'rforceC', 'zforceC', 'densC', 'potS', 'rforceS', 'zforceS', 'densS'):
            setattr(F, name, np.fromfile(f, dtype='f8', count=nt).reshape(shape))
        [F.halofac, F.rminhalo, F.rmaxhalo, F.scalehalo, F.xcen_halo, F.ycen_halo, F.zcen_halo] = \
            np.fromfile(f, dtype='f4', count=7)
        [F.numrhalo, F.cmaphalo, F.lmaxhalo, F.nmaxhalo] = np.fromfile(f, dtype='i4', count=4)
        F.xihalo = np.fromfile(f, dtype='f8', count=F.numrhalo)
        F.p0halo = np.fromfile(f, dtype='f8', count=F.numrhalo)
        F.d0halo = np.fromfile(f, dtype='f8', count=F.numrhalo)
        F.ltable = np.fromfile(f, dtype='f8', count=(F.lmaxhalo + 1))
        F.ltablehalo = F.ltable
        F.evtablehalo = np.fromfile(f, dtype='f8', count=(F.lmaxhalo + 1) * F.nmaxhalo).reshape(
            [(F.lmaxhalo + 1), F.nmaxhalo])
        F.eftablehalo = np.fromfile(f, dtype='f8', count=(F.lmaxhalo + 1) * F.nmaxhalo * F.numrhalo).reshape(
            [(F.lmaxhalo + 1), F.nmaxhalo, F.numrhalo])
        F.SL = spheresl.SL_Object()
        F.SL.expcoef = np.fromfile(f, dtype='f8', count=(F.lmaxhalo + 1) * (F.lmaxhalo + 1) * F.nmaxhalo).reshape(
            [(F.lmaxhalo + 1) * (F.lmaxhalo + 1), F.nmaxhalo])
    F.disk_use_m = F.mmax
    F.disk_use_n = F.norder
    F.halo_use_l = F.lmaxhalo
    F.halo_use_n = F.nmaxhalo
    F.SL.model_file = F.model_file
    F.SL.sph_file = F.sph_file
    F.EOF.eof_file = F.eof_file
    F.EOF.mmax = F.mmax
    return F


def make_fields(eof_file, sph_file, model_file, cos, sin, expcoef, halofac=1.0, verbose=0):
    """
    Build a Fields object from files and an existing coefficient set (what the
    reference does by hand after restore_*_coefficients: set .EOF/.SL/.halofac,
    then prep_tables).
    """
    F = Fields('memory', eof_file, sph_file, model_file, verbose=verbose)
    F.EOF = eof.EOF_Object()
    F.EOF.eof_file = eof_file
    F.EOF.cos = np.asarray(cos, dtype=np.float64)
    F.EOF.sin = np.asarray(sin, dtype=np.float64)
    F.SL = spheresl.SL_Object()
    F.SL.sph_file = sph_file
    F.SL.model_file = model_file
    F.SL.expcoef = np.asarray(expcoef, dtype=np.float64)
    F.halofac = float(halofac)
    F.prep_tables()
    F.set_field_parameters()
    return F
