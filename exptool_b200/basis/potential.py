"""
potential.py -- drop-in for exptool.basis.potential.Fields on the hot path
(potential.py:49-497): hold the disc (EOF) and halo (SL) tables and coefficients
of one snapshot and return the combined force at points.

The device state is two table handles (ops.EOFTables / ops.SLTables) holding the
coefficient contraction for the current truncation (set_field_parameters).

total_coefficients reads the PSP snapshot (exptool_b200.io.psp_io), applies the
bar-frame rotation and the centring on the device and accumulates both bases
(SURVEY.md section 8f rank 3).  Not mirrored: rotation curves, resonance helpers,
EnergyKappa.
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
                 centering=False, mutual_center=False, verbose=1, table_fp32=False):
        # table_fp32 (extension; BASELINE north_star's "<= 1e-5 where FP32 table interpolation is used"): hold the
        # contracted tables of return_forces_* / orbit integration as float -- ~1.5x faster, forces ~6e-8 from FP64
        self.table_fp32 = bool(table_fp32)
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
        Fields.total_coefficients (potential.py:98-252): read the 'star' and 'dark' components of the PSP file
        `self.filename` (or take the particle sets passed as disc= / halo=: `holder`, `.data` container or
        (x,y,z,m) tuple), optionally rotate both into the bar frame (transform, pattern.BarTransform) and
        recentre them on the mass-weighted centre of the 10^4 innermost disc particles (centering,
        mutual_center), then accumulate the EOF and SL coefficients.  The snapshot is uploaded once; the
        transforms, the centre search and the accumulation all run on the device-resident arrays.
        halofac = N_halo / N_halo = 1 in the reference (potential.py:143); pass halofac= to override.
        '''
        from ..io import particle
        torch = ops.torch
        if disc is None:
            disc = particle.Input(self.filename, comp='star', verbose=self.verbose)
        if halo is None:
            halo = particle.Input(self.filename, comp='dark', verbose=self.verbose)
        self.time = getattr(disc, 'time', self.time)
        xd, yd, zd, md = [ops.dev(a) for a in particle.particle_arrays(disc)]
        xh, yh, zh, mh = [ops.dev(a) for a in particle.particle_arrays(halo)]
        self.halofac = 1.0 if halofac is None else float(halofac)
        if self.transform:
            # pattern.py:104: bar angle from the disc (0 < R < 1), the same rotation for the halo (potential.py:147)
            self.bar_angle = -1. * ops.bar_fourier_angle(xd, yd, minr=0., maxr=1.)
            if self.verbose > 1:
                print('potential.Fields.total_coefficients: Using bar_angle {0:4.3f}'.format(self.bar_angle))
            xd, yd = ops.affine_xy(xd, yd, angle=self.bar_angle)
            xh, yh = ops.affine_xy(xh, yh, angle=self.bar_angle)
        if self.centering:
            print('potential.Fields.total_coefficients: Computing centering (centering=True)')
            ncenter = 10000
            self.xcen_disk, self.ycen_disk, self.zcen_disk = ops.inner_center_of_mass(xd, yd, zd, md, ncenter)
            if self.mutual_center:
                print('potential.Fields.total_coefficients: Using computed disk center for halo (mutual_center=True)')
                self.xcen_halo, self.ycen_halo, self.zcen_halo = self.xcen_disk, self.ycen_disk, self.zcen_disk
            else:
                # potential.py:190-200: the DISC ranking indexes the halo arrays (replicated)
                self.xcen_halo, self.ycen_halo, self.zcen_halo = ops.inner_center_of_mass(
                    xd, yd, zd, md, ncenter, values=(xh, yh, zh, mh))
            print('potential.Fields.total_coefficients: (x,y,z) = {0:6.5f},{1:6.5f},{2:6.5f}'
                  .format(float(self.xcen_disk), float(self.ycen_disk), float(self.zcen_disk)))
            xd, yd, zd = ops.affine_xy(xd, yd, zd, center=(self.xcen_disk, self.ycen_disk, self.zcen_disk))
            xh, yh, zh = ops.affine_xy(xh, yh, zh, center=(self.xcen_halo, self.ycen_halo, self.zcen_halo))
        else:
            self.xcen_disk = self.ycen_disk = self.zcen_disk = 0.
            self.xcen_halo = self.ycen_halo = self.zcen_halo = 0.
        D = particle.Particles(xd, yd, zd, md, time=self.time, filename=getattr(disc, 'filename', self.filename),
                               comp=getattr(disc, 'comp', 'star'))
        Hh = particle.Particles(xh, yh, zh, mh, time=self.time, filename=getattr(halo, 'filename', self.filename),
                                comp=getattr(halo, 'comp', 'dark'))
        self.EOF = eof.compute_coefficients(D, self.eof_file, verbose=self.verbose, no_odd=self.no_odd)
        self.SL = spheresl.compute_coefficients(Hh, self.sph_file, self.model_file, verbose=self.verbose,
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
        # table precision is a property of THIS instance's handles (no process-wide state involved)
        self._E.set_table_fp32(bool(getattr(self, 'table_fp32', False)))
        self._H.set_table_fp32(bool(getattr(self, 'table_fp32', False)))
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

    def precision(self):
        """kept for callers of round 1: the table precision now lives in this instance's device handles (_build_device)"""
        return ops.table_precision(getattr(self, 'table_fp32', False))

    # -- forces ---------------------------------------------------------------
    def _eval(self, fn, xval, yval, zval, rotpos):
        scalar = np.ndim(xval) == 0
        E, H = self.device_handles()
        with self.precision():
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


def make_fields(eof_file, sph_file, model_file, cos, sin, expcoef, halofac=1.0, verbose=0, table_fp32=False):
    """
    Build a Fields object from files and an existing coefficient set (what the
    reference does by hand after restore_*_coefficients: set .EOF/.SL/.halofac,
    then prep_tables).
    """
    F = Fields('memory', eof_file, sph_file, model_file, verbose=verbose, table_fp32=table_fp32)
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


def get_fields(simulation_directory, simulation_name, intime, eof_file, sph_file, model_file, bar_file='', nhalo=1000000,
               transform=True, fileprefix='OUT'):
    '''potential.get_fields (potential.py:1243-1288): snapshot -> (Fields with coefficients and tables, pattern speed,
    rotation frequency).  With transform the pattern speed comes from the printed bar file (pattern.BarDetermine).'''
    from ..analysis import pattern
    from ..io import particle
    infile = simulation_directory + fileprefix + '.' + simulation_name + '.%05i' % intime
    BarInstance = pattern.BarDetermine()
    if transform:
        BarInstance.read_bar(bar_file)
        BarInstance.frequency_and_derivative(spline_derivative=2)
        PSPDump = particle.Input(infile, comp='star')
        patt = pattern.find_barpattern(PSPDump.time, BarInstance, smth_order=None)
        rotfreq = patt / (2. * np.pi)
    else:
        patt = 0.
        rotfreq = 0.
    F = Fields(infile, eof_file, sph_file, model_file, nhalo=nhalo, transform=transform, no_odd=False, centering=True,
               mutual_center=True)
    F.total_coefficients()
    F.prep_tables()
    return F, patt, rotfreq
