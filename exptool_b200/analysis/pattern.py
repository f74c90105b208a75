"""
pattern.py -- BarTransform (exptool/analysis/pattern.py:64-169): rotate a snapshot into the bar frame.

The m = 2 Fourier sums and the planar rotation run on the device (ops.bar_fourier_angle, ops.affine_xy ->
bfe_bar_fourier, bfe_affine_xy); the class keeps the reference's interface (`.bar_angle`, `.data` dict with
x, y, z, vx, vy, vz, m, potE, `.time`, `.filename`, `.comp`).  Of BarDetermine only what the workflow driver
potential.get_fields needs is mirrored (read_bar, frequency_and_derivative, find_barpattern: host-side post-processing
of a printed bar file); bar tracking over file lists and the plotting helpers are out of scope.
"""
import numpy as np
from scipy.interpolate import UnivariateSpline

from .. import ops


class BarTransform():
    '''BarTransform (pattern.py:64-153): planar transformation of the particles to the bar frame'''

    def __init__(self, ParticleInstanceIn, bar_angle=None, rel_bar_angle=0., minr=0., maxr=1.):
        self.ParticleInstanceIn = ParticleInstanceIn
        self.bar_angle = bar_angle
        self.data = dict()
        if self.bar_angle is None:
            # pattern.py:104: only maxr is forwarded to bar_fourier_compute (minr stays at its default 0)
            self.bar_angle = -1. * self.bar_fourier_compute(self.ParticleInstanceIn.data['x'],
                                                            self.ParticleInstanceIn.data['y'], maxr=maxr)
        self.bar_angle += rel_bar_angle
        self.calculate_transform_and_return()

    def calculate_transform_and_return(self):
        '''pattern.py:113-152'''
        d = self.ParticleInstanceIn.data
        tx, ty = ops.affine_xy(d['x'], d['y'], angle=self.bar_angle)
        tvx, tvy = ops.affine_xy(d['vx'], d['vy'], angle=self.bar_angle)
        self.data['x'] = ops.to_host(tx)
        self.data['y'] = ops.to_host(ty)
        self.data['z'] = np.copy(d['z'])
        self.data['vx'] = ops.to_host(tvx)
        self.data['vy'] = ops.to_host(tvy)
        self.data['vz'] = np.copy(d['vz'])
        self.data['m'] = d['m']
        self.data['potE'] = d['potE']
        self.time = self.ParticleInstanceIn.time
        self.filename = self.ParticleInstanceIn.filename
        self.comp = self.ParticleInstanceIn.comp

    def bar_fourier_compute(self, posx, posy, minr=0., maxr=1.):
        '''pattern.py:155-169: m = 2 phase angle of the particles with minr < R < maxr'''
        return ops.bar_fourier_angle(posx, posy, minr=minr, maxr=maxr)


class BarDetermine():
    '''BarDetermine (pattern.py:180-472), the parts potential.get_fields uses: a printed bar file (time, position
    [, derivative] per line) -> pattern speed.'''

    def __init__(self, **kwargs):
        if 'file' in kwargs:
            try:
                self.read_bar(kwargs['file'])
                print('pattern.BarDetermine: BarInstance sucessfully read.')
            except Exception:
                print('pattern.BarDetermine: no compatible bar file found.')

    def frequency_and_derivative(self, smth_order=None, fft_order=None, spline_derivative=None, verbose=0):
        '''pattern.py:362-401: backward differences, optionally replaced by a polynomial fit of the derivative or by the
        derivative of a smoothing cubic spline of the position'''
        if (smth_order or fft_order) and verbose:
            print('Cannot assure proper functionality of both order smoothing and low pass filtering.')
        self.deriv = np.zeros_like(self.pos)
        for i in range(1, len(self.pos)):
            self.deriv[i] = (self.pos[i] - self.pos[i - 1]) / (self.time[i] - self.time[i - 1])
        if smth_order:
            self.deriv = np.poly1d(np.polyfit(self.time, self.deriv, smth_order))(self.time)
        if spline_derivative:
            spl = UnivariateSpline(self.time, self.pos, k=3, s=spline_derivative)
            self.deriv = (spl.derivative())(self.time)
            self.dderiv = np.zeros_like(self.deriv)
            for indx, timeval in enumerate(self.time):
                self.dderiv[indx] = spl.derivatives(timeval)[2]

    def read_bar(self, infile):
        '''pattern.py:439-466'''
        time, pos, deriv = [], [], []
        with open(infile) as f:
            for line in f:
                q = [float(d) for d in line.split()]
                time.append(q[0]); pos.append(q[1])
                if len(q) > 2:
                    deriv.append(q[2])
        self.time = np.array(time); self.pos = np.array(pos); self.deriv = np.array(deriv)
        if len(self.deriv) < 1:
            BarDetermine.frequency_and_derivative(self)


def find_barpattern(intime, BarInstance, smth_order=2):
    '''pattern.py:549-578: the tabulated derivative nearest in time (scalar or array of times)'''
    BarInstance.frequency_and_derivative(smth_order=smth_order)
    if np.ndim(intime) > 0:
        return np.array([BarInstance.deriv[abs(t - BarInstance.time).argmin()] for t in intime])
    return BarInstance.deriv[abs(intime - BarInstance.time).argmin()]
