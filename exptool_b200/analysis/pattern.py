"""
pattern.py -- BarTransform (exptool/analysis/pattern.py:64-169): rotate a snapshot into the bar frame.

The m = 2 Fourier sums and the planar rotation run on the device (ops.bar_fourier_angle, ops.affine_xy ->
bfe_bar_fourier, bfe_affine_xy); the class keeps the reference's interface (`.bar_angle`, `.data` dict with
x, y, z, vx, vy, vz, m, potE, `.time`, `.filename`, `.comp`).  The other classes of the reference module
(BarDetermine, pattern-speed fits) are post-processing and out of scope.
"""
import numpy as np

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
