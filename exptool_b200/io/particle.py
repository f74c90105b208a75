"""
particle.py -- the particle container types the hot path accepts
(exptool/io/particle.py:142-157) and the `Input` front end of the PSP reader
(exptool/io/particle.py:16-112): "OUT." files go to psp_io, "SPL." split files to spl_io.
"""
from . import psp_io, spl_io


class Input():
    """particle.Input (particle.py:16-112): PSP 'OUT.' and split 'SPL.' snapshots -> .header, .time, .filename, .comp and either
    `.data` (dict) or, with legacy=True, the .xpos/.ypos/... attributes."""

    def __init__(self, filename, comp=None, legacy=False, verbose=0):
        if 'SPL.' in filename:                     # particle.py:68-77
            self.style = 'SPL'
            I = spl_io.Input(filename, comp=comp, verbose=verbose)
        else:
            self.style = 'OUT'
            I = psp_io.Input(filename, comp=comp, verbose=verbose)
        self.header = I.header
        self.filename = I.filename
        self.time = I.time
        if I.comp is None:
            return
        if legacy:
            self.mass = I.data['m']
            self.xpos = I.data['x']
            self.ypos = I.data['y']
            self.zpos = I.data['z']
            self.xvel = I.data['vx']
            self.yvel = I.data['vy']
            self.zvel = I.data['vz']
            self.pote = I.data['potE']
            if I.header[I.comp]['parameters']['indexing']:
                self.id = I.data['id']
        else:
            self.data = I.data
        self.comp = I.comp


class holder(object):
    '''legacy attribute container (exptool/io/particle.py:142-157)'''
    filename = None
    comp = None
    nbodies = None
    time = None
    xpos = None
    ypos = None
    zpos = None
    xvel = None
    yvel = None
    zvel = None
    mass = None
    pote = None
    id = None


def particle_arrays(P):
    """
    (x, y, z, m) from either container the reference accepts: a `holder`
    (.xpos/.ypos/.zpos/.mass, eof.py:525) or an object with a `.data` dict
    (keys 'x','y','z','m', eof.py:582).  m may be None for force evaluation.
    Real PSP files hold float32; parity is defined on FP64,
    so everything is cast to float64 here.
    """
    if isinstance(P, holder) or (not hasattr(P, 'data') and hasattr(P, 'xpos')):
        x, y, z, m = P.xpos, P.ypos, P.zpos, P.mass
    elif hasattr(P, 'data'):
        d = P.data
        x, y, z = d['x'], d['y'], d['z']
        try:
            m = d['m']
        except (KeyError, ValueError, IndexError):
            m = None
    elif isinstance(P, (tuple, list)) and len(P) in (3, 4):
        x, y, z = P[0], P[1], P[2]
        m = P[3] if len(P) == 4 else None
    else:
        raise TypeError('unsupported particle container %r' % type(P))
    return x, y, z, m


class Particles(object):
    """In-memory particle set with both interfaces (.data dict and holder attributes)."""

    def __init__(self, x, y, z, m, time=0.0, filename='memory', comp='memory'):
        self.data = {'x': x, 'y': y, 'z': z, 'm': m}
        self.xpos, self.ypos, self.zpos, self.mass = x, y, z, m
        self.time = time
        self.filename = filename
        self.comp = comp
        self.nbodies = len(x)
