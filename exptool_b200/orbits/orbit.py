"""
orbit.py -- the return type of leapfrog_integrate (exptool/orbits/orbit.py:50):
a dict subclass with keys T,X,Y,Z,VX,VY,VZ,P,TX,TY,VTX,VTY[,FX,FY,FZ].
PSP orbit mapping and plotting are out of scope.
"""


class Orbits(dict):

    def __init__(self):
        pass
