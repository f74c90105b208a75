"""
integrate.py -- drop-in for exptool.utils.integrate.leapfrog_integrate
(integrate.py:53-190) plus a batched form for many orbits.

The time loop runs in one CUDA kernel (csrc/bfe_field.cu: leapfrog_kernel), one
thread per orbit with the phase-space state in registers; the host only sets the
field truncation, launches, and formats the Orbits dictionary.

Not mirrored: orbit-grid drivers, compute_timestep, orbit text files
(SURVEY.md section 8f rank 2).
"""
import time

import numpy as np

from ..orbits import orbit
from .. import ops


def transform(xarray, yarray, thetas):
    ''' counterclockwise planar transformation (integrate.py:37-41)'''
    new_xpos = np.cos(thetas) * xarray - np.sin(thetas) * yarray
    new_ypos = np.sin(thetas) * xarray + np.cos(thetas) * yarray
    return new_xpos, new_ypos


def clock_transform(xarray, yarray, thetas):
    ''' clockwise planar transformation (integrate.py:44-48)'''
    new_xpos = np.cos(thetas) * xarray + np.sin(thetas) * yarray
    new_ypos = -1. * np.sin(thetas) * xarray + np.cos(thetas) * yarray
    return new_xpos, new_ypos


def gen_init_step(xpos, vtan, z0=0.0, zvel0=0.):
    '''integrate.py:196-213'''
    return [xpos, 0.0, z0], [0.0, vtan, zvel0]


def _handles(FieldInstance, no_odd, halo_l, halo_n, disk_m, disk_n):
    FieldInstance.set_field_parameters(no_odd=no_odd, halo_l=halo_l, halo_n=halo_n, disk_m=disk_m, disk_n=disk_n)
    if not hasattr(FieldInstance, 'device_handles'):
        raise TypeError('leapfrog_integrate: FieldInstance must be an exptool_b200.basis.potential.Fields '
                        '(the time loop runs on the GPU; arbitrary Python force callbacks are not supported)')
    return FieldInstance.device_handles()


def leapfrog_integrate(FieldInstance, nint, dt, initpos, initvel, rotfreq=0., no_odd=False,
                       halo_l=-1, halo_n=-1, disk_m=-1, disk_n=-1, verbose=0, force=False, ap_max=1000, apse=False):
    '''
    integrate.leapfrog_integrate (integrate.py:53-190) for one orbit -> Orbits dict with
    T, X, Y, Z, VX, VY, VZ, P [, FX, FY, FZ], TX, TY, VTX, VTY, each of length `step`.
    '''
    t0 = time.time()
    E, H = _handles(FieldInstance, no_odd, halo_l, halo_n, disk_m, disk_n)
    pos0 = np.asarray(initpos, dtype=np.float64).reshape(3, 1)
    vel0 = np.asarray(initvel, dtype=np.float64).reshape(3, 1)
    state, traj, nsteps = ops.leapfrog(E, H, pos0, vel0, nint, dt, rotfreq=rotfreq, traj_stride=1,
                                       apse=apse, ap_max=ap_max)
    step = int(nsteps.cpu().numpy()[0])
    traj = traj.cpu().numpy()[:step, :, 0]
    times = np.arange(0, nint, 1) * dt
    barpos = (2. * np.pi * rotfreq * times)[0:step]
    if verbose:
        print('{0:4.3f} seconds to integrate.'.format(time.time() - t0))
    O = orbit.Orbits()
    O['T'] = times[0:step]
    O['X'] = traj[:, 0].copy(); O['Y'] = traj[:, 1].copy(); O['Z'] = traj[:, 2].copy()
    O['VX'] = traj[:, 3].copy(); O['VY'] = traj[:, 4].copy(); O['VZ'] = traj[:, 5].copy()
    O['P'] = traj[:, 6].copy()
    if force:
        O['FX'] = traj[:, 7].copy(); O['FY'] = traj[:, 8].copy(); O['FZ'] = traj[:, 9].copy()
    # integrate.py:180-188: rotation direction chosen by the sign of the bar position
    if np.min(barpos) < 0.:
        O['TX'], O['TY'] = transform(O['X'], O['Y'], barpos)
        O['VTX'], O['VTY'] = transform(O['VX'], O['VY'], barpos)
    else:
        O['TX'], O['TY'] = clock_transform(O['X'], O['Y'], barpos)
        O['VTX'], O['VTY'] = clock_transform(O['VX'], O['VY'], barpos)
    return O


def leapfrog_integrate_batch(FieldInstance, nint, dt, initpos, initvel, rotfreq=0., no_odd=False,
                             halo_l=-1, halo_n=-1, disk_m=-1, disk_n=-1, traj_stride=0, ap_max=1000, apse=False,
                             return_device=False):
    '''
    Batched extension: initpos, initvel are (3, norbit).  Returns a dict with the end
    state 'X','Y','Z','VX','VY','VZ' (norbit,), 'NSTEPS', and -- if traj_stride > 0 --
    'TRAJ' (nsave, 10, norbit) sampled every traj_stride steps (x,y,z,vx,vy,vz,pot,fx,fy,fz).
    With torch.distributed initialised each rank integrates its own block of orbits
    (no communication; outputs stay sharded).
    '''
    E, H = _handles(FieldInstance, no_odd, halo_l, halo_n, disk_m, disk_n)
    state, traj, nsteps = ops.leapfrog(E, H, initpos, initvel, nint, dt, rotfreq=rotfreq, traj_stride=traj_stride,
                                       apse=apse, ap_max=ap_max)
    if return_device:
        return dict(STATE=state, TRAJ=traj, NSTEPS=nsteps)
    s = state.cpu().numpy()
    out = dict(X=s[0], Y=s[1], Z=s[2], VX=s[3], VY=s[4], VZ=s[5], NSTEPS=nsteps.cpu().numpy())
    if traj is not None:
        out['TRAJ'] = traj.cpu().numpy()
    return out
