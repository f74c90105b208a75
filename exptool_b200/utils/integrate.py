"""
integrate.py -- drop-in for exptool.utils.integrate.leapfrog_integrate
(integrate.py:53-190) plus a batched form for many orbits.

The time loop runs in one CUDA kernel (csrc/bfe_field.cu: leapfrog_kernel), one
thread per orbit with the phase-space state in registers; the host only sets the
field truncation, launches, and formats the Orbits dictionary.

compute_timestep, the four integrate_grid* drivers and the orbit-map text format
 are mirrored too: a whole (radius, velocity[, z, vz]) grid is one
batched time-step estimate plus one leapfrog launch with a step size per orbit.
do_integrate_multi and the re-forming helpers are mirrored (the Pool fan-out becomes ranks / one batched launch);
run_time / run_time_mod (the notebook workflow drivers) are mirrored at the end of this file.  leapfrog_integrate accepts
the reference's duck-typed FieldInstance: an exptool_b200 Fields runs on the GPU, any other object with
set_field_parameters / return_forces_cart is integrated with its own force routine, as the reference would.
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
        raise TypeError('this entry point needs an exptool_b200.basis.potential.Fields (the time loop runs on the GPU); '
                        'leapfrog_integrate itself also accepts any object with set_field_parameters / return_forces_cart')
    return FieldInstance.device_handles()


def _precision(FieldInstance):
    """kept for callers of round 1: table precision lives in the Fields instance's device handles now"""
    return ops.table_precision(getattr(FieldInstance, 'table_fp32', False))


def _leapfrog_generic(FieldInstance, nint, dt, initpos, initvel, rotfreq, force, ap_max, apse):
    """integrate.leapfrog_integrate (integrate.py:94-190) for a FOREIGN field object -- anything with
    set_field_parameters / return_forces_cart, the reference's duck-typed protocol (integrate.py:61-64).  The force
    evaluation belongs to the caller's object, so there is nothing to put on the GPU: the reference's own scalar loop."""
    times = np.arange(0, nint, 1) * dt
    barpos = 2. * np.pi * rotfreq * times
    X = np.zeros(nint); Y = np.zeros(nint); Z = np.zeros(nint)
    VX = np.zeros(nint); VY = np.zeros(nint); VZ = np.zeros(nint)
    FX = np.zeros(nint); FY = np.zeros(nint); FZ = np.zeros(nint); P = np.zeros(nint)
    X[0], Y[0], Z[0] = initpos[0], initpos[1], initpos[2]
    VX[0], VY[0], VZ[0] = initvel[0], initvel[1], initvel[2]
    step = 1
    dfx, hfx, dfy, hfy, dfz, hfz, dp, hp = FieldInstance.return_forces_cart(X[0], Y[0], Z[0], rotpos=barpos[0])
    FX[0], FY[0], FZ[0], P[0] = dfx + hfx, dfy + hfy, dfz + hfz, dp + hp
    n_aps = 0
    while (n_aps < ap_max) & (step < nint):                                                   # 126
        X[step] = X[step - 1] + (VX[step - 1] * dt) + (0.5 * FX[step - 1] * (dt ** 2.))       # 129-131
        Y[step] = Y[step - 1] + (VY[step - 1] * dt) + (0.5 * FY[step - 1] * (dt ** 2.))
        Z[step] = Z[step - 1] + (VZ[step - 1] * dt) + (0.5 * FZ[step - 1] * (dt ** 2.))
        dfx, hfx, dfy, hfy, dfz, hfz, dp, hp = FieldInstance.return_forces_cart(X[step], Y[step], Z[step], rotpos=barpos[step])
        FX[step], FY[step], FZ[step], P[step] = dfx + hfx, dfy + hfy, dfz + hfz, dp + hp
        VX[step] = VX[step - 1] + (0.5 * (FX[step - 1] + FX[step]) * dt)                      # 141-143
        VY[step] = VY[step - 1] + (0.5 * (FY[step - 1] + FY[step]) * dt)
        VZ[step] = VZ[step - 1] + (0.5 * (FZ[step - 1] + FZ[step]) * dt)
        if apse and step > 1:                                                                 # 146-153
            r0 = X[step - 2] * X[step - 2] + Y[step - 2] * Y[step - 2]
            r1 = X[step - 1] * X[step - 1] + Y[step - 1] * Y[step - 1]
            r2 = X[step] * X[step] + Y[step] * Y[step]
            if (r1 > r0) & (r1 > r2):
                n_aps += 1
        step += 1
    traj = np.stack([X, Y, Z, VX, VY, VZ, P, FX, FY, FZ], axis=1)[:step]
    return step, traj


def leapfrog_integrate(FieldInstance, nint, dt, initpos, initvel, rotfreq=0., no_odd=False,
                       halo_l=-1, halo_n=-1, disk_m=-1, disk_n=-1, verbose=0, force=False, ap_max=1000, apse=False):
    '''
    integrate.leapfrog_integrate (integrate.py:53-190) for one orbit -> Orbits dict with
    T, X, Y, Z, VX, VY, VZ, P [, FX, FY, FZ], TX, TY, VTX, VTY, each of length `step`.
    '''
    t0 = time.time()
    if not hasattr(FieldInstance, 'device_handles'):
        # the reference's protocol is duck-typed (integrate.py:61-64, 89, 118): a foreign field object is integrated with
        # its own return_forces_cart, exactly as the reference would
        FieldInstance.set_field_parameters(no_odd=no_odd, halo_l=halo_l, halo_n=halo_n, disk_m=disk_m, disk_n=disk_n)
        step, traj = _leapfrog_generic(FieldInstance, nint, dt, initpos, initvel, rotfreq, force, ap_max, apse)
    else:
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


# ---------------------------------------------------------------------------
# time-step heuristic and orbit grids -- integrate.py:217-272, 760-922
# ---------------------------------------------------------------------------
def compute_timestep(FieldInstance, start_pos, start_vel, dyn_res=100., verbose=False):
    '''
    integrate.compute_timestep (integrate.py:217-272): min of the four EXP criteria / dyn_res.
    start_pos, start_vel: [x,y,z] (scalars -> float) or (3, norbit) arrays -> array of dt.
    The reference's `dfz*hfz` product (249, 251; a sum is meant) is reproduced.
    '''
    eps = 1.e-10
    pos = np.asarray(start_pos, dtype=np.float64)
    vel = np.asarray(start_vel, dtype=np.float64)
    scalar = pos.ndim == 1
    pos = pos.reshape(3, -1)
    vel = vel.reshape(3, -1)
    vtot = np.sum(vel ** 2., axis=0) ** 0.5 + eps
    dfx, hfx, dfy, hfy, dfz, hfz, dp, hp = FieldInstance.return_forces_cart(pos[0], pos[1], pos[2])
    dtr = pos[0] * (dfx + hfx) + pos[1] * (dfy + hfy) + pos[2] * (dfz * hfz)
    atot = ((dfx + hfx) ** 2. + (dfy + hfy) ** 2. + (dfz * hfz) ** 2.0) ** 0.5 + eps
    ptot = np.abs(dp + hp)
    T = [1.0 / np.sqrt(vtot + eps), np.sqrt(vtot / (atot + eps)), ptot / (np.abs(dtr) + eps),
         np.sqrt(ptot / (atot * atot + eps))]
    dt = np.min(np.array(T), axis=0) / dyn_res
    if scalar:
        return dt[0]
    return dt


def _grid_orbits(F, pos0, vel0, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max):
    """shared body of the grid drivers: per-orbit dt, one leapfrog launch, host-side bar-frame rotation"""
    E, H = _handles(F, no_odd, halo_l, -1, disk_m, -1)
    dts = np.maximum(compute_timestep(F, pos0, vel0, dyn_res=dyn_res), dt)       # integrate.py:871
    with _precision(F):
        state, traj, nsteps = ops.leapfrog(E, H, pos0, vel0, nint, dts, rotfreq=rotfreq, traj_stride=1,
                                           apse=False, ap_max=ap_max)
    traj = ops.to_host(traj)                                   # (nint, 10, norb)
    times = np.arange(0, nint, 1)[:, None] * dts[None, :]       # (nint, norb)
    barpos = 2. * np.pi * rotfreq * times
    X, Y, VX, VY = traj[:, 0], traj[:, 1], traj[:, 3], traj[:, 4]
    TX = np.empty_like(X); TY = np.empty_like(Y)
    neg = np.min(barpos, axis=0) < 0.                            # integrate.py:180-188, per orbit
    tx, ty = transform(X, Y, barpos)
    cx, cy = clock_transform(X, Y, barpos)
    TX = np.where(neg[None, :], tx, cx)
    TY = np.where(neg[None, :], ty, cy)
    return traj, TX, TY, times


def _integrate_grid_2d(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, launch_y):
    rads = np.asarray(rads, dtype=np.float64); vels = np.asarray(vels, dtype=np.float64)
    R, V = np.meshgrid(rads, vels, indexing='ij')
    n = R.size
    pos0 = np.zeros((3, n)); vel0 = np.zeros((3, n))
    pos0[1 if launch_y else 0] = R.reshape(-1)
    vel0[1] = V.reshape(-1)
    traj, TX, TY, times = _grid_orbits(F, pos0, vel0, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max)
    Oarray = np.zeros([rads.size, vels.size, 7, nint])
    for k, a in enumerate((traj[:, 0], traj[:, 1], TX, TY, traj[:, 3], traj[:, 4], times)):
        Oarray[:, :, k, :] = a.T.reshape(rads.size, vels.size, nint)
    return Oarray


def integrate_grid(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, verbose):
    '''integrate.integrate_grid (integrate.py:848-883): Oarray[rad, vel, (X,Y,TX,TY,VX,VY,T), nint], launched
    from the x axis with dt as the minimum step.'''
    return _integrate_grid_2d(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, False)


def integrate_grid_launchy(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, verbose):
    '''integrate.integrate_grid_launchy (integrate.py:886-922): same, launched from the y axis.'''
    return _integrate_grid_2d(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, True)


def _integrate_grid_3d(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, zs, vzs, launch_y):
    if (zs is None) or (vzs is None):
        print('ERROR: 3D orbit specified, but no z or vz values passed!')
        return None
    rads, vels, zs, vzs = (np.asarray(a, dtype=np.float64) for a in (rads, vels, zs, vzs))
    R, V, Z, VZ = np.meshgrid(rads, vels, zs, vzs, indexing='ij')
    n = R.size
    pos0 = np.zeros((3, n)); vel0 = np.zeros((3, n))
    pos0[1 if launch_y else 0] = R.reshape(-1)
    pos0[2] = Z.reshape(-1)
    vel0[1] = V.reshape(-1)
    vel0[2] = VZ.reshape(-1)
    traj, TX, TY, times = _grid_orbits(F, pos0, vel0, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max)
    Oarray = np.zeros([rads.size, vels.size, zs.size, vzs.size, 9, nint])
    for k, a in enumerate((traj[:, 0], traj[:, 1], traj[:, 2], TX, TY, traj[:, 3], traj[:, 4], traj[:, 5], times)):
        Oarray[:, :, :, :, k, :] = a.T.reshape(rads.size, vels.size, zs.size, vzs.size, nint)
    return Oarray


def integrate_grid_3D(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, zs, vzs, verbose):
    '''integrate.integrate_grid_3D (integrate.py:760-801): Oarray[rad,vel,z,vz,(X,Y,Z,TX,TY,VX,VY,VZ,T),nint]'''
    return _integrate_grid_3d(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, zs, vzs, False)


def integrate_grid_3D_launchy(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, zs, vzs, verbose):
    '''integrate.integrate_grid_3D_launchy (integrate.py:803-845)'''
    return _integrate_grid_3d(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, zs, vzs, True)


# ---------------------------------------------------------------------------
# the reference's process fan-out over radii -- integrate.py:290-333, 440-507.  Every orbit of the grid is one
# thread of one leapfrog launch here (across ranks when torch.distributed is up: the radii are block-partitioned
# exactly like redistribute_arrays, each rank integrates its block and the f4 blocks are gathered).
# ---------------------------------------------------------------------------
def redistribute_arrays(rads, divisions):
    '''integrate.redistribute_arrays (integrate.py:440-463): block 0 takes the remainder.'''
    from .. import parallel
    rads = np.asarray(rads)
    return [rads[lo:hi] for lo, hi in parallel.shard_bounds(rads.size, divisions)]


def re_form_orbit_arrays(array):
    '''integrate.re_form_orbit_arrays (integrate.py:469-489): concatenate along the radius axis, cast to f4.'''
    return np.concatenate([np.asarray(a) for a in array], axis=0).astype('f4')


def re_form_orbit_arrays_3D(array):
    '''integrate.re_form_orbit_arrays_3D (integrate.py:491-507)'''
    return np.concatenate([np.asarray(a) for a in array], axis=0).astype('f4')


def do_integrate_multi(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max,
                       verbose=0, nprocs=-1, threedee=False, zs=None, vzs=None, launch='x'):
    '''
    integrate.do_integrate_multi (integrate.py:290-333): the orbit grid as one f4 array
    [rad, vel, 7, nint] (or [rad, vel, z, vz, 9, nint] with threedee).  `nprocs` is accepted for call
    compatibility; the whole grid is one device launch per rank.
    '''
    import time
    from .. import parallel
    rads = np.asarray(rads, dtype=np.float64); vels = np.asarray(vels, dtype=np.float64)
    t1 = time.time()
    rank, ws = parallel.world()
    mine = redistribute_arrays(rads, ws)[rank] if ws > 1 else rads
    launch_y = (launch == 'y')
    if threedee:
        if (zs is None) or (vzs is None):
            print('ERROR: 3D orbit specified, but no z or vz values passed!')
            return None
        block = _integrate_grid_3d(mine, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, zs, vzs,
                                   launch_y)
    else:
        block = _integrate_grid_2d(mine, vels, F, nint, dt, rotfreq, no_odd, halo_l, disk_m, dyn_res, ap_max, launch_y)
    block = block.astype('f4')
    if ws > 1:
        import torch.distributed as dist
        blocks = [None] * ws
        dist.all_gather_object(blocks, block)
        block = re_form_orbit_arrays_3D(blocks) if threedee else re_form_orbit_arrays(blocks)
    if verbose > 0:
        print('Total integration calculation took {0:3.2f} seconds, or {1:3.2f} seconds per orbit.'
              .format(time.time() - t1, (time.time() - t1) / max(float(rads.size * vels.size), 1.)))
    return block


# ---------------------------------------------------------------------------
# orbit-map text files -- integrate.py:513-569, 970-1038 (one line per orbit)
# ---------------------------------------------------------------------------
def print_orbit_array(f, OrbitArray):
    '''integrate.print_orbit_array (513-537): `nsteps R0 V0 dT x.. y.. tx.. ty.. vx.. vy..`'''
    for rad in range(0, OrbitArray.shape[0]):
        for vel in range(0, OrbitArray.shape[1]):
            nsteps = OrbitArray.shape[-1]
            print(nsteps, OrbitArray[rad, vel, 0, 0], OrbitArray[rad, vel, 5, 0], OrbitArray[rad, vel, -1, 1],
                  end=' ', file=f)
            for row in range(6):
                for x in range(0, nsteps):
                    print(OrbitArray[rad, vel, row, x], end=' ', file=f)
            print('', file=f)


def print_orbit_array_3D(f, OrbitArray):
    '''integrate.print_orbit_array_3D (539-569)'''
    for rad in range(0, OrbitArray.shape[0]):
        for vel in range(0, OrbitArray.shape[1]):
            for z in range(0, OrbitArray.shape[2]):
                for vz in range(0, OrbitArray.shape[3]):
                    nsteps = OrbitArray.shape[-1]
                    O = OrbitArray[rad, vel, z, vz]
                    print(nsteps, O[0, 0], O[6, 0], O[2, 0], O[7, 0], O[-1, 1], end=' ', file=f)
                    for row in range(8):
                        for x in range(0, nsteps):
                            print(O[row, x], end=' ', file=f)
                    print('', file=f)


def read_integrations(infile):
    '''integrate.read_integrations (970-1001)'''
    D = {k: {} for k in ('R_0', 'V_0', 'dT', 'X', 'Y', 'TX', 'TY', 'VX', 'VY')}
    with open(infile, 'r') as f:
        for linenum, line in enumerate(f):
            d = [float(q) for q in line.split()]
            npoints = int(d[0])
            D['R_0'][linenum] = d[1]; D['V_0'][linenum] = d[2]; D['dT'][linenum] = d[3]
            for k, key in enumerate(('X', 'Y', 'TX', 'TY', 'VX', 'VY')):
                D[key][linenum] = d[(k * npoints) + 4:((k + 1) * npoints) + 4]
    return D


def read_integrations_3D(infile):
    '''integrate.read_integrations_3D (1003-1038)'''
    D = {k: {} for k in ('R_0', 'V_0', 'Z_0', 'VZ_0', 'dT', 'X', 'Y', 'Z', 'TX', 'TY', 'VX', 'VY', 'VZ')}
    with open(infile, 'r') as f:
        for linenum, line in enumerate(f):
            d = [float(q) for q in line.split()]
            npoints = int(d[0])
            D['R_0'][linenum] = d[1]; D['V_0'][linenum] = d[2]; D['Z_0'][linenum] = d[3]
            D['VZ_0'][linenum] = d[4]; D['dT'][linenum] = d[5]
            for k, key in enumerate(('X', 'Y', 'Z', 'TX', 'TY', 'VX', 'VY', 'VZ')):
                D[key][linenum] = d[(k * npoints) + 6:((k + 1) * npoints) + 6]
    return D


# ---------------------------------------------------------------------------
# workflow drivers -- integrate.py:571-755 (what the notebooks call)
# ---------------------------------------------------------------------------
def run_time(simulation_directory, simulation_name, eof_file, sph_file, model_file, intime, rads, vels, nint, dt, no_odd, halo_l,
             max_m, dyn_res, ap_max, verbose, nprocs=-1, omegap=-1., orbitfile='', transform=True, fileprefix='OUT', launch='x'):
    '''integrate.run_time (integrate.py:571-635): fields of one snapshot -> orbit grid -> omap text file.  As in the
    reference, get_fields is called without a bar file here.'''
    from ..basis import potential
    if verbose:
        print('exptool.integrate.run_time: in directory {}, run {} at output {}, with transform={}'.format(
            simulation_directory, simulation_name, intime, transform))
    F, patt, rotfreq = potential.get_fields(simulation_directory, simulation_name, intime, eof_file, sph_file, model_file,
                                            transform=transform, fileprefix=fileprefix)
    if omegap >= 0.:
        patt = omegap
    rotfreq = -1. * abs(patt / (2. * np.pi))
    OrbitArray = do_integrate_multi(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, max_m, dyn_res, ap_max, verbose=verbose,
                                    nprocs=nprocs, launch=launch)
    fname = orbitfile if orbitfile != '' else simulation_directory + 'omap_' + str(intime) + '.txt'
    with open(fname, 'w') as f:
        print_orbit_array(f, OrbitArray)


def run_time_mod(simulation_directory, simulation_name, eof_file, sph_file, model_file, intime, rads, vels, nint, dt, no_odd,
                 halo_l, max_m, dyn_res, ap_max, verbose, nprocs=-1, omegap=-1., orbitfile='', transform=False, save_field=True,
                 field_file_name='', field_file=None, bar_file='', fileprefix='OUT', threedee=False, zs=None, vzs=None,
                 launch='x'):
    '''integrate.run_time_mod (integrate.py:638-755): as run_time, with the frozen-field file (save / restore), an explicit
    bar file for the transform, and the 3-D launch grid.'''
    from ..basis import potential
    from ..analysis import pattern
    from ..io import particle
    if verbose:
        print('exptool.integrate.run_time: in directory {}, run {} at output {}, with transform={}'.format(
            simulation_directory, simulation_name, intime, transform))
    if transform is True and bar_file == '':
        print('error - no bar file supplied for transform')
        return
    if field_file is None:
        F, patt, rotfreq = potential.get_fields(simulation_directory, simulation_name, intime, eof_file, sph_file, model_file,
                                                transform=transform, bar_file=bar_file, fileprefix=fileprefix)
        if save_field is True:
            if str(field_file_name) != '':
                F.save_field(str(field_file_name))
            else:
                print('saving field file with default name (field_file) to local directory')
                F.save_field('field_file')
    else:
        print('field file supplied! File name:')
        print(field_file_name)
        F = potential.restore_field(str(field_file_name))
        if transform:
            BarInstance = pattern.BarDetermine()
            BarInstance.read_bar(bar_file)
            BarInstance.frequency_and_derivative(spline_derivative=2)
            infile = simulation_directory + fileprefix + '.' + simulation_name + '.{0:05d}'.format(intime)
            PSPDump = particle.Input(infile, comp='star')
            patt = pattern.find_barpattern(PSPDump.time, BarInstance, smth_order=None)
        else:
            patt = 0.
    if omegap >= 0.:
        patt = omegap
    rotfreq = -1. * abs(patt / (2. * np.pi))
    if threedee is False:
        print('launching from ', launch, ' axis')
        OrbitArray = do_integrate_multi(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, max_m, dyn_res, ap_max,
                                        verbose=verbose, nprocs=nprocs, launch=launch)
        if orbitfile == '':
            print('Saving 2D orbit')
        fname = orbitfile if orbitfile != '' else simulation_directory + 'omap_' + str(intime) + '.txt'
        with open(fname, 'w') as f:
            print_orbit_array(f, OrbitArray)
    else:
        if (zs is None) or (vzs is None):
            print('ERROR: 3D orbit specified, but no z or vz values passed!')
            return
        OrbitArray = do_integrate_multi(rads, vels, F, nint, dt, rotfreq, no_odd, halo_l, max_m, dyn_res, ap_max,
                                        verbose=verbose, nprocs=nprocs, threedee=threedee, zs=zs, vzs=vzs, launch=launch)
        if orbitfile == '':
            print('Saving 3D orbit')
        fname = orbitfile if orbitfile != '' else simulation_directory + 'omap_3D_' + str(intime) + '.txt'
        with open(fname, 'w') as f:
            print_orbit_array_3D(f, OrbitArray)
