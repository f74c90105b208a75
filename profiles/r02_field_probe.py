"""Round-2 probe of the key-ordered per-point paths (bfe_orbit_sort.cu): C3-like point sets and C4-like orbit batches
under different options, against the caller-order kernels, with a bit-equality check.
   python profiles/r02_field_probe.py [--n 4000000] [--norb 1000000] [--steps 120] [--out gpurun_out/r02_field_probe.json]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S
import bench_configs as BC

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=4000000)
ap.add_argument('--norb', type=int, default=1000000)
ap.add_argument('--steps', type=int, default=120)
ap.add_argument('--lmax', type=int, default=6)
ap.add_argument('--out', default='')
ap.add_argument('--skip-points', action='store_true')
ap.add_argument('--skip-orbits', action='store_true')
ap.add_argument('--chunks', default='262144,524288,1048576,2097152')
ap.add_argument('--resorts', default='2,3,4,6,8,16')
ap.add_argument('--stage', type=int, default=-1, help='option stage_eval (default: library default)')
ap.add_argument('--subbits', type=int, default=-1, help='option key_subbits')
ap.add_argument('--opt', action='append', default=[], help='name=value library option, repeatable')
args = ap.parse_args()

if args.stage >= 0:
    ops.set_option('stage_eval', args.stage)
if args.subbits >= 0:
    ops.set_option('key_subbits', args.subbits)
for kv in args.opt:
    k_, v_ = kv.split('=')
    ops.set_option(k_, int(v_))
E = BC.eof_handle(); H = BC.sl_handle(args.lmax)
n = args.n; nd = n // 2
pd = BC.dev_particles('disc', nd, 3003); ph = BC.dev_particles('halo', n - nd, 3503)
c, s = E.accumulate(*[p[:1000000] for p in pd]); ch = H.accumulate(*[p[:1000000] for p in ph])
E.contract(c * 0.025, s * 0.025); H.contract(ch)
x = torch.cat([pd[0], ph[0]]); y = torch.cat([pd[1], ph[1]]); z = torch.cat([pd[2], ph[2]])


def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3      # us


res = dict(n=n, norb=args.norb, steps=args.steps, lmax=args.lmax)
defaults = {k: ops.get_option(k) for k in ('field_sort_min', 'field_sort_chunk', 'orbit_sort_min', 'orbit_resort', 'table_fp32', 'pdl')}
res['defaults'] = defaults
if not args.skip_points:
    pts = {}
    for f32 in (0, 1):
        ops.set_option('table_fp32', f32)
        ops.set_option('field_sort_min', 0)
        ref = ops.field_force_cart(E, H, x, y, z, rotpos=0.3)
        pts['caller_order_us_per_1e6_f32tab%d' % f32] = timeit(lambda: ops.field_force_cart(E, H, x, y, z, rotpos=0.3)) / n * 1e6
        ops.set_option('field_sort_min', 1)
        for chunk in [int(c) for c in args.chunks.split(',')]:
            ops.set_option('field_sort_chunk', chunk)
            got = ops.field_force_cart(E, H, x, y, z, rotpos=0.3)
            eq = bool(torch.equal(got, ref))
            t = timeit(lambda: ops.field_force_cart(E, H, x, y, z, rotpos=0.3)) / n * 1e6
            pts['key_order_chunk%d_us_per_1e6_f32tab%d' % (chunk, f32)] = t
            pts['key_order_chunk%d_equal_f32tab%d' % (chunk, f32)] = eq
        if f32 == 0:
            for pdl in (0, 1):
                ops.set_option('pdl', pdl); ops.set_option('field_sort_chunk', defaults['field_sort_chunk'])
                pts['key_order_default_chunk_pdl%d_us_per_1e6' % pdl] = timeit(lambda: ops.field_force_cart(E, H, x, y, z, rotpos=0.3)) / n * 1e6
            ops.set_option('pdl', defaults['pdl'])
            # disc-only and halo-only halves
            for name, sl in (('disc', slice(0, nd)), ('halo', slice(nd, n))):
                xs, ys, zs = x[sl].contiguous(), y[sl].contiguous(), z[sl].contiguous()
                ops.set_option('field_sort_min', 0)
                t0 = timeit(lambda: ops.field_force_cart(E, H, xs, ys, zs, rotpos=0.3)) / xs.numel() * 1e6
                ops.set_option('field_sort_min', 1)
                t1 = timeit(lambda: ops.field_force_cart(E, H, xs, ys, zs, rotpos=0.3)) / xs.numel() * 1e6
                pts['%s_only_us_per_1e6 [caller, key order]' % name] = [t0, t1]
    for k, v in defaults.items():
        ops.set_option(k, v)
    res['points'] = pts
    print(json.dumps(pts, indent=1), flush=True)

if not args.skip_orbits:
    norb, nint = args.norb, args.steps
    dd = S.exponential_disc(norb, 4004)
    pos0 = np.stack(dd[:3])
    ops.set_option('field_sort_min', 0)
    a = ops.field_force_cart(E, H, pos0[0], pos0[1], pos0[2]).cpu().numpy()
    ops.set_option('field_sort_min', defaults['field_sort_min'])
    R = np.sqrt(pos0[0] ** 2 + pos0[1] ** 2) + 1e-12
    fr = ((a[0] + a[1]) * pos0[0] + (a[2] + a[3]) * pos0[1]) / R
    vc = np.sqrt(np.maximum(-R * fr, 1e-12))
    rng = np.random.default_rng(44)
    f = rng.uniform(0.6, 1.1, norb)
    vel0 = np.stack([-pos0[1] / R * vc * f, pos0[0] / R * vc * f, 0.1 * vc * rng.standard_normal(norb)])
    P0, V0 = ops.dev(pos0), ops.dev(vel0)
    orb = {}
    for f32 in (0, 1):
        ops.set_option('table_fp32', f32)
        ops.set_option('orbit_resort', 0)
        ref, _, _ = ops.leapfrog(E, H, P0, V0, nint, 3e-4, rotfreq=-5.0)
        t = timeit(lambda: ops.leapfrog(E, H, P0, V0, nint, 3e-4, rotfreq=-5.0), reps=2, warm=1)
        orb['plain_ns_per_orbit_step_f32tab%d' % f32] = t * 1e3 / (norb * (nint - 1))
        for K in [int(k) for k in args.resorts.split(',')]:
            ops.set_option('orbit_resort', K); ops.set_option('orbit_sort_min', 1)
            got, _, _ = ops.leapfrog(E, H, P0, V0, nint, 3e-4, rotfreq=-5.0)
            eq = bool(torch.equal(got, ref))
            t = timeit(lambda: ops.leapfrog(E, H, P0, V0, nint, 3e-4, rotfreq=-5.0), reps=2, warm=1)
            orb['resort%d_ns_per_orbit_step_f32tab%d' % (K, f32)] = t * 1e3 / (norb * (nint - 1))
            orb['resort%d_equal_f32tab%d' % (K, f32)] = eq
    for k, v in defaults.items():
        ops.set_option(k, v)
    res['orbits'] = orb
    print(json.dumps(orb, indent=1), flush=True)
if args.out:
    with open(args.out, 'w') as fh:
        json.dump(res, fh, indent=1)
