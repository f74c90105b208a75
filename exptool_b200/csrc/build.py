"""
build.py -- compile libbfe.so (sm_100a) in-tree with nvcc.

    python exptool_b200/csrc/build.py [--force] [--verbose]

Output: exptool_b200/libbfe.so (git-ignored; travels to the GPU box with the snapshot).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.environ.get('BFE_BUILD_OUT') or os.path.join(PKG, 'libbfe.so')     # BFE_BUILD_OUT + BFE_NVCC_FLAGS: variant builds
SOURCES = ['bfe_eof.cu', 'bfe_sl.cu', 'bfe_field.cu', 'bfe_sort.cu', 'bfe_sl_sort.cu', 'bfe_host.cu', 'bfe_ingest.cu', 'bfe_blocks.cu', 'bfe_peer.cu', 'bfe_orbit_sort.cu', 'bfe_peak.cu']
HEADERS = ['bfe_device.cuh', 'bfe_sortcore.cuh', 'bfe_internal.h', os.path.join('..', '..', 'include', 'bfe.h')]
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc_path():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) < t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    nvcc = nvcc_path()
    objdir = os.path.join(HERE, 'build', os.path.basename(OUT).replace('.so', '') if os.environ.get('BFE_BUILD_OUT') else '')
    os.makedirs(objdir, exist_ok=True)
    flags = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--use_fast_math=false'][:-1]
    flags += ['-Xptxas', '-v'] if verbose else []
    flags += [f for f in os.environ.get('BFE_NVCC_FLAGS', '').split() if f]
    procs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(HERE, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(objdir, s.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + ARCH + flags + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [nvcc] + ARCH + ['-shared', '-o', OUT] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
