"""Build libpavgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pav_b200.build [--force] [-v] [--variant NAME -DMACRO=VALUE ...]

The shared library lands at pav_b200/libpavgpu.so (git-ignored, travels with gpurun snapshots).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libpavgpu.so')
SOURCES = ['seqstore.cu', 'cigar.cu', 'density.cu', 'nccl_bcast.cu', 'lift.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr'] + os.environ.get('PAVGPU_NVCC_DEFS', '').split()   # e.g. -DHOM_MIN_BLOCKS=5 for tuning runs


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(REPO, 'include', 'pavgpu.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defs=()):
    """Default library, or with ``variant`` a tuning build ``libpavgpu.<variant>.so`` compiled with extra ``defs`` (-D flags):
    several variants can be built here (no GPU needed), travel to the GPU box with the tree and be A/B-timed in one call by
    pointing ``PAVGPU_LIB`` at them (pav_b200/_capi.py) -- no nvcc run on GPU time."""
    lib_path = LIB if variant is None else os.path.join(HERE, f'libpavgpu.{variant}.so')
    if variant is None and not force and not _stale():
        build_pyrows()
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    obj_dir = os.path.join(HERE, 'build') if variant is None else os.path.join(HERE, 'build', variant)
    os.makedirs(obj_dir, exist_ok=True)
    for s in srcs:
        o = os.path.join(obj_dir, os.path.basename(s) + '.o')
        objs.append(o)
        cmd = [NVCC] + FLAGS + list(defs) + ['-I', os.path.join(REPO, 'include'), '-I', CSRC, '-c', s, '-o', o]
        if verbose:
            cmd.insert(1, '-Xptxas')
            cmd.insert(2, '-v')
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out.decode())
        if p.returncode:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    tmp = lib_path + f'.{os.getpid()}.tmp'
    cmd = [NVCC, '-shared', '-o', tmp] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static', '-ldl', '-lrt', '-lpthread']
    subprocess.check_call(cmd)
    os.replace(tmp, lib_path)
    build_pyrows(force)
    return lib_path


PYROWS = os.path.join(HERE, '_pyrows.so')


def build_pyrows(force=False):
    """CPython helper (host-side row formatting), gcc, in-tree."""
    import sysconfig
    src = os.path.join(CSRC, 'pyrows.c')
    if not force and os.path.exists(PYROWS) and os.path.getmtime(PYROWS) >= os.path.getmtime(src):
        return PYROWS
    import numpy
    inc = sysconfig.get_paths()['include']
    tmp = PYROWS + f'.{os.getpid()}.tmp'
    subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-std=gnu11', '-I', inc, '-I', numpy.get_include(), '-o', tmp, src])
    os.replace(tmp, PYROWS)
    return PYROWS


if __name__ == '__main__':
    # python -m pav_b200.build [--force] [-v] [--variant NAME -DX=1 -DY=2 ...]
    name = sys.argv[sys.argv.index('--variant') + 1] if '--variant' in sys.argv else None
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, variant=name, defs=[a for a in sys.argv[1:] if a.startswith('-D')]))
