"""Build libempanada_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m empanada_b200.build [--force] [--verbose]
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB = os.environ.get('EMP_B200_LIB') or os.path.join(LIB_DIR, 'libempanada_b200.so')   # override: tuning experiments only

NVCC_FLAGS = [
    '-O3', '-std=c++17',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo',
    '-Xcompiler', '-fPIC,-fvisibility=hidden,-pthread',
    '-fmad=false',          # never contract a*b+c: bit parity with torch CPU needs rn(dy*dy) first
    '-cudart', 'static',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'empanada_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = []
    procs = []
    extra = os.environ.get('EMP_NVCC_EXTRA', '').split()      # e.g. -DEMP_ASSIGN_CTAS_PER_SM=4 (tuning experiments)
    obj_dir = os.path.join(HERE, 'build' + ('_' + str(abs(hash(tuple(extra))) % 10000) if extra else ''))
    os.makedirs(obj_dir, exist_ok=True)
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(' '.join(cmd) + '\n' + out + '\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc, '-shared', '-cudart', 'static', '-gencode', 'arch=compute_100a,code=sm_100a',
           '-Xcompiler', '-pthread', '-o', LIB] + objs
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
