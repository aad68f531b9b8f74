"""Build libb200ipm.so in-tree with nvcc for sm_100a (no torch extension machinery: plain C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'b200ipm.cu')
OUT = os.path.join(HERE, 'libb200ipm.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))] + \
       [os.path.join(os.path.dirname(HERE), 'include', 'b200ipm.h')]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-shared',
           '-Xcompiler', '-fPIC', '-o', OUT, SRC]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
