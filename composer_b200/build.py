'''
Builds ``libcomposer_b200.so`` (the sm_100a kernels + the C ABI of
``include/composer_b200.h``) in-tree with nvcc.  No torch dependency: the
library only needs the CUDA runtime.

    python -m composer_b200.build [--force]
'''

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PACKAGE_DIR, 'csrc')
BUILD_DIR = os.path.join(PACKAGE_DIR, 'build')
LIBRARY = os.path.join(PACKAGE_DIR, 'libcomposer_b200.so')
STAMP = LIBRARY + '.sha256'          # hash of the sources the library was built from (travels with it)
SOURCES = ['gemm.cu', 'elementwise.cu', 'attention.cu', 'attention_fwd_tc.cu', 'attention_tc.cu', 'decode.cu', 'decode_mega.cu', 'engine.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
    for candidate in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if candidate and (os.path.sep not in candidate or os.path.exists(candidate)):
            return candidate
    return 'nvcc'


def source_hash():
    '''SHA-256 over every file the library is built from (csrc/, include/) and the compiler flags.'''

    digest = hashlib.sha256(' '.join(NVCC_FLAGS + SOURCES).encode())
    for root in (CSRC, os.path.join(os.path.dirname(PACKAGE_DIR), 'include')):
        for name in sorted(os.listdir(root)):
            path = os.path.join(root, name)
            if os.path.isfile(path):
                digest.update(name.encode())
                with open(path, 'rb') as handle:
                    digest.update(handle.read())
    return digest.hexdigest()


def is_current():
    '''True when the library exists and was built from exactly the sources in the tree (hash, not mtime).'''

    if not (os.path.exists(LIBRARY) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as handle:
        return handle.read().strip() == source_hash()


def build(force=False, verbose=False):
    '''Compiles every .cu for sm_100a and links the shared library. Returns its path.'''

    if not force and is_current():
        return LIBRARY

    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(source):
        obj = os.path.join(BUILD_DIR, source.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, source), '-o', obj]
        if verbose:
            print(' '.join(cmd))
        result = subprocess.run(cmd, capture_output=True, text=True)
        if result.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (source, result.stdout, result.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objects = list(pool.map(compile_one, SOURCES))

    cmd = [nvcc, '-shared', '-o', LIBRARY] + objects + ['-lcudart']
    result = subprocess.run(cmd, capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (result.stdout, result.stderr))
    with open(STAMP, 'w') as handle:
        handle.write(source_hash() + '\n')
    return LIBRARY


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
