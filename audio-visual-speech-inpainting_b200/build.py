"""In-tree build of libavsi_b200.so (nvcc, sm_100a only).

The shared library is written next to this file so that it travels with the repo snapshot
to the GPU box; it is git-ignored (history stays source-only).
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libavsi_b200.so')
BUILD_DIR = os.path.join(HERE, 'csrc', 'build')

SOURCES = ['capi.cu', 'frontend.cu', 'istft.cu', 'video.cu', 'gemm_sm100.cu', 'lstm.cu', 'lstm4.cu', 'lstm4_bwd.cu', 'loss.cu', 'ctc.cu', 'optim.cu', 'stats.cu', 'dropout.cu', 'ctc_decode.cu', 'features_extra.cu', 'tfrecord_host.cu', 'ssnn.cu', 'resample.cu']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-std=c++17', '-O3', '-lineinfo',
              '-Xcompiler', '-fPIC']


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build libavsi_b200.so')
    return exe


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ and link libavsi_b200.so.  Returns the library path."""
    srcs = sources()
    headers = [os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'sm100_ptx.cuh'), os.path.join(os.path.dirname(HERE), 'include', 'avsi_b200.h')]
    stamp = os.path.join(BUILD_DIR, 'stamp.txt')
    dig = _digest(srcs + headers)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        if verbose and r.stderr:
            print(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    with open(stamp, 'w') as f:
        f.write(dig)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force=True, verbose=True))
