"""Build recipe of the sm_100a decoder library (nvcc, in-tree output).

    python -m offsetguided_b200.build          # builds offsetguided_b200/libogdecoder.so

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_PATH = os.path.join(PKG_DIR, 'libogdecoder.so')
SOURCES = ['og_api.cu', 'og_nms_topk.cu', 'og_fused.cu', 'og_limbs.cu', 'og_group.cu', 'og_resize.cu']
HEADERS = ['og_common.cuh', 'og_interp.cuh', 'og_prep.cuh', os.path.join('..', '..', 'include', 'og_decoder.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '-fmad=false',                 # contraction only where the code asks for fmaf explicitly
    '-Xcompiler', '-fPIC', '-shared',
    '-cudart', 'static',
]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False, defines=None, out=None):
    """Compile the library if it is missing or older than its sources.
    ``defines`` / ``out`` build a tuning variant (e.g. {'OG_K1_UNROLL': 4}) elsewhere."""
    if out is None and not force and not needs_build():
        return LIB_PATH
    out = out or LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS
    for key, val in (defines or {}).items():
        cmd.append('-D%s=%s' % (key, val))
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ['-o', out]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
