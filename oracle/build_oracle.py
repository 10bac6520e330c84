"""TEST INFRASTRUCTURE — build recipe of the C oracle (oracle/og_oracle.c -> oracle/libogoracle.so).

The reference is pure Python (no C sources to compile into oracle/_ref), so this is the
only native checker.  `python -m oracle.build_oracle`
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'og_oracle.c')
LIB = os.path.join(HERE, 'libogoracle.so')


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ['gcc', '-O2', '-ffp-contract=off', '-fopenmp', '-shared', '-fPIC', '-o', LIB, SRC, '-lm']
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('gcc failed:\n' + proc.stdout + proc.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
