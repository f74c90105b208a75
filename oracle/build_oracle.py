"""build_oracle.py -- compile the oracle's C restatement (test infrastructure) with gcc.
Output: oracle/_build/libbfe_oracle.so (git-ignored; travels to the GPU box with the snapshot)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_build', 'libbfe_oracle.so')
SRC = os.path.join(HERE, 'bfe_oracle.c')


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -O2 without -ffast-math / FMA contraction: the arithmetic stays a faithful FP64 restatement
    cmd = ['gcc', '-O2', '-fopenmp', '-ffp-contract=off', '-fPIC', '-shared', SRC, '-o', OUT, '-lm']
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force=True))
