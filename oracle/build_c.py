"""Builds oracle/_build/libthb_oracle.so from oracle/thb_oracle_c.c (gcc -O2 -fopenmp).  Test infrastructure."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "thb_oracle_c.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libthb_oracle.so")


def build(force=False):
    if os.path.exists(OUT) and not force and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", OUT, SRC, "-lm"])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
