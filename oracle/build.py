"""Builds the C part of the CPU oracle (test infrastructure only) into oracle/_build/."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SO = os.path.join(OUT, "libs3oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "knn_oracle.c")
    os.makedirs(OUT, exist_ok=True)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src, "-o", SO, "-lm"])
    return SO


if __name__ == "__main__":
    print(build(True))
