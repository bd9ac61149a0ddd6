"""Build the oracle's C restatement (oracle/ipp_oracle.c -> oracle/_build/libipp_oracle.so).

TEST INFRASTRUCTURE: the product library never links this.  The reference itself is pure Python
(no C/C++ on the hot path), so there is no ``oracle/_ref`` binary to compile: the reference was
executed in the build container to produce ``tests/golden`` instead (tests/golden/make_golden.py).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ipp_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libipp_oracle.so")


def build_oracle(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
