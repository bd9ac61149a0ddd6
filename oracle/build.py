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


LIB_NATIVE = os.path.join(OUT_DIR, "libipp_oracle_native.so")


def build_oracle(force: bool = False, native: bool = False) -> str:
    """``native``: -O3 -march=native for the machine this runs on (the CPU-baseline leg of bench.py builds it on the box
    whose cores it times; the portable -O2 build is the one that ships with the snapshot and checks parity)."""
    os.makedirs(OUT_DIR, exist_ok=True)
    lib = LIB_NATIVE if native else LIB
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= os.path.getmtime(SRC):
        return lib
    opt = ["-O3", "-march=native"] if native else ["-O2"]
    cmd = ["gcc"] + opt + ["-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-o", lib, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{res.stdout}\n{res.stderr}")
    return lib


if __name__ == "__main__":
    print(build_oracle(force=True))
