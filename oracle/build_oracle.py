"""Build recipe for the CPU oracle (test infrastructure).

    python oracle/build_oracle.py

compiles oracle/velocity_oracle.c -> oracle/_build/libvelocity_oracle.so with gcc.  The reference
(/root/reference) is pure Python + opencv-python, so there is no C/C++ reference source to compile
into oracle/_ref; the oracle is a restatement ("port") pinned by tests/golden/.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "velocity_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libvelocity_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
           "-fvisibility=hidden", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
