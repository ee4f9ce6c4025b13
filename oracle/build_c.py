"""Builds the plain-C part of the oracle (oracle/length_regulator.c) into oracle/_build/. Test infrastructure."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "liboracle_lr.so")


def build() -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    src = os.path.join(HERE, "length_regulator.c")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", OUT, src, "-lm"], check=True)
    return OUT


if __name__ == "__main__":
    print(build())
