"""Builds oracle/_build/libnms_oracle.so from oracle/nms_oracle.c (test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libnms_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "nms_oracle.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", OUT, src])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
