"""Compile the C restatement of EPMGP (oracle/epmgp_oracle.c) with gcc.  TEST INFRASTRUCTURE."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "epmgp_oracle.c")
OUT = os.path.join(HERE, "_build", "libepmgp_oracle.so")


def build_oracle(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    r = subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", OUT, SRC, "-lm"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build_oracle(force=True))
