"""Compile fabolas_b200/csrc a second time with g++ against the fiber-based CUDA emulator
(tests/cusim/cusim.h).  TEST INFRASTRUCTURE: the resulting library is loaded only by tests."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "fabolas_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libfabolas_sim.so")


def build_sim(force=False):
    srcs = [os.path.join(CSRC, "api.cu")]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cusim.h"),
                                                                 os.path.join(ROOT, "include", "fabolas_b200.h")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(map(os.path.getmtime, deps)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-shared", "-fPIC", "-DFAB_CUSIM", "-DCUSIM_DEFINE_GLOBALS",
           "-include", os.path.join(HERE, "cusim.h"), "-x", "c++"] + srcs + ["-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("sim build failed:\n" + r.stderr[-6000:])
    return OUT


if __name__ == "__main__":
    print(build_sim(force=True))
