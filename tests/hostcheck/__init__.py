"""Host build of the device math headers -- TEST INFRASTRUCTURE ONLY (see hostcheck.cpp)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.abspath(os.path.join(HERE, "..", "..", "differentiable_ransac_b200", "csrc"))
SO = os.path.join(HERE, "_hostcheck.so")


def build(force=False):
    src = os.path.join(HERE, "hostcheck.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-pthread", "-I", CSRC,
                               "-x", "c++", src, "-o", SO])
    return SO


def load():
    return ctypes.CDLL(build())
