"""TEST INFRASTRUCTURE: loads tests/emul/libathk_emul.so (host-compiled operator bodies) so index
logic can be debugged without a GPU.  Never imported by the athena_b200 package."""
import ctypes
import os
import subprocess

from athena_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))


def load_emul(build=True):
    so = os.path.join(HERE, "libathk_emul.so")
    if build:
        subprocess.run(["make", "-s", "-C", HERE], check=True, capture_output=True)
    return capi.bind(ctypes.CDLL(so))
