#!/usr/bin/env python
"""cuview -- lists the CUDA devices the blocks can be given, the way the reference's `clview`
(lib/clview.cc:130-231) lists OpenCL platforms/devices with the ids to put into a block's
(platform type, device selector, platform id, device id) fields.  The B200 build has one
"platform" (CUDA); a block's devId picks the ordinal when devSelector == 2 (specific).
usage: python tools/cuview.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gr_clenabled_b200 import capi


def main():
    lib = capi.load()
    n = lib.clb200_device_count()
    print("libclenabled_b200 %s" % lib.clb200_version().decode())
    if n <= 0:
        print("No CUDA devices found.")
        return 1
    print("Platform Id: 0\nPlatform Name: NVIDIA CUDA (sm_100a build)\n")
    for i in range(n):
        buf = C.create_string_buffer(256)
        capi.check(lib.clb200_device_name(i, buf, 256))
        print("  Device Id: %d" % i)
        print("  Device Name: %s" % buf.value.decode())
        print("  Device Type: GPU")
        print("  Multiprocessors: %d" % lib.clb200_device_sm_count(i))
        print("  Block fields: openCLPlatformType=1 (GPU), devSelector=2 (specific), platformId=0, devId=%d\n" % i)
    return 0


if __name__ == "__main__":
    sys.exit(main())
