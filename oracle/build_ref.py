"""Builds oracle/_ref/libekf_eigen.so from oracle/ekf_eigen.cpp against the reference's vendored
Eigen (read-only include from /root/reference/thirdparty/eigen).  Only possible where /root/reference
exists (the authoring container); the GPU box uses the prebuilt file.  -march=x86-64-v3 (AVX2+FMA)
instead of the reference's -march=native so the binary runs on the GPU box's host CPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
EIGEN = "/root/reference/thirdparty/eigen"
OUT = os.path.join(HERE, "_ref", "libekf_eigen.so")


def main():
    if not os.path.isdir(EIGEN):
        print("reference Eigen not present; keeping prebuilt", OUT)
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-DNDEBUG", "-DEIGEN_INITIALIZE_MATRICES_BY_ZERO", "-std=c++17", "-fPIC", "-shared", "-I", EIGEN,
           os.path.join(HERE, "ekf_eigen.cpp"), "-o", OUT]
    subprocess.check_call(cmd)
    print("built", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
