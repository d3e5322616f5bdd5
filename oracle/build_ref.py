"""Builds the checkers that come from the REFERENCE itself into oracle/_ref/ (git-ignored; the built files travel to the GPU box):

* libekf_eigen.so — oracle/ekf_eigen.cpp (the reference's Eigen expression sequence of MHGating + UpdateJosephForm) against
  the reference's vendored Eigen 3.3.9;
* libxivo_ref_G<g>_F<f>.so — the reference's OWN estimator for the point-cloud path: its unmodified sources
  (src/{estimator,estimator_accessors,update,manager,feature,oos,group,graph,graphbase,mm,options,param,camera_manager,helpers,
  geometry,imu,princedormand,rk4,factory,tracker,fastbrief}.cpp, common/utils.cpp, vendored jsoncpp) compiled with g++ from where
  they lie under /root/reference, plus oracle/ref_wrap.cpp (C entry points) and oracle/ref_shim/ (type-only stand-ins for the
  OpenCV and glog headers, which are not installed here, and a no-op Canvas).  The reference's own build system is not used.
  kMaxGroup / kMaxFeature are compile-time in the reference (src/core.h:92-105): one library per (G, F).

Flags = the reference's effective Release flags (CMakeLists.txt:25-32: CMAKE_CXX_FLAGS "... -funroll-loops" followed by
CMAKE_CXX_FLAGS_RELEASE "-O3 -DNDEBUG") except -march=native, which cannot travel: the parity pin uses an x86-64-v3 (AVX2 + FMA)
build; for TIMING (bench.py's reference arm) a second build of the bench variant with -march=x86-64-v4 (AVX-512: what `native`
means on the GPU boxes' host CPUs) is made as libxivo_ref_G<g>_F<f>_v4.so and used when the host CPU has AVX-512.

Only possible where /root/reference exists (the authoring container); the GPU box uses the prebuilt files."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
EIGEN = REF + "/thirdparty/eigen"
OUT = os.path.join(HERE, "_ref")
BUILD = os.path.join(HERE, "_build", "ref")

REF_SRC = ["estimator", "estimator_accessors", "update", "manager", "feature", "oos", "group", "graph", "graphbase", "mm", "options", "param",
           "camera_manager", "helpers", "geometry", "imu", "princedormand", "rk4", "factory", "tracker", "fastbrief"]
JSON_SRC = ["json_reader", "json_value", "json_writer"]
VARIANTS = [(4, 14), (15, 30)]  # (kMaxGroup, kMaxFeature): BASELINE configs[1] (N = 89) and the reference default (N = 203)


def newer(srcs, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


def build_eigen_update():
    out = os.path.join(OUT, "libekf_eigen.so")
    src = os.path.join(HERE, "ekf_eigen.cpp")
    if newer([src], out):
        subprocess.check_call(["g++", "-O3", "-funroll-loops", "-march=x86-64-v3", "-DNDEBUG", "-DEIGEN_INITIALIZE_MATRICES_BY_ZERO", "-std=c++17", "-fPIC", "-shared", "-I", EIGEN, src, "-o", out])
        print("built", out)


def build_reference_estimator(G, F, march="x86-64-v3", suffix=""):
    out = os.path.join(OUT, f"libxivo_ref_G{G}_F{F}{suffix}.so")
    shim = os.path.join(HERE, "ref_shim")
    own = [os.path.join(HERE, "ref_wrap.cpp"), os.path.join(shim, "canvas_stub.cpp")]
    shim_files = [os.path.join(dp, f) for dp, _, fs in os.walk(shim) for f in fs]
    if not newer(own + shim_files + [__file__], out):
        return
    bdir = os.path.join(BUILD, f"G{G}_F{F}{suffix}")
    os.makedirs(bdir, exist_ok=True)
    inc = ["-I", shim, "-I", REF + "/src", "-I", REF + "/common", "-I", EIGEN, "-I", REF + "/thirdparty/sophus", "-I", REF + "/thirdparty/jsoncpp/include",
           "-I", REF + "/thirdparty/DBoW2/include", "-I", REF + "/thirdparty/pnp", "-I", REF + "/thirdparty/pnp/lambdatwist"]
    # the reference's effective flags (CMakeLists.txt:25-40) minus -march=native.  EIGEN_INITIALIZE_MATRICES_BY_ZERO: the reference
    # reads Eigen matrices it never initialised (src/estimator.cpp:183-190: only the diagonals of Ka / Kg are set, then
    # IMU::IMU CHECKs that Ca is upper triangular, src/imu.cpp:23-25) — with the macro they are the zeros the code assumes.
    flags = ["-O3", "-DNDEBUG", "-funroll-loops", f"-march={march}", "-std=c++17", "-fPIC", "-w", "-DSOPHUS_USE_BASIC_LOGGING", "-DGOOGLE_STRIP_LOG=1", "-DEIGEN_INITIALIZE_MATRICES_BY_ZERO",
             f"-DEKF_MAX_GROUPS={G}", f"-DEKF_MAX_FEATURES={F}"]
    units = [(REF + f"/src/{n}.cpp", n) for n in REF_SRC] + [(REF + "/common/utils.cpp", "utils")] + \
            [(REF + f"/thirdparty/jsoncpp/src/lib_json/{n}.cpp", n) for n in JSON_SRC] + [(own[0], "ref_wrap"), (own[1], "canvas_stub")]

    def cc(u):
        src, name = u
        obj = os.path.join(bdir, name + ".o")
        r = subprocess.run(["g++"] + flags + inc + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"reference source {src} failed to compile:\n{r.stderr[-3000:]}")
        return obj

    with ThreadPoolExecutor(max(1, min(8, os.cpu_count() or 1))) as ex:
        objs = list(ex.map(cc, units))
    subprocess.check_call(["g++", "-shared", "-o", out] + objs)
    print("built", out)


def main():
    if not os.path.isdir(EIGEN):
        print("reference not present; keeping the prebuilt files in", OUT)
        return 0
    os.makedirs(OUT, exist_ok=True)
    build_eigen_update()
    for G, F in VARIANTS:
        build_reference_estimator(G, F)
    build_reference_estimator(4, 14, "x86-64-v4", "_v4")  # timing build of the bench variant for AVX-512 hosts
    return 0


if __name__ == "__main__":
    sys.exit(main())
