"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Build recipe for the CPU oracle (g++ only; no Eigen, no cmake). Two variants are
built from the same sources with identical results bit for bit:
  liboracle_fma.so      -mfma -mavx2   (std::fma inlined to the hardware FMA)
  liboracle_generic.so  baseline x86-64 (std::fma through libm, still correctly rounded)
Both use -ffp-contract=off so that only the explicit fma() calls are fused: the
canonical operation order of gi_oracle.hpp is then the same on every host.

`oracle/_ref/` (the real reference compiled from /root/reference) is NOT built:
every reference translation unit includes <Eigen/...> (include/jrl-qp/defs.h:5-6)
and Eigen is neither vendored in /root/reference nor installed in this image.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SOURCES = ["gi_oracle.cpp", "gi_oracle_capi.cpp", "decomp_oracle.cpp", "decomp_oracle_capi.cpp", "warm_oracle.cpp", "warm_oracle_capi.cpp", "kkt_oracle.cpp", "block_oracle.cpp"]
HEADERS = ["gi_oracle.hpp", "decomp_oracle.hpp", "warm_oracle.hpp", "block_oracle.hpp"]
BASE_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-pthread", "-DNDEBUG", "-Wall", "-Wextra"]
VARIANTS = {"fma": ["-mfma", "-mavx2"], "generic": []}
# A third build of the SAME sources for timing only (bench.py cpu_baseline.fast_port): -O3, AVX2 + FMA, and the compiler
# free to contract, reassociate and vectorise (-ffast-math): NOT bit-pinned, never used as a checker. It brackets what an
# optimised Eigen build of the reference could reach on the same cores beside the bit-pinned port.
FAST_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-shared", "-ffast-math", "-mfma", "-mavx2", "-funroll-loops", "-pthread", "-DNDEBUG"]


def _srcs():
    return [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = _srcs() + [os.path.join(HERE, h) for h in HEADERS if os.path.exists(os.path.join(HERE, h))]
    return any(os.path.getmtime(d) > t for d in deps)


def cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    fl = line.split()
                    return "fma" in fl and "avx2" in fl
    except OSError:
        pass
    return False


def build(force=False, verbose=False):
    """Compile both variants if missing or stale. Returns {variant: path}."""
    os.makedirs(OUT, exist_ok=True)
    paths = {}
    for name, extra in VARIANTS.items():
        target = os.path.join(OUT, f"liboracle_{name}.so")
        if force or _stale(target):
            cmd = ["g++"] + BASE_FLAGS + extra + _srcs() + ["-o", target]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.run(cmd, check=True)
        paths[name] = target
    return paths


def build_fast(force=False, verbose=False):
    """The timing-only variant (see FAST_FLAGS). Returns its path, or None on a host without AVX2 + FMA."""
    if not cpu_has_fma():
        return None
    os.makedirs(OUT, exist_ok=True)
    target = os.path.join(OUT, "liboracle_fast.so")
    if force or _stale(target):
        cmd = ["g++"] + FAST_FLAGS + _srcs() + ["-o", target]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return target


def library_path():
    paths = build()
    return paths["fma"] if cpu_has_fma() else paths["generic"]


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True), build_fast(force="--force" in sys.argv, verbose=True))
