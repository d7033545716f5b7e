"""Build libplc.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libplc.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
    "--expt-relaxed-constexpr",
]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), "include", "plc.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into libplc.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libplc.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + _sources() + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


def build_kprof() -> str:
    """Developer build with in-kernel cycle counters compiled in (-DPLC_KPROF) -> tools/ubench/libplc_kprof.so;
    load it with PLC_LIB=tools/ubench/libplc_kprof.so (tools/kprof*.py)."""
    out = os.path.join(os.path.dirname(HERE), "tools", "ubench", "libplc_kprof.so")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-DPLC_KPROF"] + _sources() + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return out


def build_variant(defines, name: str) -> str:
    """Developer A/B build with extra -D flags -> tools/ubench/<name>.so (select it with PLC_LIB=...)."""
    out = os.path.join(os.path.dirname(HERE), "tools", "ubench", name + ".so")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v"] + ["-D" + d for d in defines] + _sources() + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    with open(out + ".ptxas.log", "w") as f:
        f.write(r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 2:], sys.argv[i + 1]))
    elif "--kprof" in sys.argv:
        print(build_kprof())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
