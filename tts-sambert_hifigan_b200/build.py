"""nvcc recipe for libhfg_b200.so (sm_100a only; built in-tree so the .so
travels to the GPU box with the repo snapshot)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libhfg_b200.so")
# Same sources with -DHFG_TUNING: HFG_TC_* environment knobs, clock64 timelines and the "switch parts of the
# kernel off" experiments exist only in this build (tools/, kernel-variant tests; loaded via HFG_LIB_PATH).
LIB_TUNING = os.path.join(LIB_DIR, "libhfg_b200_tuning.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--threads", "4",           # the .cu files of one library compile side by side
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) \
        + glob.glob(os.path.join(INCLUDE, "*.h"))


def is_stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(p) > t for p in _deps())


def _compile(lib: str, extra, verbose: bool):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError(f"nvcc not found: cannot build {os.path.basename(lib)}")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed building {os.path.basename(lib)}")


def build(force: bool = False, verbose: bool = False, tuning: bool = True) -> str:
    """Production library, and (tuning=True) the instrumented build next to it."""
    jobs = []
    if force or is_stale(LIB):
        jobs.append((LIB, [], verbose))
    if tuning and (force or is_stale(LIB_TUNING)):
        jobs.append((LIB_TUNING, ["-DHFG_TUNING"], False))
    if len(jobs) > 1:                        # the two libraries are independent: build them at the same time
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(jobs)) as ex:
            for f in [ex.submit(_compile, *j) for j in jobs]:
                f.result()
    elif jobs:
        _compile(*jobs[0])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
