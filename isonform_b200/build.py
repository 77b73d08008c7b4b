"""In-tree build of libisonform_b200.so (nvcc, sm_100a only).

    python -m isonform_b200.build [--force]

The shared library lands next to this file so that it travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libisonform_b200.so")
SO_DEBUG = os.path.join(HERE, "libisonform_b200_dbg.so")          # -DISOF_CHECK: device-side bounds assertions
SOURCES = ["capi.cu", "minimizers.cu", "sg_align.cu", "sg_dp_tb1.cu", "sg_dp_tb2.cu", "sg_dp_tb3.cu", "sg_dp_rb.cu", "sg_dp_rb2.cu", "sg_short.cu", "myers.cu", "spans.cu", "host_wis.cpp", "host_poa.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "isonform_b200.h"))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, debug=None):
    """debug=None builds both libraries, False/True only the release / assertion build."""
    if debug is None:
        with ThreadPoolExecutor(max_workers=2) as ex:          # release and assertion build side by side
            for f in [ex.submit(build, force, verbose, False), ex.submit(build, force, False, True)]:
                f.result()
        return SO
    deps = _deps()
    target = SO_DEBUG if debug else SO
    if not force and not _stale(target, deps):
        return target
    objs = []
    extra = ["-DISOF_CHECK"] if debug else []

    headers = [d for d in deps if d.endswith((".cuh", ".h"))]

    def compile_one(src):
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + (".dbg.o" if debug else ".o"))
        if not force and not verbose and not _stale(obj, headers + [os.path.join(CSRC, src)]):
            return obj                       # only the sources that changed (or all of them after a header change)
        cmd = [NVCC] + FLAGS + ["-Xcompiler", "-ffp-contract=off"] + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", target] + objs + ["-Xlinker", "--exclude-libs=ALL", "-lcudart_static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
