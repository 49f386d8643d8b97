"""In-tree build of the C-ABI library (``lantern_b200/liblantern_b200.so``) with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the repo snapshot.  ``build()`` is incremental:
it recompiles only when a source or header is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from typing import List

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "liblantern_b200.so")
SOURCES = ["abi.cu", "accept.cu", "sample.cu", "kv_compact.cu", "neighbors.cu", "neighbors_tc.cu", "session.cu", "dyntree.cu", "draft_sample.cu", "greedy.cu", "call.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def _deps() -> List[str]:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "lantern_b200.h"))
    return deps


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, variant: str = "", flags: List[str] = ()) -> str:
    """``variant``: an A/B or profiling build next to the product library (``lantern_b200/variants/lib_<variant>.so``,
    extra nvcc ``flags``); select it at run time with LANTERN_B200_LIB."""
    lib = LIB
    if variant:
        os.makedirs(os.path.join(PKG, "variants"), exist_ok=True)
        lib = os.path.join(PKG, "variants", f"lib_{variant}.so")
    elif not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    log = []
    objdir = os.path.join(PKG, "build" + ("_" + variant if variant else ""))
    os.makedirs(objdir, exist_ok=True)
    procs = []
    extra = os.environ.get("LANTERN_EXTRA_NVCC_FLAGS", "").split() + list(flags)
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(f"$ {' '.join(link)}\n{r.stdout}")
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(objdir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:   # python -m lantern_b200.build --variant trace -DLANTERN_WALK_TRACE
        i = sys.argv.index("--variant")
        print(build(variant=sys.argv[i + 1], flags=sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
