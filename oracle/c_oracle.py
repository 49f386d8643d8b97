"""TEST INFRASTRUCTURE - builds and binds the plain-C restatement of the neighbour-table oracle
(oracle/neighbors_oracle.c).  ``build()`` is called by ``__graft_entry__.build()`` (building the checker is not using
it); ``neighbor_table`` / ``fp32_order_mismatches`` are for tests/ and bench.py's CPU legs only."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "neighbors_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liblantern_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: cannot build the C oracle")
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [gcc, "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"C oracle build failed:\n{r.stdout}")
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.lantern_oracle_neighbor_table.restype = C.c_int
        lib.lantern_oracle_neighbor_table.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        lib.lantern_oracle_fp32_order_mismatches.restype = C.c_int64
        lib.lantern_oracle_fp32_order_mismatches.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        _lib = lib
    return _lib


def neighbor_table(E: np.ndarray, K: int | None = None) -> np.ndarray:
    """Same result as ``lantern_oracle.neighbor_table`` (bit for bit), all host cores."""
    E = np.ascontiguousarray(E, dtype=np.float32)
    N, d = E.shape
    K = N - 1 if K is None else int(K)
    out = np.empty((N, K), dtype=np.int32)
    rc = load().lantern_oracle_neighbor_table(E.ctypes.data, N, d, K, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("lantern_oracle_neighbor_table failed")
    return out


def fp32_order_mismatches(E: np.ndarray, exact: np.ndarray) -> int:
    """Positions of ``exact`` [N, K] at which the fp32 expanded-form distance order (what the reference's
    ``torch.cdist`` + ``topk`` computes, ties by id) differs - reported, never gated."""
    E = np.ascontiguousarray(E, dtype=np.float32)
    exact = np.ascontiguousarray(exact, dtype=np.int32)
    N, d = E.shape
    return int(load().lantern_oracle_fp32_order_mismatches(E.ctypes.data, N, d, exact.shape[1], exact.ctypes.data))
