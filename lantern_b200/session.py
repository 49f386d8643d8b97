"""Host-buffer session: the verify step for callers whose logits live in host memory (``lantern_session_*``).

``HostSession.step`` takes CPU tensors / numpy arrays.  Page-locked logits (``tensor.pin_memory()``) with a
lazy-eligible window are read in place over PCIe — only the rows the walk visits cross the bus; anything else is staged
through device memory.  Results come back as numpy arrays; the call is synchronous.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _abi
from .verify import Verifier


@dataclass
class HostResult:
    accept_length: np.ndarray     # [B] int32
    best_candidate: np.ndarray    # [B] int32
    token: np.ndarray             # [B] int32
    n_draws: np.ndarray           # [B] int32
    flags: np.ndarray             # [B] int32 (LANTERN_OUT_*; bits 8..15 = logits rows read for the item)
    path_tokens: np.ndarray       # [B, D] int32
    select_indices: np.ndarray    # [B, D] int32
    sample_p: Optional[np.ndarray]
    in_place: bool                # True: logits read in place from page-locked memory; False: staged copy

    @property
    def rows_read(self) -> np.ndarray:
        return (self.flags >> 8) & 0xFF


def _addr(x) -> int:
    return x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data


class HostSession:
    """Sized once for (B, T, L, D) upper bounds; ``verifier`` supplies the family, warp knobs and the table."""

    def __init__(self, verifier: Verifier, n_items: int, n_rows: int, n_paths: int, depth: int,
                 logits_dtype: torch.dtype = torch.float32, n_uniforms: Optional[int] = None):
        self.lib = _abi.load()
        self.ver = verifier
        self.dims = (n_items, n_rows, n_paths, depth)
        self.n_uniforms = n_rows + 1 if n_uniforms is None else n_uniforms
        proto = torch.empty(0, dtype=logits_dtype).new_empty((1, 1, verifier.family.vocab)).expand(n_items, n_rows, -1)
        cfg = verifier._cfg(n_items, n_rows, n_paths, depth, proto, False, self.n_uniforms, (0, 0))
        cfg.item_stride, cfg.row_stride = n_rows * verifier.family.vocab, verifier.family.vocab
        table = None
        if verifier.nbr_table is not None:
            table = np.ascontiguousarray(verifier.nbr_table.cpu().numpy().astype(np.int32))
        self._h = C.c_void_p()
        _abi.check(self.lib.lantern_session_create(C.byref(cfg), None if table is None else table.ctypes.data,
                                                   0 if table is None else table.shape[0], C.byref(self._h)))

    def close(self):
        if self._h:
            self.lib.lantern_session_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def step(self, logits_cond, logits_uncond, tree_tokens, retrieve, *, uniforms=None, row_kinds=None,
             philox=(0, 0), want_sample_p: bool = False) -> HostResult:
        """logits_*: [B, T, V] CPU tensors or numpy arrays (contiguous last dim), tree_tokens [B, T] int32, retrieve
        [B, L, D] int32 (-1 padded), uniforms [B, n] fp32 or None (device Philox stream), row_kinds [B, T] uint8."""
        B, T, V = logits_cond.shape
        L, D = retrieve.shape[-2:]
        as_t = (lambda x: x if isinstance(x, torch.Tensor) else torch.from_numpy(x))
        lc = as_t(logits_cond)
        n_uni = 0 if uniforms is None else uniforms.shape[1]
        cfg = self.ver._cfg(B, T, L, D, lc, retrieve.ndim == 2, n_uni, philox)
        ain = _abi.AcceptIn()
        ain.logits_cond = _addr(logits_cond)
        ain.logits_uncond = None if logits_uncond is None else _addr(logits_uncond)
        ain.tree_tokens, ain.retrieve = _addr(tree_tokens), _addr(retrieve)
        ain.uniforms = None if uniforms is None else _addr(uniforms)
        ain.row_kinds = None if row_kinds is None else _addr(row_kinds)
        ints = {n: np.zeros(B, dtype=np.int32) for n in ("accept_length", "best_candidate", "token", "n_draws", "flags")}
        path, sel = np.zeros((B, D), dtype=np.int32), np.zeros((B, D), dtype=np.int32)
        sp = np.zeros((B, V), dtype=np.float32) if want_sample_p else None
        aout = _abi.AcceptOut()
        for n, a in ints.items():
            setattr(aout, n, a.ctypes.data)
        aout.path_tokens, aout.select_indices = path.ctypes.data, sel.ctypes.data
        if sp is not None:
            aout.sample_p = sp.ctypes.data
        _abi.check(self.lib.lantern_session_step(self._h, C.byref(cfg), C.byref(ain), C.byref(aout)))
        return HostResult(ints["accept_length"], ints["best_candidate"], ints["token"], ints["n_draws"], ints["flags"],
                          path, sel, sp, int(self.lib.lantern_session_last_route(self._h)) == 1)
