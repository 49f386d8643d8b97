"""Device-tensor front end of the fused verify step (``lantern_accept_fused``).

PyTorch is plumbing here: it owns device memory and streams; every computation happens in the
CUDA library behind ``include/lantern_b200.h``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi


@dataclass(frozen=True)
class FamilySpec:
    """Per-family constants of the acceptance walk."""
    name: str
    family_id: int
    vocab: int
    col0: int
    ncols: int
    tok_offset: int = 0
    syntax_tokens: Tuple[int, ...] = ()
    newline_token: int = -1
    eoi_token: int = -1

    def resized(self, ncols: int, vocab: Optional[int] = None) -> "FamilySpec":
        """Same structure on a smaller vocabulary (tests, sweeps)."""
        if self.col0 == 0 and not self.syntax_tokens and self.family_id != _abi.FAMILY_ANOLE:
            return FamilySpec(self.name, self.family_id, ncols, 0, ncols, 0)
        vocab = vocab or max(ncols + 16, 8832)
        return FamilySpec(self.name, self.family_id, vocab, self.col0, ncols, self.tok_offset,
                          self.syntax_tokens, self.newline_token, self.eoi_token)


def vanilla(vocab: int) -> FamilySpec:
    """Plain EAGLE verification (models/drafters/utils.py:333-410)."""
    return FamilySpec("vanilla", _abi.FAMILY_VANILLA, vocab, 0, vocab)


LLAMAGEN = FamilySpec("llamagen", _abi.FAMILY_LLAMAGEN, 16384, 0, 16384)
ANOLE = FamilySpec("anole", _abi.FAMILY_ANOLE, 65536, 4, 8192, 4)
LUMINA = FamilySpec("lumina_mgpt", _abi.FAMILY_LUMINA, 65536, 4, 8192, 4, (8196, 8197, 8803, 8828), 8803, 8196)
FAMILIES = {"llamagen": LLAMAGEN, "anole": ANOLE, "lumina_mgpt": LUMINA}

_DTYPES = {torch.float32: _abi.F32, torch.bfloat16: _abi.BF16, torch.float16: _abi.F16}


@dataclass
class StaticTree:
    """Kernel-ready description of a static draft tree (shared by all items of a launch)."""
    retrieve: torch.Tensor      # [L, D] int32
    node_qrow: torch.Tensor     # [T] int32: row of draft_op with the node's sibling-group distribution
    sib_off: torch.Tensor       # [T+1] int32
    sib_idx: torch.Tensor       # [nnz] int32: earlier siblings (tree positions)
    n_q_rows: int


@dataclass
class VerifyResult:
    accept_length: torch.Tensor    # [B] int32
    best_candidate: torch.Tensor   # [B] int32
    token: torch.Tensor            # [B] int32
    path_tokens: torch.Tensor      # [B, D] int32
    select_indices: torch.Tensor   # [B, D] int32
    n_draws: torch.Tensor          # [B] int32
    flags: torch.Tensor            # [B] int32
    sample_p: Optional[torch.Tensor] = None   # [B, V] fp32


# lantern_accept_phases schedules: every tree row streamed through the statistics kernel, statistics computed lazily
# inside the walk, or chosen per batch by the library (profiles/sweep_r1.md)
_SCHEDULES = {"streamed": 3, "lazy": 6, "auto": 8}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Verifier:
    """One configured verification step; ``step`` launches it on the current CUDA stream."""

    def __init__(self, family: FamilySpec, *, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0,
                 cfg_scale: float = 1.0, lantern: bool = False, lantern_k: int = 1000, lantern_delta: float = 0.1,
                 nbr_table: Optional[torch.Tensor] = None, static_tree: Optional[StaticTree] = None,
                 device: Optional[torch.device] = None, schedule: str = "auto"):
        self.lib = _abi.load()
        if schedule not in _SCHEDULES:
            raise ValueError(f"schedule must be one of {sorted(_SCHEDULES)}")
        self.schedule = schedule
        self.family = family
        self.temperature = float(temperature)
        self.top_k = int(top_k)
        self.top_p = float(top_p)
        self.cfg_scale = float(cfg_scale)
        self.lantern = bool(lantern)
        self.lantern_k = int(lantern_k)
        self.lantern_delta = float(lantern_delta)
        self.static_tree = static_tree
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.temperature <= 1e-5:
            raise ValueError("temperature <= 1e-5 selects greedy decoding; use evaluate_posterior_greedy")
        self.nbr_table = None
        if self.lantern:
            if nbr_table is None:
                raise ValueError("lantern=True needs the neighbour table")
            self.nbr_table = self._as_table(nbr_table)
            if not (1 <= self.lantern_k <= self.nbr_table.shape[1]):
                raise ValueError(f"lantern_k={self.lantern_k} outside [1, {self.nbr_table.shape[1]}]")
        self._work = None
        self._cfg_cache = {}
        self._out_cache = {}

    def _as_table(self, t) -> torch.Tensor:
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(t.astype(np.int32))
        return t.to(device=self.device, dtype=torch.int32).contiguous()

    def _cfg(self, B, T, L, D, logits: torch.Tensor, retrieve_shared: bool, n_uniforms: int,
             philox: Tuple[int, int]) -> _abi.AcceptCfg:
        f = self.family
        c = _abi.AcceptCfg()
        c.n_items, c.n_rows, c.n_paths, c.depth = B, T, L, D
        c.vocab, c.col0, c.ncols = f.vocab, f.col0, f.ncols
        c.logits_dtype = _DTYPES[logits.dtype]
        c.item_stride, c.row_stride = logits.stride(0), logits.stride(1)
        c.family = f.family_id
        c.static_tree = 1 if self.static_tree is not None else 0
        c.cfg_scale, c.temperature, c.top_p, c.top_k = self.cfg_scale, self.temperature, self.top_p, self.top_k
        c.lantern, c.lantern_k = int(self.lantern), self.lantern_k
        c.lantern_delta = self.lantern_delta
        c.lantern_delta_m1 = float(np.float32(self.lantern_delta - 1.0))
        c.table_cols = int(self.nbr_table.shape[1]) if self.nbr_table is not None else 0
        c.tok_offset = f.tok_offset
        c.n_syntax = len(f.syntax_tokens)
        for i, tkn in enumerate(f.syntax_tokens):
            c.syntax_tokens[i] = tkn
        c.newline_token, c.eoi_token = f.newline_token, f.eoi_token
        c.retrieve_shared = int(retrieve_shared)
        c.n_uniforms = n_uniforms
        c.n_q_rows = self.static_tree.n_q_rows if self.static_tree is not None else 0
        c.philox_seed, c.philox_step = int(philox[0]) & (2**64 - 1), int(philox[1]) & (2**64 - 1)
        return c

    def cached_cfg(self, B, T, L, D, logits: torch.Tensor, shared: bool, n_uni: int, philox=(0, 0),
                   bonus_uniform_last: bool = False):
        """(lantern_accept_cfg, workspace bytes) for a launch shape; the struct is cached per shape and only its per-call
        fields are refreshed."""
        ckey = (B, T, L, D, logits.dtype, logits.stride(0), logits.stride(1), shared, n_uni)
        cached = self._cfg_cache.get(ckey)
        if cached is None:
            cfg = self._cfg(B, T, L, D, logits, shared, n_uni, philox)
            if len(self._cfg_cache) > 32:
                self._cfg_cache.clear()
            cached = self._cfg_cache[ckey] = (cfg, int(self.lib.lantern_accept_workspace_bytes(C.byref(cfg))))
        cfg, need = cached
        cfg.philox_seed, cfg.philox_step = int(philox[0]) & (2**64 - 1), int(philox[1]) & (2**64 - 1)
        cfg.bonus_uniform_last = int(bonus_uniform_last)
        return cfg, need

    def workspace(self, need: int, dev) -> torch.Tensor:
        if self._work is None or self._work.numel() < need or self._work.device != dev:
            self._work = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
        return self._work

    def step(self, logits_cond: torch.Tensor, logits_uncond: Optional[torch.Tensor], tree_tokens: torch.Tensor,
             retrieve: Optional[torch.Tensor] = None, *, row_kinds: Optional[torch.Tensor] = None,
             uniforms: Optional[torch.Tensor] = None, philox: Tuple[int, int] = (0, 0),
             node_q: Optional[torch.Tensor] = None, draft_op: Optional[torch.Tensor] = None,
             sib_tokens: Optional[torch.Tensor] = None, want_sample_p: bool = False,
             phases: Optional[int] = None, bonus_uniform_last: bool = False) -> VerifyResult:
        """logits_*: [B, T, V] (fp32/bf16/fp16, last dim contiguous); tree_tokens: [B, T] int32;
        retrieve: [B, L, D] or [L, D] int32 (defaults to the static tree's).  Asynchronous."""
        if phases is None:
            phases = _SCHEDULES[self.schedule]
        if logits_cond.dim() != 3 or logits_cond.stride(2) != 1:
            raise ValueError("logits must be [B, T, V] with a contiguous last dimension")
        if logits_uncond is not None and (logits_uncond.shape != logits_cond.shape
                                          or logits_uncond.stride() != logits_cond.stride()
                                          or logits_uncond.dtype != logits_cond.dtype):
            raise ValueError("logits_uncond must match logits_cond in shape, strides and dtype")
        B, T, V = logits_cond.shape
        if V != self.family.vocab:
            raise ValueError(f"vocab {V} != family vocab {self.family.vocab}")
        if retrieve is None:
            if self.static_tree is None:
                raise ValueError("retrieve indices are required for dynamic trees")
            retrieve = self.static_tree.retrieve
        shared = retrieve.dim() == 2
        L, D = retrieve.shape[-2:]
        for name, t in (("tree_tokens", tree_tokens), ("retrieve", retrieve)):
            if t.dtype != torch.int32 or not t.is_contiguous() or not t.is_cuda:
                raise ValueError(f"{name} must be a contiguous int32 CUDA tensor")
        n_uni = 0
        if uniforms is not None:
            if uniforms.dtype != torch.float32 or not uniforms.is_contiguous() or uniforms.shape[0] != B:
                raise ValueError("uniforms must be contiguous fp32 [B, n]")
            n_uni = uniforms.shape[1]
        cfg, need = self.cached_cfg(B, T, L, D, logits_cond, shared, n_uni, philox, bonus_uniform_last)
        ain = _abi.AcceptIn()
        ain.logits_cond, ain.logits_uncond = _ptr(logits_cond), _ptr(logits_uncond)
        ain.tree_tokens, ain.retrieve = _ptr(tree_tokens), _ptr(retrieve)
        ain.row_kinds = _ptr(row_kinds)
        ain.nbr_table = _ptr(self.nbr_table)
        ain.uniforms = _ptr(uniforms)
        if self.static_tree is not None:
            st = self.static_tree
            if node_q is None or draft_op is None:
                raise ValueError("static trees need node_q [B,T] and draft_op [B,R,V]")
            if draft_op.dtype != torch.float32 or not draft_op.is_contiguous() or draft_op.shape != (B, st.n_q_rows, V):
                raise ValueError(f"draft_op must be contiguous fp32 [B, {st.n_q_rows}, {V}]")
            if node_q.dtype != torch.float32 or not node_q.is_contiguous() or node_q.shape != (B, T):
                raise ValueError("node_q must be contiguous fp32 [B, T]")
            stoks = tree_tokens if sib_tokens is None else sib_tokens
            ain.node_q, ain.draft_op = _ptr(node_q), _ptr(draft_op)
            ain.node_qrow, ain.sib_off, ain.sib_idx = _ptr(st.node_qrow), _ptr(st.sib_off), _ptr(st.sib_idx)
            ain.sib_tokens, ain.sib_tokens_stride = _ptr(stoks), stoks.stride(0)
        dev = logits_cond.device
        ints = torch.empty(B * (5 + 2 * D), dtype=torch.int32, device=dev)
        res = VerifyResult(ints[0:B], ints[B:2 * B], ints[2 * B:3 * B],
                           ints[5 * B:5 * B + B * D].view(B, D), ints[5 * B + B * D:].view(B, D),
                           ints[3 * B:4 * B], ints[4 * B:5 * B])
        res.ints = ints      # [accept_length | best_candidate | token | n_draws | flags] x B, then the two [B, D] tables
        if want_sample_p:
            res.sample_p = torch.empty(B, V, dtype=torch.float32, device=dev)
        aout = _abi.AcceptOut()
        aout.accept_length, aout.best_candidate, aout.token = _ptr(res.accept_length), _ptr(res.best_candidate), _ptr(res.token)
        aout.path_tokens, aout.select_indices = _ptr(res.path_tokens), _ptr(res.select_indices)
        aout.n_draws, aout.flags, aout.sample_p = _ptr(res.n_draws), _ptr(res.flags), _ptr(res.sample_p)
        self.workspace(need, dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _abi.check(self.lib.lantern_accept_phases(C.byref(cfg), C.byref(ain), C.byref(aout), self._work.data_ptr(),
                                                  self._work.numel(), stream, phases))
        res._keepalive = (logits_cond, logits_uncond, tree_tokens, retrieve, row_kinds, uniforms, node_q, draft_op)
        return res


def _greedy(self, logits_cond: torch.Tensor, logits_uncond: Optional[torch.Tensor], tree_tokens: torch.Tensor,
            retrieve: torch.Tensor, want_row: bool = True) -> VerifyResult:
    """Greedy verification (temperature 0): ``lantern_accept_greedy``.  Same tensor layout as ``step``; the warp
    knobs are ignored, ``lantern`` selects the relaxed variant.  ``sample_p`` of the result is the CFG-mixed logits row
    of the last accepted node (the reference's third return value) and ``token`` its argmax."""
    if logits_cond.dim() != 3 or logits_cond.stride(2) != 1:
        raise ValueError("logits must be [B, T, V] with a contiguous last dimension")
    if logits_uncond is not None and (logits_uncond.shape != logits_cond.shape
                                      or logits_uncond.stride() != logits_cond.stride()
                                      or logits_uncond.dtype != logits_cond.dtype):
        raise ValueError("logits_uncond must match logits_cond in shape, strides and dtype")
    B, T, V = logits_cond.shape
    shared = retrieve.dim() == 2
    L, D = retrieve.shape[-2:]
    for name, t in (("tree_tokens", tree_tokens), ("retrieve", retrieve)):
        if t.dtype != torch.int32 or not t.is_contiguous() or not t.is_cuda:
            raise ValueError(f"{name} must be a contiguous int32 CUDA tensor")
    cfg = self._cfg(B, T, L, D, logits_cond, shared, 0, (0, 0))
    ain = _abi.AcceptIn()
    ain.logits_cond, ain.logits_uncond = _ptr(logits_cond), _ptr(logits_uncond)
    ain.tree_tokens, ain.retrieve = _ptr(tree_tokens), _ptr(retrieve)
    ain.nbr_table = _ptr(self.nbr_table)
    dev = logits_cond.device
    ints = torch.empty(B * (5 + 2 * D), dtype=torch.int32, device=dev)
    res = VerifyResult(ints[0:B], ints[B:2 * B], ints[2 * B:3 * B],
                       ints[5 * B:5 * B + B * D].view(B, D), ints[5 * B + B * D:].view(B, D),
                       ints[3 * B:4 * B], ints[4 * B:5 * B])
    res.ints = ints
    if want_row:
        res.sample_p = torch.empty(B, V, dtype=torch.float32, device=dev)
    aout = _abi.AcceptOut()
    aout.accept_length, aout.best_candidate, aout.token = _ptr(res.accept_length), _ptr(res.best_candidate), _ptr(res.token)
    aout.path_tokens, aout.select_indices = _ptr(res.path_tokens), _ptr(res.select_indices)
    aout.n_draws, aout.flags, aout.sample_p = _ptr(res.n_draws), _ptr(res.flags), _ptr(res.sample_p)
    work = torch.empty(max(int(self.lib.lantern_accept_greedy_workspace_bytes(C.byref(cfg))), 4), dtype=torch.uint8,
                       device=dev)
    _abi.check(self.lib.lantern_accept_greedy(C.byref(cfg), C.byref(ain), C.byref(aout), work.data_ptr(), work.numel(),
                                              torch.cuda.current_stream(dev).cuda_stream))
    res._keepalive = (logits_cond, logits_uncond, tree_tokens, retrieve, work)
    return res


Verifier.greedy = _greedy


def sample_tokens(probs: torch.Tensor, uniforms: torch.Tensor) -> torch.Tensor:
    """Inverse-CDF draw per row of ``probs`` [n, V] fp32 (the build's ``torch.multinomial(p, 1)``)."""
    lib = _abi.load()
    if probs.dim() != 2 or probs.dtype != torch.float32 or probs.stride(1) != 1 or not probs.is_cuda:
        raise ValueError("probs must be a CUDA fp32 [n, V] tensor with contiguous rows")
    n, V = probs.shape
    u = uniforms.to(device=probs.device, dtype=torch.float32).contiguous()
    out = torch.empty(n, dtype=torch.int32, device=probs.device)
    _abi.check(lib.lantern_sample_tokens(probs.data_ptr(), probs.stride(0), n, V, u.data_ptr(), out.data_ptr(),
                                         torch.cuda.current_stream(probs.device).cuda_stream))
    return out


def philox_uniforms(seed: int, step: int, item: int, n: int) -> np.ndarray:
    """Host copy of the device Philox stream (same words the kernel draws)."""
    lib = _abi.load()
    out = np.empty(n, dtype=np.float32)
    lib.lantern_philox_uniforms(seed & (2**64 - 1), step & (2**64 - 1), item, n, out.ctypes.data)
    return out
