"""Static-tree drafter sampling — SURVEY.md 8(f) row N3 (``Model.sample``, models/drafters/cnets_llamagen.py:924-940).

``sample(logits, logits_processor, k)`` keeps the reference's return triple ``(sampled_indices [R,k],
sampled_probs [R,k], probabilities [R,V])``; the draw uses the device Philox stream (``seed``, ``step``) instead of
``torch.multinomial`` (exponential race: same law, reproducible).  ``logits`` may be the CFG-doubled ``[2, R, V]``
drafter output, in which case the CFG mix is fused (``cfg_scale``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _abi, verify
from .posterior import _warp_knobs


def sample(logits: torch.Tensor, logits_processor, k: int = 10, *, uncond: Optional[torch.Tensor] = None,
           cfg_scale: float = 1.0, family: Optional[verify.FamilySpec] = None, seed: int = 0, step: int = 0
           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    lib = _abi.load()
    if logits.dim() != 2 or logits.stride(1) != 1 or not logits.is_cuda:
        raise ValueError("logits must be a CUDA [R, V] tensor with contiguous rows")
    R, V = logits.shape
    fam = family or verify.vanilla(V)
    t, p, tk = _warp_knobs(logits_processor)
    cfg = _abi.AcceptCfg()
    cfg.n_items, cfg.n_rows, cfg.n_paths, cfg.depth = 1, R, 1, 1
    cfg.vocab, cfg.col0, cfg.ncols = V, fam.col0, fam.ncols
    cfg.logits_dtype = verify._DTYPES[logits.dtype]
    cfg.item_stride, cfg.row_stride = R * logits.stride(0), logits.stride(0)
    cfg.family = fam.family_id
    cfg.cfg_scale, cfg.temperature, cfg.top_p, cfg.top_k = cfg_scale, t, p, tk
    cfg.philox_seed, cfg.philox_step = seed & (2**64 - 1), step & (2**64 - 1)
    ain = _abi.AcceptIn()
    ain.logits_cond = logits.data_ptr()
    if uncond is not None:
        if uncond.shape != logits.shape or uncond.stride() != logits.stride() or uncond.dtype != logits.dtype:
            raise ValueError("uncond must match logits")
        ain.logits_uncond = uncond.data_ptr()
    dev = logits.device
    probs = torch.empty(R, V, dtype=torch.float32, device=dev)
    idx = torch.empty(R, k, dtype=torch.int32, device=dev)
    cond = torch.empty(R, k, dtype=torch.float32, device=dev)
    need = lib.lantern_accept_workspace_bytes(C.byref(cfg))
    work = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    _abi.check(lib.lantern_draft_sample(C.byref(cfg), C.byref(ain), k, probs.data_ptr(), idx.data_ptr(), cond.data_ptr(),
                                        work.data_ptr(), work.numel(), torch.cuda.current_stream(dev).cuda_stream))
    return idx.long(), cond, probs
