"""Reference-compatible call surface of the verification step.

Same names, argument order and return conventions as the reference:

* module functions of ``models/drafters/utils.py``: ``prepare_logits_processor`` (:36), ``tree_decoding`` (:309),
  ``evaluate_posterior`` (:333), ``update_inference_inputs`` (:413);
* ``VerifyMixin`` — the methods of ``EaModel`` in ``models/ea_model_llamagen.py`` / ``ea_model_anole.py``
  (``tree_decoding`` :908, ``evaluate_posterior`` :709, ``evaluate_posterior_v1`` :464,
  ``update_inference_inputs`` :935);
* ``LuminaVerifyMixin`` — the methods of ``EaLumina_mGPT`` (``ea_model_lumina_mgpt.py`` :556, :610, :731).

``patch_reference(cls)`` installs the mixin methods on a reference class so its ``generate()`` loop runs unmodified.
All arithmetic happens in the CUDA library; this file only marshals tensors.  Two input forms are accepted:

* the reference's gathered ``logits [L, D, V]`` (already CFG-mixed / masked by its own ``tree_decoding``), or
* a ``TreeLogits`` handle returned by the ``tree_decoding`` shims here, which keeps the raw ``[2, T, V]`` logits so
  the CFG mix, masking, top-k and the leaf-path gather are fused into the kernel (nothing of size [L, D, V] is built).

Uniforms: ``rng="python"`` replays the reference's stream exactly — the walk consumes ``random.random()`` values in
the reference's order and the module RNG is advanced by exactly the number of draws the reference would have made;
``rng="philox"`` uses the device Philox stream (no host round trip for the uniforms).
"""
from __future__ import annotations

import random
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi, verify
from .verify import FamilySpec, StaticTree, Verifier

TOPK = 10


# ------------------------------------------------------------------------------------------------
# prepare_logits_processor
# ------------------------------------------------------------------------------------------------
class LogitsWarp(list):
    """What ``prepare_logits_processor`` returns: behaves like the HF ``LogitsProcessorList`` the reference builds
    (callable as ``proc(input_ids, scores)``, empty when temperature <= 1e-5) and carries the knobs the fused
    kernel needs."""

    def __init__(self, temperature: float = 0.0, top_p: float = 0.0, top_k: int = 0):
        super().__init__()
        self.temperature, self.top_p, self.top_k = float(temperature), float(top_p), int(top_k)
        if temperature > 1e-5:
            if temperature != 1.0:
                self.append(("temperature", temperature))
            if 1e-8 <= top_p < 1.0:
                self.append(("top_p", top_p))
            if top_k > 0:
                self.append(("top_k", top_k))

    def __call__(self, input_ids, scores: torch.Tensor) -> torch.Tensor:
        for name, val in self:
            if name == "temperature":
                scores = scores / val
            elif name == "top_p":
                srt, idx = torch.sort(scores, descending=False, stable=True)
                cum = srt.softmax(dim=-1).cumsum(dim=-1)
                remove = cum <= (1 - val)
                remove[..., -1:] = False
                scores = scores.masked_fill(remove.scatter(-1, idx, remove), -float("inf"))
            else:
                k = min(val, scores.size(-1))
                kth = torch.topk(scores, k)[0][..., -1, None]
                scores = scores.masked_fill(scores < kth, -float("inf"))
        return scores


def prepare_logits_processor(temperature: float = 0.0, repetition_penalty: float = 0.0, top_p: float = 0.0,
                             top_k: int = 0) -> LogitsWarp:
    """drafters/utils.py:36-52 (the repetition penalty is never enabled by the reference callers)."""
    if repetition_penalty > 1.0:
        raise NotImplementedError("repetition_penalty is not part of the verification path")
    return LogitsWarp(temperature, top_p, top_k)


def _warp_knobs(proc) -> Tuple[float, float, int]:
    if proc is None:
        raise ValueError("logits_processor is None: greedy decoding, use evaluate_posterior_greedy")
    if isinstance(proc, LogitsWarp):
        return proc.temperature, proc.top_p if 1e-8 <= proc.top_p < 1.0 else 1.0, proc.top_k
    t, p, k = 1.0, 1.0, 0          # a HF LogitsProcessorList built by the reference itself
    for w in proc:
        n = type(w).__name__
        if n == "TemperatureLogitsWarper":
            t = float(w.temperature)
        elif n == "TopPLogitsWarper":
            p = float(w.top_p)
        elif n == "TopKLogitsWarper":
            k = int(w.top_k)
        else:
            raise NotImplementedError(f"unsupported logits processor {n}")
    return t, p, k


# ------------------------------------------------------------------------------------------------
# Fused-path handle
# ------------------------------------------------------------------------------------------------
@dataclass
class TreeLogits:
    """Returned by the ``tree_decoding`` shims in place of the gathered ``[L, D, V]`` tensor."""
    cond: torch.Tensor                   # [1, T, V]
    uncond: Optional[torch.Tensor]       # [1, T, V] or None
    cfg_scale: float
    retrieve_indices: torch.Tensor       # [L, D] int64 (as the reference holds it)
    row_kinds: Optional[torch.Tensor] = None   # [1, T] uint8 (Lumina)
    top_k: int = 0                       # Lumina: InterleavedTopKLogitsWarper applied in tree_decoding

    @property
    def device(self):
        return self.cond.device

    @property
    def shape(self):
        L, D = self.retrieve_indices.shape
        return (L, D, self.cond.shape[-1])


def lumina_row_kinds(position_ids_plus1: torch.Tensor, image_start_idx: int, h: int = 48, w: int = 48) -> torch.Tensor:
    """Row classes of ``MultiModalLogitsProcessor`` (ea_model_lumina_mgpt.py:45-86): index arithmetic only."""
    n = position_ids_plus1.long() - (image_start_idx + 1 + 2)
    kinds = torch.zeros_like(n, dtype=torch.uint8)
    kinds[((n + 1) % (w + 1)) == 0] = _abi.ROW_NEWLINE
    kinds[(n + 1) == (w + 1) * h + 1] = _abi.ROW_EOI
    return kinds


def lumina_process_rows(rows: torch.Tensor, kinds: torch.Tensor, fam: FamilySpec, top_k: int) -> torch.Tensor:
    """``MultiModalLogitsProcessor`` + ``InterleavedTopKLogitsWarper`` (ea_model_lumina_mgpt.py:45-112) on CFG-mixed
    rows [T, V] - the unfused (``lantern_fused=False``) path of a stand-in that carries no processor objects."""
    out = torch.full_like(rows, -float("inf"))
    img = kinds == _abi.ROW_IMAGE
    out[img, fam.col0:fam.col0 + fam.ncols] = rows[img, fam.col0:fam.col0 + fam.ncols]
    out[kinds == _abi.ROW_NEWLINE, fam.newline_token] = 0
    out[kinds == _abi.ROW_EOI, fam.eoi_token] = 0
    if top_k > 0:
        kth = torch.topk(out, min(top_k, out.shape[-1]))[0][..., -1, None]
        out = out.masked_fill(out < kth, -float("inf"))
    return out


# ------------------------------------------------------------------------------------------------
# Core marshalling
# ------------------------------------------------------------------------------------------------
_verifiers = {}
_static_cache = {}


def _get_verifier(fam: FamilySpec, temp, top_p, top_k, cfg_scale, lantern, k, delta, table, static: Optional[StaticTree],
                  device) -> Verifier:
    key = (fam, float(temp), float(top_p), int(top_k), float(cfg_scale), bool(lantern), int(k), float(delta),
           id(table), id(static), str(device))
    v = _verifiers.get(key)
    if v is None:
        if len(_verifiers) > 64:
            _verifiers.clear()
        v = Verifier(fam, temperature=temp, top_k=top_k, top_p=top_p, cfg_scale=cfg_scale, lantern=lantern,
                     lantern_k=k, lantern_delta=delta, nbr_table=table, static_tree=static, device=device)
        v._table_ref = table
        _verifiers[key] = v
    return v


_table_cache = {}


def device_table(nearest_latents, device, cols: int) -> torch.Tensor:
    """The reference keeps ``nearest_latents`` as a uint16 / int64 numpy ``[N, N-1]`` array in host memory and slices
    it per candidate (ea_model_llamagen.py:143,744).  Here the first ``cols`` columns live on the device as int32."""
    key = (id(nearest_latents), str(device), cols)
    t = _table_cache.get(key)
    if t is None:
        if isinstance(nearest_latents, torch.Tensor):
            t = nearest_latents[:, :cols].to(device=device, dtype=torch.int32).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(nearest_latents[:, :cols]).astype(np.int32)).to(device)
        if len(_table_cache) > 8:
            _table_cache.clear()
        _table_cache[key] = (t, nearest_latents)     # keep the source alive so id() stays unique
        return t
    return t[0]


def _draw_python_uniforms(n: int):
    state = random.getstate()
    u = [random.random() for _ in range(n)]
    return state, u


class _HostIO:
    """Staging for the batch-1 drop-in call, one per (device, host thread): pinned host buffers for the uniforms going
    in and the five result integers coming out, so that one call costs two async copies and a single stream
    synchronisation.  A call is synchronous, so a thread never has two calls in flight on one buffer; different
    threads (one patched model each) get their own."""

    def __init__(self, device):
        self.u_host = torch.empty(4096, dtype=torch.float32).pin_memory()
        self.u_np = self.u_host.numpy()
        self.u_dev = torch.empty(4096, dtype=torch.float32, device=device)
        self.r_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.r_np = self.r_host.numpy()
        # context of lantern_posterior_call (the one-FFI-call form of a batch-1 evaluate_posterior)
        self.call = None
        self.call_u = None
        self.call_out = np.zeros(5 + 2 * 64, dtype=np.int32)

    MAX_ROWS, MAX_CELLS = 1024, 16384

    def call_handle(self):
        if self.call is None:
            import ctypes as C
            lib = _abi.load()
            h = C.c_void_p()
            _abi.check(lib.lantern_call_create(self.MAX_ROWS, self.MAX_CELLS, C.byref(h)))
            self.call = h
            ptr = lib.lantern_call_uniforms(h)
            self.call_u = np.ctypeslib.as_array((C.c_float * (self.MAX_ROWS + 1)).from_address(ptr))
        return self.call


_host_io = {}


def _io(device) -> _HostIO:
    key = (str(device), threading.get_ident())
    io = _host_io.get(key)
    if io is None:
        io = _host_io[key] = _HostIO(device)
    return io


def _tree_inputs(logits, candidates: torch.Tensor, stream):
    """Kernel-side form of what the reference hands to evaluate_posterior: (cond [1,T,V], uncond, cfg_scale,
    tree_tokens [1,T] int32, retrieve [1,L,D] int32, row_kinds, T, Lumina top-k)."""
    fused = isinstance(logits, TreeLogits)
    device = logits.device
    cand = candidates if candidates.device == device else candidates.to(device)
    L, D = cand.shape
    if fused:
        cond, uncond, cfg_scale = logits.cond, logits.uncond, logits.cfg_scale
        ri64 = logits.retrieve_indices
        if ri64.device != device or ri64.dtype != torch.int64 or not ri64.is_contiguous():
            ri64 = ri64.to(device=device, dtype=torch.int64).contiguous()
        if cand.dtype != torch.int64 or not cand.is_contiguous():
            cand = cand.to(torch.int64).contiguous()
        T = cond.shape[1]
        ibuf = torch.empty(T + L * D, dtype=torch.int32, device=device)
        tokens, retrieve = ibuf[:T].view(1, T), ibuf[T:].view(1, L, D)
        _abi.check(_abi.load().lantern_tree_from_candidates(cand.data_ptr(), ri64.data_ptr(), L, D, T,
                                                            tokens.data_ptr(), retrieve.data_ptr(), stream.cuda_stream))
        return cond, uncond, cfg_scale, tokens, retrieve, logits.row_kinds, T, logits.top_k
    V = logits.shape[-1]
    cond = logits.reshape(1, L * D, V)
    if cond.stride(2) != 1:
        cond = cond.contiguous()
    T = L * D
    tokens = cand.reshape(-1).to(torch.int32)
    ids = torch.arange(T, device=device, dtype=torch.int32).view(L, D)
    retrieve = torch.where(cand >= 0, ids, torch.full_like(ids, -1)).contiguous()[None]
    tokens = torch.where(tokens >= 0, tokens, torch.zeros_like(tokens)).view(1, T).contiguous()
    return cond, None, 1.0, tokens, retrieve, None, T, None


def _verify_greedy(fam: FamilySpec, logits, candidates: torch.Tensor, lantern=False, lantern_k=1000, lantern_delta=0.1,
                   nearest_latents=None):
    """Greedy branches (``logits_processor is None``): drafters/utils.py:356-369, ea_model_anole.py:789-902.  Returns the
    reference's greedy triple: (best_candidate 0-d int64 on the device, accept_length 0-d int64 on the device,
    logits[best, accept_length] [V])."""
    device = logits.device
    stream = torch.cuda.current_stream(device)
    cond, uncond, cfg_scale, tokens, retrieve, _, _, _ = _tree_inputs(logits, candidates, stream)
    table, k = None, int(lantern_k)
    if lantern:
        k = min(k, fam.ncols - 1)
        table = device_table(nearest_latents, device, min(k + 1, fam.ncols - 1))
    ver = _get_verifier(fam, 1.0, 1.0, 0, cfg_scale, lantern, k, lantern_delta, table, None, device)
    res = ver.greedy(cond, uncond, tokens, retrieve)
    row = res.sample_p[0]
    row._lantern_argmax = res.token[0]     # device scalar: sample_bonus_token(do_sample=False) reuses it
    return res.best_candidate[0].long(), res.accept_length[0].long(), row


def _verify_one_call(fam: FamilySpec, logits: "TreeLogits", candidates: torch.Tensor, temp, top_p, top_k, lantern,
                     lantern_k, lantern_delta, nearest_latents, static_inputs, rng, philox, want_sample_p):
    """The fused-handle form of ``_verify`` behind ONE FFI call (``lantern_posterior_call``): uniforms upload, tree inputs
    from ``candidates`` / ``retrieve_indices``, the fused step, the result read-back and the stream synchronisation all
    happen inside the library.  Same results as the tensor-by-tensor route below."""
    import ctypes as C
    device = logits.device
    cond, uncond = logits.cond, logits.uncond
    cand = candidates
    if cand.device != device or cand.dtype != torch.int64 or not cand.is_contiguous():
        cand = cand.to(device=device, dtype=torch.int64).contiguous()
    ri64 = logits.retrieve_indices
    if ri64.device != device or ri64.dtype != torch.int64 or not ri64.is_contiguous():
        ri64 = ri64.to(device=device, dtype=torch.int64).contiguous()
    L, D = cand.shape
    T = cond.shape[1]
    if fam.family_id == _abi.FAMILY_LUMINA and logits.top_k is not None:
        top_k = logits.top_k
    table, k = None, int(lantern_k)
    if lantern:
        k = min(k, fam.ncols - 1)
        table = device_table(nearest_latents, device, min(k + 1, fam.ncols - 1))
    static = static_inputs[0] if static_inputs is not None else None
    ver = _get_verifier(fam, temp, top_p, top_k, logits.cfg_scale, lantern, k, lantern_delta, table, static, device)
    io = _io(device)
    handle = io.call_handle()
    n_u, state = 0, None
    if rng == "python":
        state, u = _draw_python_uniforms(T)
        io.call_u[:T] = u                                   # float64 -> float32, round to nearest like torch
        io.call_u[T] = float(torch.rand(()).item())         # bonus draw (see _verify for the RNG-state note)
        n_u = T + 1
    cfg, need = ver.cached_cfg(1, T, L, D, cond, False, n_u, philox, rng == "python")
    ain = _abi.AcceptIn()
    ain.logits_cond, ain.logits_uncond = cond.data_ptr(), (uncond.data_ptr() if uncond is not None else None)
    ain.row_kinds = logits.row_kinds.data_ptr() if logits.row_kinds is not None else None
    ain.nbr_table = ver.nbr_table.data_ptr() if ver.nbr_table is not None else None
    keep = None
    if static_inputs is not None:
        st, node_q, draft_op, sib_tokens = static_inputs
        ain.node_q, ain.draft_op = node_q.data_ptr(), draft_op.data_ptr()
        ain.node_qrow, ain.sib_off, ain.sib_idx = st.node_qrow.data_ptr(), st.sib_off.data_ptr(), st.sib_idx.data_ptr()
        ain.sib_tokens, ain.sib_tokens_stride = sib_tokens.data_ptr(), sib_tokens.stride(0)
        keep = static_inputs
    sample_p = torch.empty(cond.shape[-1], dtype=torch.float32, device=device) if want_sample_p else None
    work = ver.workspace(need, device)
    out = io.call_out
    _abi.check(ver.lib.lantern_posterior_call(handle, C.byref(cfg), C.byref(ain), cand.data_ptr(), ri64.data_ptr(), n_u,
                                              sample_p.data_ptr() if sample_p is not None else None, work.data_ptr(),
                                              work.numel(), out.ctypes.data,
                                              torch.cuda.current_stream(device).cuda_stream))
    a, best, token, draws = int(out[0]), int(out[1]), int(out[2]), int(out[3])
    if state is not None:                                    # advance the module RNG exactly like the reference
        random.setstate(state)
        for _ in range(draws - 1):
            random.random()
    if sample_p is None:
        sample_p = torch.empty(0, device=device)
    sample_p._lantern_token = token
    del keep
    return torch.tensor(best), a, sample_p


def _verify(fam: FamilySpec, logits, candidates: torch.Tensor, *, temp=1.0, top_p=1.0, top_k=0, lantern=False,
            lantern_k=1000, lantern_delta=0.1, nearest_latents=None, static_inputs=None, rng="python",
            philox=(0, 0), want_sample_p=True):
    """Returns (best_candidate 0-d int64 CPU tensor, accept_length int, sample_p [V] with ``_lantern_token``)."""
    device = logits.device
    if (isinstance(logits, TreeLogits) and logits.cond.shape[0] == 1 and logits.cond.shape[1] <= _HostIO.MAX_ROWS
            and candidates.numel() <= _HostIO.MAX_CELLS and candidates.shape[1] <= 64):
        return _verify_one_call(fam, logits, candidates, temp, top_p, top_k, lantern, lantern_k, lantern_delta,
                                nearest_latents, static_inputs, rng, philox, want_sample_p)
    stream = torch.cuda.current_stream(device)
    cond, uncond, cfg_scale, tokens, retrieve, kinds, T, lumina_top_k = _tree_inputs(logits, candidates, stream)
    if fam.family_id == _abi.FAMILY_LUMINA and lumina_top_k is not None:
        top_k = lumina_top_k
    table, k = None, int(lantern_k)
    if lantern:
        n_codes = fam.ncols
        k = min(k, n_codes - 1)
        table = device_table(nearest_latents, device, min(k + 1, n_codes - 1))
    static = None
    kw = {}
    if static_inputs is not None:
        static, node_q, draft_op, sib_tokens = static_inputs
        kw = dict(node_q=node_q, draft_op=draft_op, sib_tokens=sib_tokens)
    ver = _get_verifier(fam, temp, top_p, top_k, cfg_scale, lantern, k, lantern_delta, table, static, device)
    io = _io(device)
    uniforms = None
    state = None
    if rng == "python":
        state, u = _draw_python_uniforms(T)
        n_u = T + 1
        if n_u > io.u_np.shape[0]:
            raise ValueError(f"tree of {T} nodes exceeds the uniform staging buffer")
        io.u_np[:T] = u                                   # float64 -> float32, round to nearest like torch
        # The bonus-token draw comes from torch's CPU generator.  The reference's torch.multinomial advances the CUDA
        # generator instead (by an amount that is an implementation detail of ATen), so after a seeded run the torch RNG
        # states differ; the python `random` stream - the one that decides acceptance - is replayed exactly.
        io.u_np[T] = float(torch.rand(()).item())
        uniforms = io.u_dev[:n_u].view(1, n_u)
        uniforms.copy_(io.u_host[:n_u].view(1, n_u), non_blocking=True)
    res = ver.step(cond, uncond, tokens, retrieve, row_kinds=kinds, uniforms=uniforms,
                   philox=philox, want_sample_p=want_sample_p, bonus_uniform_last=(rng == "python"), **kw)
    # accept_length, best_candidate, token, n_draws, flags are the first five ints of one buffer (B = 1)
    io.r_host[:5].copy_(res.ints[:5], non_blocking=True)
    stream.synchronize()
    a, best, token, draws = (int(x) for x in io.r_np[:4])
    if state is not None:                                # advance the module RNG exactly like the reference
        random.setstate(state)
        for _ in range(draws - 1):
            random.random()
    sample_p = res.sample_p[0] if want_sample_p else torch.empty(0, device=device)
    sample_p._lantern_token = token
    return torch.tensor(best), a, sample_p


# ------------------------------------------------------------------------------------------------
# module-level functions (vanilla EAGLE helpers, drafters/utils.py)
# ------------------------------------------------------------------------------------------------
def evaluate_posterior(logits, candidates, logits_processor, rng: str = "python"):
    """drafters/utils.py:333-410."""
    if logits_processor is None or len(logits_processor) == 0 and not getattr(logits_processor, "temperature", 1) > 1e-5:
        if isinstance(logits, TreeLogits):
            return _verify_greedy(verify.vanilla(logits.shape[-1]), logits, candidates)
        return evaluate_posterior_greedy(logits, candidates)
    t, p, k = _warp_knobs(logits_processor)
    fam = verify.vanilla(logits.shape[-1])
    return _verify(fam, logits, candidates, temp=t, top_p=p, top_k=k, rng=rng)


def evaluate_posterior_greedy(logits: torch.Tensor, candidates: torch.Tensor):
    """Greedy branch (drafters/utils.py:356-369): argmax match, cumprod, longest prefix — index arithmetic on [L, D]."""
    top = torch.argmax(logits[:, :-1], dim=-1)
    mask = (candidates[:, 1:].to(logits.device) == top).int()
    acc = torch.cumprod(mask, dim=1).sum(dim=1)
    a = acc.max()
    best = torch.tensor(0, dtype=torch.long, device=candidates.device) if a == 0 else torch.argmax(acc).to(torch.long)
    return best, a, logits[best, a]


def tree_decoding(model, tree_candidates, past_key_values, tree_position_ids, input_ids, retrieve_indices,
                  fused: bool = True):
    """drafters/utils.py:309-327.  The target forward is the caller's model; only the gather is replaced."""
    position_ids = tree_position_ids + input_ids.shape[1]
    outputs, tree_logits, hidden_state = model(tree_candidates, output_orig=True, past_key_values=past_key_values,
                                               position_ids=position_ids)
    if fused:
        return TreeLogits(tree_logits[:1], None, 1.0, retrieve_indices), hidden_state, outputs
    return tree_logits[0, retrieve_indices], hidden_state, outputs


def sample_bonus_token(sample_p: torch.Tensor, do_sample: bool = True) -> torch.Tensor:
    """``torch.multinomial(prob, 1)`` / ``argmax`` of update_inference_inputs (ea_model_llamagen.py:976-982); returns
    ``[1, 1]`` int64.  Reuses the token the fused step already drew when ``sample_p`` came from it."""
    tok = getattr(sample_p, "_lantern_token", None)
    if not do_sample:
        am = getattr(sample_p, "_lantern_argmax", None)
        return (torch.argmax(sample_p) if am is None else am.long())[None, None]
    if tok is None:
        u = torch.rand(1)
        tok = int(verify.sample_tokens(sample_p.float().view(1, -1), u)[0])
    return torch.tensor([[tok]], dtype=torch.long, device=sample_p.device)


def kv_compact(past_key_values_data_list: Sequence[torch.Tensor], select_indices: torch.Tensor, prev_len: int) -> int:
    """The per-slab ``dst.copy_(tgt)`` loop of update_inference_inputs (ea_model_llamagen.py:962-967) as one launch
    per device: slab ``[2*layers, batch, heads, S_max, head_dim]``."""
    import ctypes as C
    lib = _abi.load()
    n = int(select_indices.numel())
    for data in past_key_values_data_list:
        dev = data.device
        sel = select_indices
        if sel.device != dev or sel.dtype not in (torch.int64, torch.int32) or not sel.is_contiguous():
            sel = sel.to(device=dev, dtype=torch.int64).contiguous()
        cfg = _abi.KvCfg()
        cfg.n_slabs, cfg.elem_bytes = 1, data.element_size()
        cfg.n_outer = data.shape[0] * data.shape[1] * data.shape[2]
        cfg.outer_per_batch, cfg.n_batch = 1, 1          # one prompt: every outer slice uses select row 0
        cfg.s_max, cfg.head_dim, cfg.max_keep = data.shape[3], data.shape[4], n
        # batch-1 call: slab pointer and lengths travel by value, the index tensor is used as it is (no H2D copy)
        cfg.slab0, cfg.prev_len0, cfg.n_keep0 = data.data_ptr(), int(prev_len), n
        cfg.select_i64 = int(sel.dtype == torch.int64)
        _abi.check(lib.lantern_kv_compact(C.byref(cfg), None, sel.data_ptr(), None, None,
                                          torch.cuda.current_stream(dev).cuda_stream))
    return prev_len + n


def update_inference_inputs(input_ids, candidates, best_candidate, accept_length, retrieve_indices, logits_processor,
                            new_token, past_key_values_data_list, current_length_data, model, hidden_state_new,
                            sample_p):
    """drafters/utils.py:413-468."""
    prev_input_len = input_ids.shape[1]
    select_indices = retrieve_indices[best_candidate, : accept_length + 1] + prev_input_len
    input_ids = torch.cat([input_ids, candidates[None, best_candidate, : accept_length + 1].to(input_ids.device)], dim=-1)
    new_len = kv_compact(past_key_values_data_list, select_indices, prev_input_len)
    current_length_data.fill_(new_len)
    retrieve_hidden_state_new = hidden_state_new[:, retrieve_indices]
    accept_hidden_state_new = retrieve_hidden_state_new[:, best_candidate, : accept_length + 1]
    token = sample_bonus_token(sample_p, logits_processor is not None)
    draft_tokens, retrieve_indices, tree_mask, tree_position_ids = model.ea_layer.topK_genrate(
        accept_hidden_state_new, input_ids=torch.cat((input_ids, token.to(input_ids.device)), dim=1),
        head=model.base_model.lm_head, logits_processor=logits_processor)
    new_token += accept_length + 1
    return input_ids, draft_tokens, retrieve_indices, tree_mask, tree_position_ids, new_token, None, token


# ------------------------------------------------------------------------------------------------
# static-tree extras -> kernel inputs
# ------------------------------------------------------------------------------------------------
def _static_inputs(fused: bool, candidates, cart_candidates_prob, op, p_indices, tree_candidates, b_indices, device,
                   retrieve_indices=None):
    """Map ``evaluate_posterior_v1``'s extra arguments (ea_model_llamagen.py:464-477) onto the kernel's arrays."""
    L, D = candidates.shape
    key = (id(p_indices), id(b_indices), fused, L, D, str(device))
    cached = _static_cache.get(key)
    counts = [int(o.shape[0]) for o in op]
    offs = np.concatenate([[0], np.cumsum(counts)])
    if cached is None:
        if fused:
            ri = retrieve_indices.cpu().numpy()
            T = int(ri.max()) + 1
            node_of = lambda j, i: int(ri[j, i])
        else:
            T = L * D
            node_of = lambda j, i: j * D + i
        qrow = np.zeros(T, dtype=np.int32)
        sibs: List[List[int]] = [[] for _ in range(T)]
        for j in range(L):
            for i in range(1, D):
                n = node_of(j, i)
                if n < 0:
                    continue
                qrow[n] = int(offs[i - 1]) + int(p_indices[j][i])
                b = b_indices[j][i]
                sibs[n] = b.tolist() if isinstance(b, torch.Tensor) else list(b)
        sib_off = np.zeros(T + 1, dtype=np.int32)
        flat: List[int] = []
        for n in range(T):
            flat.extend(sibs[n])
            sib_off[n + 1] = len(flat)
        ri_dev = (retrieve_indices.to(device=device, dtype=torch.int32).contiguous() if fused else
                  torch.arange(T, dtype=torch.int32, device=device).view(L, D))
        st = StaticTree(retrieve=ri_dev, node_qrow=torch.from_numpy(qrow).to(device),
                        sib_off=torch.from_numpy(sib_off).to(device),
                        sib_idx=torch.tensor(flat if flat else [0], dtype=torch.int32, device=device),
                        n_q_rows=int(offs[-1]))
        if len(_static_cache) > 16:
            _static_cache.clear()
        _static_cache[key] = (st, p_indices, b_indices)
    else:
        st = cached[0]
    T = st.node_qrow.shape[0]
    cq = cart_candidates_prob.to(device=device, dtype=torch.float32)
    if fused:
        node_q = torch.ones(T, dtype=torch.float32, device=device)
        ri = retrieve_indices.to(device)
        m = ri >= 0
        node_q[ri[m]] = cq[m]
    else:
        node_q = cq.reshape(-1).contiguous()
    draft_op = torch.cat([o.to(device=device, dtype=torch.float32) for o in op], dim=0)[None].contiguous()
    sib_tokens = tree_candidates.to(device=device, dtype=torch.int32).contiguous().view(1, -1)
    return st, node_q.view(1, T), draft_op, sib_tokens


# ------------------------------------------------------------------------------------------------
# EaModel methods (LlamaGen / Anole)
# ------------------------------------------------------------------------------------------------
class VerifyMixin:
    """Drop-in methods for ``EaModel`` (LlamaGen / Anole).  ``self`` needs ``nearest_latents`` (lantern) and, for
    Anole, ``image_token_offset = 4``; ``lantern_family`` may name the family explicitly."""

    lantern_rng = "python"
    lantern_fused = True

    def _family_name(self) -> str:
        name = getattr(self, "lantern_family", None)
        if name is None:
            name = "anole" if getattr(self, "image_token_offset", 0) == 4 else "llamagen"
        return name

    def _family(self, vocab: int) -> FamilySpec:
        fam = verify.FAMILIES[self._family_name()]
        if fam.vocab == vocab:
            return fam
        ncols = getattr(self, "lantern_image_tokens", fam.ncols if fam.col0 else vocab)
        return fam.resized(ncols, vocab)

    def tree_decoding(self, tree_candidates, past_key_values, tree_position_ids, input_ids, retrieve_indices, cfg_scale,
                      attention_mask=None, input_position_diff=0):
        """ea_model_llamagen.py:907-932 / ea_model_anole.py:904-933 with the CFG mix, the non-image mask and the
        [L, D, V] gather left to the fused kernel."""
        position_ids = tree_position_ids + input_ids.shape[1]
        if self._family_name() == "anole" or input_position_diff:
            position_ids = position_ids.unsqueeze(0)
            position_ids = torch.cat([position_ids, position_ids - input_position_diff], dim=0)
        if attention_mask is not None:
            remaining = input_ids.shape[1] + tree_candidates.shape[1] - attention_mask.shape[1]
            attention_mask = torch.cat([attention_mask, attention_mask.new_ones((attention_mask.shape[0], remaining))], dim=1)
        outputs, tree_logits, hidden_state = self(input_ids=tree_candidates, output_orig=True,
                                                  past_key_values=past_key_values, position_ids=position_ids,
                                                  attention_mask=attention_mask)
        half = tree_logits.shape[0] // 2
        if not self.lantern_fused:
            mixed = tree_logits[half:] + (tree_logits[:half] - tree_logits[half:]) * cfg_scale
            return mixed[0, retrieve_indices], hidden_state, outputs
        handle = TreeLogits(tree_logits[:1], tree_logits[half:half + 1], float(cfg_scale), retrieve_indices)
        return handle, hidden_state, outputs

    def evaluate_posterior(self, logits, candidates, logits_processor=None, lantern=False, lantern_k=1000,
                           lantern_delta=0.1):
        """ea_model_llamagen.py:709-787 / ea_model_anole.py:709-788 (sampling branch)."""
        if logits_processor is None:
            # greedy branches (ea_model_anole.py:789-902); gathered logits without relaxation stay index arithmetic
            if lantern or isinstance(logits, TreeLogits):
                return _verify_greedy(self._family(logits.shape[-1]), logits, candidates, lantern=lantern,
                                      lantern_k=lantern_k, lantern_delta=lantern_delta,
                                      nearest_latents=getattr(self, "nearest_latents", None))
            return evaluate_posterior_greedy(logits, candidates)
        t, p, k = _warp_knobs(logits_processor)
        fam = self._family(logits.shape[-1])
        return _verify(fam, logits, candidates, temp=t, top_p=p, top_k=k, lantern=lantern, lantern_k=lantern_k,
                       lantern_delta=lantern_delta, nearest_latents=getattr(self, "nearest_latents", None),
                       rng=self.lantern_rng)

    def evaluate_posterior_v1(self, logits, candidates, logits_processor, cart_candidates_prob, op, p_indices,
                              tree_candidates, b_indices, lantern=False, lantern_k=1000, lantern_delta=0.1):
        """ea_model_llamagen.py:463-669 / ea_model_anole.py:464-669 (static tree, LANTERN++)."""
        if logits_processor is None:     # same greedy code as the dynamic method (ea_model_anole.py:478-595)
            return _verify_greedy(self._family(logits.shape[-1]), logits, candidates, lantern=lantern,
                                  lantern_k=lantern_k, lantern_delta=lantern_delta,
                                  nearest_latents=getattr(self, "nearest_latents", None))
        t, p, k = _warp_knobs(logits_processor)
        fam = self._family(logits.shape[-1])
        fused = isinstance(logits, TreeLogits)
        si = _static_inputs(fused, candidates, cart_candidates_prob, op, p_indices, tree_candidates, b_indices,
                            logits.device, logits.retrieve_indices if fused else None)
        return _verify(fam, logits, candidates, temp=t, top_p=p, top_k=k, lantern=lantern, lantern_k=lantern_k,
                       lantern_delta=lantern_delta, nearest_latents=getattr(self, "nearest_latents", None),
                       static_inputs=si, rng=self.lantern_rng)

    def update_inference_inputs(self, input_ids, candidates, best_candidate, accept_length, retrieve_indices,
                                logits_processor, new_token, past_key_values_data_list, current_length_data,
                                hidden_state_new, sample_p, cfg_scale, input_position_diff=None, attention_mask=None,
                                static_tree=False):
        """ea_model_llamagen.py:934-999 / ea_model_anole.py:935-1006."""
        prev_input_len = input_ids.shape[1]
        select_indices = retrieve_indices[best_candidate, : accept_length + 1] + prev_input_len
        input_ids = torch.cat([input_ids, candidates[None, best_candidate, : accept_length + 1]], dim=-1)
        new_len = kv_compact(past_key_values_data_list, select_indices, prev_input_len)
        current_length_data.fill_(new_len)
        retrieve_hidden_state_new = hidden_state_new[:, retrieve_indices]
        accept_hidden_state_new = retrieve_hidden_state_new[:, best_candidate, : accept_length + 1]
        token = sample_bonus_token(sample_p, logits_processor is not None)
        ea_input_ids = torch.cat((input_ids, token.to(input_ids.device)), dim=1).repeat(2, 1)
        extra = {}
        if input_position_diff is not None:
            extra = dict(input_position_diff=input_position_diff, attention_mask=attention_mask)
        if static_tree:
            tree_logits = self.ea_layer.topK_genrate_v1(accept_hidden_state_new, input_ids=ea_input_ids,
                                                        head=self.base_model.lm_head, logits_processor=logits_processor,
                                                        cfg_scale=cfg_scale, **extra)
            new_token += accept_length + 1
            return input_ids, tree_logits, new_token, None, token
        draft_tokens, retrieve_indices, tree_mask, tree_position_ids = self.ea_layer.topK_genrate(
            accept_hidden_state_new, input_ids=ea_input_ids, head=self.base_model.lm_head,
            logits_processor=logits_processor, cfg_scale=cfg_scale, **extra)
        new_token += accept_length + 1
        return input_ids, draft_tokens, retrieve_indices, tree_mask, tree_position_ids, new_token, None, token


# ------------------------------------------------------------------------------------------------
# EaLumina_mGPT methods
# ------------------------------------------------------------------------------------------------
class LuminaVerifyMixin:
    """Drop-in methods for ``EaLumina_mGPT``.  ``self`` needs ``cfg_mode``, ``cfg_scale``, ``eagle_version``,
    ``image_start_token_id_index``, ``nearest_latents`` and the image top-k (``lantern_image_top_k``, default 2000)."""

    lantern_rng = "python"
    lantern_fused = True
    lantern_image_top_k = 2000

    def _family(self, vocab: int) -> FamilySpec:
        fam = verify.LUMINA
        return fam if fam.vocab == vocab else fam.resized(getattr(self, "lantern_image_tokens", fam.ncols), vocab)

    def _image_top_k(self) -> int:
        """The k of the InterleavedTopKLogitsWarper the reference applies in tree_decoding: ``generate(top_k=...)``
        appends it to ``self.internal_logits_processors`` on every call (ea_model_lumina_mgpt.py:822-823) and
        tree_decoding reads entry [1] (:605).  ``lantern_image_top_k`` is only the fallback for stand-ins that carry
        no processor list."""
        procs = getattr(self, "internal_logits_processors", None)
        if procs is not None and len(procs) > 1 and hasattr(procs[1], "image_top_k"):
            return int(procs[1].image_top_k)
        return int(self.lantern_image_top_k)

    def tree_decoding(self, tree_candidates, attention_mask, past_key_values, tree_position_ids, input_ids,
                      retrieve_indices):
        """ea_model_lumina_mgpt.py:556-608: both CFG modes run the caller's target model; the CFG mix, the
        MultiModalLogitsProcessor row classes, the top-k and the gather are left to the fused kernel
        (``lantern_fused=False`` keeps the reference's own post-processing and returns the gathered [L, D, V])."""
        position_ids = tree_position_ids + input_ids.shape[1]
        if self.cfg_mode == "parallel":
            tc = torch.cat((tree_candidates, tree_candidates), dim=0)
            pos2 = torch.cat((position_ids[None], position_ids[None] - self.image_start_token_id_index), dim=0)
            _, tl, hs = self(input_ids=tc, attention_mask=attention_mask, output_orig=True,
                             past_key_values=past_key_values, position_ids=pos2)
            tree_logits, uncond_tree_logits = torch.split(tl, [1, 1])
            hidden_states, uncond_hidden_states = torch.split(hs, [1, 1])
        else:
            _, tree_logits, hidden_states = self(input_ids=tree_candidates, output_orig=True,
                                                 past_key_values=past_key_values["cond"], position_ids=position_ids)
            _, uncond_tree_logits, uncond_hidden_states = self(
                input_ids=tree_candidates, output_orig=True, past_key_values=past_key_values["uncond"],
                position_ids=position_ids - self.image_start_token_id_index)
        top_k = self._image_top_k()
        if not self.lantern_fused:
            mixed = uncond_tree_logits + self.cfg_scale * (tree_logits - uncond_tree_logits)
            procs = getattr(self, "internal_logits_processors", None)
            if procs is not None and len(procs) > 1:       # the reference's own processors (:600-605)
                rows = procs[0](mixed[0], image_start_token_id_index=self.image_start_token_id_index,
                                position_ids=position_ids + 1)
                rows = procs[1](rows)
            else:
                rows = lumina_process_rows(mixed[0], lumina_row_kinds(position_ids + 1, self.image_start_token_id_index),
                                           self._family(mixed.shape[-1]), top_k)
            return rows[retrieve_indices], hidden_states, uncond_hidden_states
        kinds = lumina_row_kinds(position_ids + 1, self.image_start_token_id_index)[None].contiguous()
        handle = TreeLogits(tree_logits, uncond_tree_logits, float(self.cfg_scale), retrieve_indices, kinds, top_k)
        return handle, hidden_states, uncond_hidden_states

    def evaluate_posterior(self, logits, candidates, cart_candidates_prob=None, original_prob=None, p_indices=None,
                           tree_candidates=None, b_indices=None, do_sample=True, lantern=False, lantern_k=1000,
                           lantern_delta=0.1):
        """ea_model_lumina_mgpt.py:610-729."""
        if not do_sample:
            raise NotImplementedError("Greedy decoding is not implemented yet")     # same as the reference (:728-729)
        fam = self._family(logits.shape[-1])
        si = None
        if self.eagle_version == 1:
            assert cart_candidates_prob is not None, "Cartesian candidate probabilities are required for EAGLE v1"
            assert original_prob is not None, "Original probabilities are required for EAGLE v1"
            assert tree_candidates is not None, "Tree candidates are required for EAGLE v1"
            assert p_indices is not None, "Parent indices are required for EAGLE v1"
            assert b_indices is not None, "B indices are required for EAGLE v1"
            fused = isinstance(logits, TreeLogits)
            si = _static_inputs(fused, candidates, cart_candidates_prob, original_prob, p_indices, tree_candidates,
                                b_indices, logits.device, logits.retrieve_indices if fused else None)
        return _verify(fam, logits, candidates, temp=1.0, top_p=1.0, top_k=0, lantern=lantern, lantern_k=lantern_k,
                       lantern_delta=lantern_delta, nearest_latents=getattr(self, "nearest_latents", None),
                       static_inputs=si, rng=self.lantern_rng)

    def update_inference_inputs(self, input_ids, attention_mask, candidates, best_candidate, accept_length,
                                retrieve_indices, do_sample, new_token, past_key_values_data, current_length_data,
                                hidden_states_new, uncond_hidden_states_new, sample_p):
        """ea_model_lumina_mgpt.py:731-799."""
        if self.cfg_mode == "parallel":
            prev = input_ids.shape[1]
            sel = retrieve_indices[best_candidate, : accept_length + 1] + prev
            current_length_data.fill_(kv_compact(past_key_values_data, sel, prev))
            input_ids = torch.cat([input_ids[None, 0], candidates[None, best_candidate, : accept_length + 1].to(input_ids.device)], dim=-1)
        else:
            for key in ("cond", "uncond"):
                prev = input_ids.shape[1] - (0 if key == "cond" else self.image_start_token_id_index)
                sel = retrieve_indices[best_candidate, : accept_length + 1] + prev
                current_length_data[key].fill_(kv_compact(past_key_values_data[key], sel, prev))
            input_ids = torch.cat([input_ids, candidates[None, best_candidate, : accept_length + 1].to(input_ids.device)], dim=-1)
        accept_hidden = hidden_states_new[:, retrieve_indices][:, best_candidate, : accept_length + 1]
        accept_uncond = uncond_hidden_states_new[:, retrieve_indices][:, best_candidate, : accept_length + 1]
        token = sample_bonus_token(sample_p, do_sample)
        output = self.ea_layer.topK_generate(
            hidden_states=accept_hidden, uncond_hidden_states=accept_uncond,
            input_ids=torch.cat((input_ids, token.to(input_ids.device)), dim=-1), attention_mask=attention_mask,
            head=self.base_model.lm_head, logits_processors=self.drafter_logits_processors,
            tree_type="static" if self.eagle_version == 1 else "dynamic")
        new_token += accept_length + 1
        return input_ids, output, new_token, token


def patch_reference(cls, lumina: bool = False, rng: str = "python", fused: bool = True):
    """Install the B200 verification path on a reference model class (``EaModel`` or ``EaLumina_mGPT``)."""
    src = LuminaVerifyMixin if lumina else VerifyMixin
    names = ["_family", "tree_decoding", "evaluate_posterior", "update_inference_inputs"]
    if not lumina:
        names += ["evaluate_posterior_v1", "_family_name"]
    for n in names:
        setattr(cls, n, getattr(src, n))
    cls.lantern_rng, cls.lantern_fused = rng, fused
    if lumina:
        cls.lantern_image_top_k = LuminaVerifyMixin.lantern_image_top_k
    return cls
