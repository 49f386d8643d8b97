"""Run casegen cases through the CUDA path (lantern_b200.Verifier) — shared by the GPU tests, smoke and bench."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

import casegen as C
from lantern_b200 import trees, verify
from lantern_b200 import choices as CH

_FAM = {"llamagen": verify.LLAMAGEN, "anole": verify.ANOLE, "lumina_mgpt": verify.LUMINA}


def family_spec(b: C.Built) -> verify.FamilySpec:
    p = b.params
    if p["family"] == "vanilla":
        return verify.vanilla(b.fam.vocab)
    base = _FAM[p["family"]]
    if b.fam.ncols == base.ncols and b.fam.vocab == base.vocab:
        return base
    return base.resized(b.fam.ncols, b.fam.vocab)


def pad_retrieve(ri_list: Sequence[np.ndarray]) -> np.ndarray:
    L = max(r.shape[0] for r in ri_list)
    D = max(r.shape[1] for r in ri_list)
    out = np.full((len(ri_list), L, D), -1, dtype=np.int32)
    for i, r in enumerate(ri_list):
        out[i, :r.shape[0], :r.shape[1]] = r
    return out


def run_cases(cases: List[C.Built], dtype=torch.float32, want_sample_p: bool = True, device="cuda",
              philox: Optional[tuple] = None, phases: int = 3):
    """All cases must share family / knobs / T (they become one batched launch)."""
    b0 = cases[0]
    p = b0.params
    fam = family_spec(b0)
    dev = torch.device(device)
    B = len(cases)
    st = None
    if b0.static is not None:
        st = trees.generate_tree_buffers(CH.tree(p["static_tree"]), device=dev)["static_tree"]
    table = None
    if p["lantern"]:
        table = torch.from_numpy(b0.table.astype(np.int32)).to(dev)
    k = min(int(p["lantern_k"]), b0.fam.ncols - 1)
    v = verify.Verifier(fam, temperature=p["temperature"], top_k=p["top_k"], top_p=p["top_p"],
                        cfg_scale=p["cfg_scale"], lantern=p["lantern"], lantern_k=k,
                        lantern_delta=p["lantern_delta"], nbr_table=table, static_tree=st, device=dev)
    cond = torch.from_numpy(np.stack([c.cond for c in cases])).to(dev).to(dtype)
    uncond = None
    if b0.uncond is not None:
        uncond = torch.from_numpy(np.stack([c.uncond for c in cases])).to(dev).to(dtype)
    tokens = torch.from_numpy(np.stack([c.tree.tokens for c in cases]).astype(np.int32)).to(dev)
    kinds = None
    if b0.row_kinds is not None:
        kinds = torch.from_numpy(np.stack([c.row_kinds for c in cases]).astype(np.uint8)).to(dev)
    uni = None
    if philox is None:
        uni = torch.from_numpy(np.stack([c.uniforms for c in cases]).astype(np.float32)).to(dev)
    kw = {}
    if st is not None:
        T = b0.tree.T
        node_q = np.ones((B, T), dtype=np.float32)
        for i, c in enumerate(cases):
            ri = c.tree.retrieve_indices
            m = ri >= 0
            node_q[i, ri[m]] = c.static.cart_prob[m]
        kw["node_q"] = torch.from_numpy(node_q).to(dev)
        kw["draft_op"] = torch.from_numpy(np.stack([np.concatenate(c.static.op, axis=0) for c in cases])).to(dev)
        retrieve = None
    else:
        retrieve = torch.from_numpy(pad_retrieve([c.tree.retrieve_indices for c in cases])).to(dev)
    res = v.step(cond, uncond, tokens, retrieve, row_kinds=kinds, uniforms=uni,
                 philox=philox or (0, 0), want_sample_p=want_sample_p, phases=phases, **kw)
    torch.cuda.synchronize()
    return res


def compare(res, i: int, orc, tol: float = 1e-5, check_token: bool = True):
    """Assert item i of a VerifyResult equals the oracle's StepResult."""
    a = int(res.accept_length[i])
    assert a == orc.accept_length, f"accept_length {a} != {orc.accept_length}"
    assert int(res.best_candidate[i]) == orc.best_candidate
    assert int(res.n_draws[i]) == orc.n_uniforms
    assert bool(int(res.flags[i]) & 1) == orc.residual_tail
    assert res.path_tokens[i, :a + 1].tolist() == orc.accepted_tokens.tolist()
    assert res.select_indices[i, :a + 1].tolist() == orc.select_indices.tolist()
    assert (res.path_tokens[i, a + 1:] == -1).all()
    if check_token:
        assert int(res.token[i]) == orc.token, f"token {int(res.token[i])} != {orc.token}"
    if res.sample_p is not None:
        assert_probs_close(res.sample_p[i].cpu().numpy(), orc.sample_p, tol)


def assert_probs_close(got: np.ndarray, want: np.ndarray, rtol: float = 1e-5):
    """north_star tolerance: 1e-5 relative in fp32.  Static-tree residuals are differences
    ``max(p - q, 0)`` of nearly equal numbers, so an absolute floor of 1e-6 x the largest entry is allowed
    (cancellation amplifies the ~1e-7 relative difference between two correct softmax implementations)."""
    atol = 1e-6 * float(want.max())
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bad = err > rtol * np.abs(want) + atol
    assert not bad.any(), (f"{int(bad.sum())} probabilities differ; worst abs err {err.max():.3e} "
                           f"at {int(err.argmax())}: got {got[err.argmax()]!r} want {want[err.argmax()]!r}")
    big = want > 10 * atol
    assert np.all(got[big] > 0), "support of sample_p differs"
    assert np.all(got[want == 0] <= atol)
