"""Neighbour-table build, host-buffer session and sweep edges on the GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

import casegen as CG
import cuda_runner as R
from lantern_b200 import _abi, codebook, verify
from oracle import lantern_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,d", [(64, 8), (512, 8), (1000, 8), (1024, 256), (4096, 8)])
def test_neighbor_table_bit_exact(N, d):
    rng = np.random.default_rng(N + d)
    E = rng.standard_normal((N, d)).astype(np.float32)
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    E[N // 2] = E[3]                                # exact duplicate rows -> distance ties broken by id
    want = O.neighbor_table(E)
    got = codebook.build_neighbor_table(torch.from_numpy(E).cuda()).cpu().numpy()
    assert got.shape == (N, N - 1) and np.array_equal(got, want)
    short = codebook.build_neighbor_table(torch.from_numpy(E).cuda(), k=17).cpu().numpy()
    assert np.array_equal(short, want[:, :17])


def test_reference_file_format_roundtrip(tmp_path):
    E = torch.randn(256, 8, device="cuda")
    t = codebook.build_neighbor_table(E)
    path = codebook.save_reference_format(t, str(tmp_path))
    raw = np.load(path)
    assert path.endswith("top_255_indices.npy") and raw.dtype == np.uint16 and raw.shape == (256, 255)
    back = codebook.load_neighbor_table(path, cols=11, device="cuda")
    assert back.dtype == torch.int32 and torch.equal(back, t[:, :11])


def _session_step(built, want_sample_p=False, pinned=False):
    """Drive lantern_session_* with plain host (numpy) buffers, or page-locked ones (pinned=True)."""
    lib = _abi.load()
    b0 = built[0]
    p = b0.params
    fam = R.family_spec(b0)
    B, T = len(built), b0.tree.T
    k = min(int(p["lantern_k"]), b0.fam.ncols - 1)
    table = np.ascontiguousarray(b0.table.astype(np.int32))
    ver = verify.Verifier(fam, temperature=p["temperature"], top_k=p["top_k"], cfg_scale=p["cfg_scale"], lantern=True,
                          lantern_k=k, lantern_delta=p["lantern_delta"], nbr_table=torch.from_numpy(table).cuda())
    cond = np.ascontiguousarray(np.stack([c.cond for c in built]))
    uncond = np.ascontiguousarray(np.stack([c.uncond for c in built]))
    keep = None
    if pinned:     # torch pinned tensors own the memory; numpy views keep the .ctypes plumbing below unchanged
        keep = [torch.from_numpy(cond).pin_memory(), torch.from_numpy(uncond).pin_memory()]
        cond, uncond = keep[0].numpy(), keep[1].numpy()
    tokens = np.ascontiguousarray(np.stack([c.tree.tokens for c in built]).astype(np.int32))
    ri = R.pad_retrieve([c.tree.retrieve_indices for c in built])
    uni = np.ascontiguousarray(np.stack([c.uniforms for c in built]).astype(np.float32))
    L, D = ri.shape[1:]
    cfg = ver._cfg(B, T, L, D, torch.from_numpy(cond), False, uni.shape[1], (0, 0))
    sess = C.c_void_p()
    _abi.check(lib.lantern_session_create(C.byref(cfg), table.ctypes.data, table.shape[0], C.byref(sess)))
    ain = _abi.AcceptIn()
    ain.logits_cond, ain.logits_uncond = cond.ctypes.data, uncond.ctypes.data
    ain.tree_tokens, ain.retrieve, ain.uniforms = tokens.ctypes.data, ri.ctypes.data, uni.ctypes.data
    out = {n: np.zeros(B, dtype=np.int32) for n in ("accept_length", "best_candidate", "token", "n_draws", "flags")}
    path, sel = np.zeros((B, D), dtype=np.int32), np.zeros((B, D), dtype=np.int32)
    sp = np.zeros((B, fam.vocab), dtype=np.float32)
    aout = _abi.AcceptOut()
    for n in out:
        setattr(aout, n, out[n].ctypes.data)
    aout.path_tokens, aout.select_indices = path.ctypes.data, sel.ctypes.data
    if want_sample_p:
        aout.sample_p = sp.ctypes.data
    for _ in range(2):                                   # second step reuses the session's buffers
        _abi.check(lib.lantern_session_step(sess, C.byref(cfg), C.byref(ain), C.byref(aout)))
    out["route"] = int(lib.lantern_session_last_route(sess))
    lib.lantern_session_destroy(sess)
    return out, path, sel, sp


@pytest.mark.parametrize("family,kw", [("llamagen", dict(ncols=4096, top_k=500, boost=11.0)),
                                       ("anole", dict(ncols=2048, top_k=400, boost=11.0)),
                                       ("lumina_mgpt", dict())])
@pytest.mark.parametrize("pinned", [False, True])
def test_host_buffer_session_matches_oracle(family, kw, pinned):
    built, orcs, seed = [], [], 12000
    while len(built) < 5:
        b = CG.build(dict(family=family, seed=seed, depth=5 if family == "lumina_mgpt" else 4, **kw))
        seed += 1
        o = CG.oracle_step(b)
        if o.margin >= 1e-5:
            built.append(b)
            orcs.append(o)
    out, path, sel, sp = _session_step(built, want_sample_p=True, pinned=pinned)
    # page-locked logits with a lazy-eligible window are read in place (zero-copy); everything else is staged
    lazy_ok = built[0].fam.ncols in (2048, 4096, 8192, 16384)
    assert out["route"] == (1 if pinned and lazy_ok else 0)
    for i, o in enumerate(orcs):
        a = int(out["accept_length"][i])
        rows_read = (int(out["flags"][i]) >> 8) & 0xFF
        assert 1 <= rows_read <= a + 2
        assert a == o.accept_length and int(out["best_candidate"][i]) == o.best_candidate
        assert int(out["token"][i]) == o.token and int(out["n_draws"][i]) == o.n_uniforms
        assert path[i, :a + 1].tolist() == o.accepted_tokens.tolist() and sel[i, :a + 1].tolist() == o.select_indices.tolist()
        R.assert_probs_close(sp[i], o.sample_p)


@pytest.mark.parametrize("B,T", [(1, 1), (3, 2), (2, 33), (17, 59), (5, 256)])
def test_accept_sweep_shapes(B, T):
    """BASELINE configs[4]: tree 1-256 x batch sizes; every item equals its own oracle."""
    built, orcs, seed = [], [], 20000 + 10 * T
    while len(built) < B:
        b = CG.build(dict(family="llamagen", ncols=2048, tree="random", total_tokens=T, top_k=300, lantern_k=100,
                          boost=10.0, seed=seed))
        seed += 1
        o = CG.oracle_step(b)
        if o.margin >= 1e-5:
            built.append(b)
            orcs.append(o)
    res = R.run_cases(built)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)


def test_host_session_python_api():
    """lantern_b200.session.HostSession: pinned torch tensors (in place) and numpy arrays (staged) give the oracle's
    results; the in-place route reads no more than accept_length + 2 rows per prompt."""
    from lantern_b200.session import HostSession
    built, orcs, seed = [], [], 12500
    while len(built) < 4:
        b = CG.build(dict(family="lumina_mgpt", seed=seed, depth=5))
        seed += 1
        o = CG.oracle_step(b)
        if o.margin >= 1e-5:
            built.append(b)
            orcs.append(o)
    b0 = built[0]
    fam = R.family_spec(b0)
    p = b0.params
    k = min(int(p["lantern_k"]), b0.fam.ncols - 1)
    ver = verify.Verifier(fam, temperature=p["temperature"], top_k=p["top_k"], cfg_scale=p["cfg_scale"], lantern=True,
                          lantern_k=k, lantern_delta=p["lantern_delta"],
                          nbr_table=torch.from_numpy(b0.table.astype(np.int32)).cuda())
    cond = np.ascontiguousarray(np.stack([c.cond for c in built]))
    uncond = np.ascontiguousarray(np.stack([c.uncond for c in built]))
    tokens = np.ascontiguousarray(np.stack([c.tree.tokens for c in built]).astype(np.int32))
    ri = R.pad_retrieve([c.tree.retrieve_indices for c in built])
    uni = np.ascontiguousarray(np.stack([c.uniforms for c in built]).astype(np.float32))
    with HostSession(ver, len(built), b0.tree.T, ri.shape[1], ri.shape[2], n_uniforms=uni.shape[1]) as sess:
        staged = sess.step(cond, uncond, tokens, ri, uniforms=uni, want_sample_p=True)
        pinned = sess.step(torch.from_numpy(cond).pin_memory(), torch.from_numpy(uncond).pin_memory(), tokens, ri,
                           uniforms=uni, want_sample_p=True)
    assert not staged.in_place and pinned.in_place
    # the caller refills the same page-locked buffers every step: in-place reads must see the new contents (no stale
    # device-side caching of host memory across steps).  Rotate the prompts inside the same pinned tensors.
    pc, pu = torch.from_numpy(cond).pin_memory(), torch.from_numpy(uncond).pin_memory()
    with HostSession(ver, len(built), b0.tree.T, ri.shape[1], ri.shape[2], n_uniforms=uni.shape[1]) as sess2:
        first = sess2.step(pc, pu, tokens, ri, uniforms=uni)
        perm = [1, 2, 3, 0]
        pc.copy_(torch.from_numpy(cond[perm]))
        pu.copy_(torch.from_numpy(uncond[perm]))
        again = sess2.step(pc, pu, np.ascontiguousarray(tokens[perm]), np.ascontiguousarray(ri[perm]),
                           uniforms=np.ascontiguousarray(uni[perm]))
    assert first.in_place and again.in_place
    for i, src in enumerate(perm):
        assert int(again.accept_length[i]) == orcs[src].accept_length and int(again.token[i]) == orcs[src].token
        assert int(first.accept_length[i]) == orcs[i].accept_length and int(first.token[i]) == orcs[i].token
    for r in (staged, pinned):
        for i, o in enumerate(orcs):
            assert int(r.accept_length[i]) == o.accept_length and int(r.token[i]) == o.token
            assert int(r.best_candidate[i]) == o.best_candidate and int(r.n_draws[i]) == o.n_uniforms
            R.assert_probs_close(r.sample_p[i], o.sample_p)
    assert np.all(pinned.rows_read <= pinned.accept_length + 2) and np.all(pinned.rows_read >= 1)


@pytest.mark.parametrize("N,d", [(512, 8), (1000, 8), (1024, 256), (2048, 40)])
def test_tensor_core_distance_gemm_within_bound(N, d):
    """tcgen05 TF32 distance GEMM against fp64 distances: inside the error bound the exact re-rank relies on."""
    lib = _abi.load()
    rng = np.random.default_rng(N * 7 + d)
    E = rng.standard_normal((N, d)).astype(np.float32)
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    Ed = torch.from_numpy(E).cuda()
    ld = (N + 3) & ~3
    D = torch.full((N, ld), -1.0, device="cuda")
    _abi.check(lib.lantern_debug_dist_gemm(Ed.data_ptr(), N, d, D.data_ptr(), ld, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = D[:, :N].cpu().numpy().astype(np.float64)
    E64 = E.astype(np.float64)
    want = ((E64[:, None, :] - E64[None, :, :]) ** 2).sum(-1) if N <= 1024 else \
        (E64 ** 2).sum(1)[:, None] + (E64 ** 2).sum(1)[None, :] - 2 * E64 @ E64.T
    err = np.abs(got - want).max()   # the diagonal (self, true distance 0) obeys the same bound
    assert err <= 1.5 * (0.0025 + 2 * 6e-5), f"max |D~ - d^2| = {err}"


@pytest.mark.parametrize("N,d,K", [(4096, 8, 101), (2048, 256, 65), (8192, 8, 1001), (1000, 8, 33)])
def test_tensor_core_neighbor_table_bit_exact(N, d, K):
    rng = np.random.default_rng(N + d + K)
    E = rng.standard_normal((N, d)).astype(np.float32)
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    E[5] = E[N // 3]                                   # duplicate row: tie broken by id
    want = O.neighbor_table(E, K)
    got, route = codebook.build_neighbor_table(torch.from_numpy(E).cuda(), k=K, return_route=True)
    got = got.cpu().numpy()
    assert route.tolist() == [1, 0], "tensor-core route sent rows to the fp64 kernel"
    assert np.array_equal(got, want)
    os_env = __import__("os").environ
    os_env["LANTERN_NBR_EXACT_ONLY"] = "1"             # the all-fp64 route gives the same table
    try:
        exact, route = codebook.build_neighbor_table(torch.from_numpy(E).cuda(), k=K, return_route=True)
        exact = exact.cpu().numpy()
        assert route.tolist() == [2, N]
    finally:
        del os_env["LANTERN_NBR_EXACT_ONLY"]
    assert np.array_equal(exact, want)


@pytest.mark.parametrize("top_k,temp,cfg", [(0, 1.0, False), (200, 1.0, True), (500, 0.8, True)])
def test_draft_sample_matches_oracle(top_k, temp, cfg):
    """Static-tree drafter sampling (next-row N3): tokens bit-exact, conditional probabilities and `op` within 1e-5."""
    from lantern_b200 import draft_sample, posterior, synth
    R_, V = 5, 4096
    cond = synth.gauss(77, (R_, V), stream=1)
    uncond = cond + synth.gauss(77, (R_, V), stream=2) * np.float32(0.25) if cfg else None
    proc = posterior.prepare_logits_processor(temperature=temp, top_p=1.0, top_k=top_k)
    idx, cp, probs = draft_sample.sample(torch.from_numpy(cond).cuda(), proc, k=10,
                                         uncond=torch.from_numpy(uncond).cuda() if cfg else None, cfg_scale=3.0,
                                         seed=99, step=4)
    torch.cuda.synchronize()
    for r in range(R_):
        row = O.cfg_mix(cond[r], uncond[r], 3.0) if cfg else cond[r]
        oi, ocp, op = O.draft_sample(row, O.Warp(temp, 1.0, top_k), 10, 99, 4, r)
        assert idx[r].tolist() == oi.tolist()
        R.assert_probs_close(probs[r].cpu().numpy(), op)
        assert np.allclose(cp[r].cpu().numpy(), ocp, rtol=1e-5, atol=1e-7)
        assert len(set(oi.tolist())) == 10                       # without replacement


def test_draft_sample_law():
    """The exponential race draws the first token with probability p (chi-square-free sanity check on a tiny vocab)."""
    from lantern_b200 import draft_sample, posterior
    V = 8
    logits = torch.tensor([[2.0, 1.0, 0.0, -1.0, 0.5, 1.5, -0.5, 0.2]], device="cuda").repeat(1, 1)
    p = torch.softmax(logits[0], 0).cpu().numpy()
    proc = posterior.prepare_logits_processor(temperature=1.0, top_p=1.0, top_k=0)
    counts = np.zeros(V)
    n = 4000
    big = logits.repeat(n, 1).contiguous()
    idx, _, _ = draft_sample.sample(big, proc, k=3, seed=7, step=1)
    first = idx[:, 0].cpu().numpy()
    for t in first:
        counts[t] += 1
    assert np.abs(counts / n - p).max() < 0.03
