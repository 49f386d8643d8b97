"""The reference's generate() loop bodies driven through the drop-in methods for the families / branches that
tests/test_dropin_loop_gpu.py (LlamaGen, dynamic tree) does not cover:

* Lumina-mGPT, dynamic tree, parallel CFG          ea_model_lumina_mgpt.py:936-1005, :556-608, :731-799
* Lumina-mGPT, static tree (eagle_version 1), sequential CFG (two KV caches, two current_length tensors)
* Anole, dynamic tree, input_position_diff != 0 + attention mask      ea_model_anole.py:1090-1148, :904-933
* LlamaGen, static_tree=True branch                                    ea_model_llamagen.py:1109-1125

Each step is checked against the oracle on the same uniforms (python `random` replayed), and the bookkeeping the
reference relies on is asserted: accepted tokens appended, KV rows compacted to the accepted positions (value of a KV
row = its absolute position), current_length advanced, bonus token consistent with sample_p.  Targets and drafters are
stand-ins (the 7B forwards are out of scope): synthetic logits in which every drafted token is boosted in its parent's
row, so walks go several levels deep."""
import random

import numpy as np
import pytest
import torch

from lantern_b200 import choices as CH
from lantern_b200 import posterior as PO
from lantern_b200 import synth, trees
from oracle import lantern_oracle as O

pytestmark = pytest.mark.gpu
H, S_MAX = 16, 512
MARGIN = 1e-5


def _mask_of(tree):
    T = tree.T
    mask = torch.zeros(1, 1, T, T)
    for i in range(T):
        a = i
        while a >= 0:
            mask[0, 0, i, a] = 1
            a = int(tree.parent[a])
    return mask


class _Kv:
    """KV slab [2*layers, batch, heads, S, head_dim]; a row's value is its absolute position."""

    def __init__(self, dev, batch):
        self.data = torch.zeros(4, batch, 2, S_MAX, 8, device=dev, dtype=torch.bfloat16)

    def write(self, pos_rows):
        """The target appends the T tree tokens at slots len .. len+T-1 (KVCache.cat, drafters/kv_cache.py:38-52), node
        i at slot len + i; position ids (len + depth) only feed RoPE.  The root's position id is len."""
        for bi, pos in enumerate(pos_rows):
            slots = int(pos[0]) + torch.arange(pos.shape[0], device=pos.device)
            self.data[:, bi, :, slots, :] = slots.to(torch.bfloat16)[None, None, :, None]


def _check_kv(slab, batch_index, prev_len, select_positions):
    got = slab[0, batch_index, 0, prev_len:prev_len + len(select_positions), 0].float().cpu().numpy()
    assert got.tolist() == [float(x) for x in select_positions]


def _replayed_uniforms(state_before, n):
    after = random.getstate()
    random.setstate(state_before)
    u = [random.random() for _ in range(n)]
    random.setstate(after)
    return u


# ------------------------------------------------------------------------------------------------------------------
# Lumina-mGPT
# ------------------------------------------------------------------------------------------------------------------
NC = 2048                                   # image tokens of the stand-in (ids 4 .. 4+NC-1); vocab keeps 8196 / 8803
FAM_L = O.small_family(O.LUMINA, NC)


class _LuminaDrafter:
    def __init__(self, owner):
        self.owner, self.calls = owner, 0

    def topK_generate(self, hidden_states, uncond_hidden_states, input_ids, attention_mask, head, logits_processors,
                      tree_type):
        self.calls += 1
        o = self.owner
        dev = input_ids.device
        assert hidden_states.shape == uncond_hidden_states.shape and hidden_states.shape[1] >= 1
        root = int(input_ids[0, -1])
        if tree_type == "static":
            counts = synth.static_group_counts(o.tree_choices)
            d = synth.static_draft(700 + self.calls, counts, FAM_L.vocab, FAM_L.col0, FAM_L.col1, sharp=1.0)
            o.static_draft = d
            return (torch.from_numpy(d.ss_token).to(dev), torch.from_numpy(d.ss_prob).to(dev),
                    [torch.from_numpy(x).to(dev) for x in d.op])
        tree = synth.eagle2_tree(700 + self.calls, 40, 5)
        synth.assign_tokens(700 + self.calls, tree, FAM_L.col0, FAM_L.col1, root_token=root)
        # a real drafter proposes the newline token where the target will demand it: do so on even calls
        if self.calls % 2 == 0:
            pos0 = input_ids.shape[1] - 1       # position of the root token
            kinds = PO.lumina_row_kinds(torch.from_numpy(tree.depth) + pos0 + 1, o.image_start_token_id_index).numpy()
            seen = set()
            for i in range(1, tree.T):
                par = int(tree.parent[i])
                if kinds[par] == O.ROW_NEWLINE and par not in seen:
                    tree.tokens[i] = O.LUMINA_NEWLINE_TOKEN
                    seen.add(par)
        o.tree = tree
        return (torch.from_numpy(tree.tokens)[None].to(dev), torch.from_numpy(tree.retrieve_indices).to(dev),
                _mask_of(tree).to(dev), torch.from_numpy(tree.depth).to(dev))


class _Lumina(PO.LuminaVerifyMixin):
    image_start_token_id_index = 9
    lantern_image_tokens = NC

    def __init__(self, dev, cfg_mode, eagle_version, top_k):
        self.dev, self.cfg_mode, self.eagle_version = dev, cfg_mode, eagle_version
        self.cfg_scale = 3.0
        self.ea_layer = _LuminaDrafter(self)
        self.base_model = type("B", (), {"lm_head": None})()
        self.drafter_logits_processors = []
        tk = type("InterleavedTopKLogitsWarper", (), {})()
        tk.image_top_k = top_k
        self.internal_logits_processors = [object(), tk]            # as generate(top_k=...) leaves them (:822-823)
        self.nearest_latents = synth.neighbor_table(0, NC, 101)
        self.calls = 0
        self.tree = None
        self.kv = {"cond": _Kv(dev, 2 if cfg_mode == "parallel" else 1)}
        if cfg_mode != "parallel":
            self.kv["uncond"] = _Kv(dev, 1)
        self.seen_positions = []

    def __call__(self, input_ids=None, attention_mask=None, output_orig=True, past_key_values=None, position_ids=None):
        T = input_ids.shape[1]
        tree = self.tree
        assert input_ids[0].tolist() == tree.tokens.tolist()
        if self.cfg_mode == "parallel":
            self.calls += 1
            assert input_ids.shape[0] == 2 and position_ids.shape == (2, T)
            assert torch.equal(position_ids[1], position_ids[0] - self.image_start_token_id_index)
            self.step_logits = synth.tree_logits(300 + self.calls, tree, FAM_L.vocab, cfg=True, boost=11.0)
            # parallel CFG keeps both branches in one left-padded cache: same slots for both batch rows
            self.kv["cond"].write([position_ids[0], position_ids[0]])
            self.seen_positions.append(position_ids[0].clone())
            tl = torch.from_numpy(np.stack(self.step_logits)).to(self.dev)
            return None, tl, torch.zeros(2, T, H, device=self.dev)
        # sequential CFG: the cond call comes first, then the uncond call with shifted positions
        which = "cond" if past_key_values == "cond" else "uncond"
        if which == "cond":
            self.calls += 1
            self.step_logits = synth.tree_logits(300 + self.calls, tree, FAM_L.vocab, cfg=True, boost=8.0)
            self.seen_positions.append(position_ids.clone())
        else:
            assert torch.equal(position_ids, self.seen_positions[-1] - self.image_start_token_id_index)
        self.kv[which].write([position_ids])
        tl = torch.from_numpy(self.step_logits[0 if which == "cond" else 1])[None].to(self.dev)
        return None, tl, torch.zeros(1, T, H, device=self.dev)


def _lumina_oracle(m, tree, top_k, kinds, uniforms, lantern_k, delta, static=None):
    cond, uncond = m.step_logits
    return O.verify_step(cond, uncond, 3.0, tree.tokens, tree.retrieve_indices, np.asarray(uniforms), FAM_L,
                         O.Warp(1.0, 1.0, top_k), True, lantern_k, delta, m.nearest_latents, static=static,
                         row_kinds=kinds)


@pytest.mark.parametrize("top_k", [400, 1500])
def test_lumina_dynamic_parallel_cfg_loop(top_k):
    dev = torch.device("cuda")
    m = _Lumina(dev, "parallel", 2, top_k)
    random.seed(77)
    torch.manual_seed(77)
    isi = m.image_start_token_id_index
    # n = generated image tokens so far = len - (isi + 3); start 6 tokens before the end of an image row (w = 48)
    n0 = 42
    input_ids = torch.randint(FAM_L.col0, FAM_L.col1, (2, isi + 3 + n0), device=dev)
    attn_mask = torch.ones(2, input_ids.shape[1], dtype=torch.long, device=dev)
    cur_len = torch.zeros(4, dtype=torch.long)
    output = m.ea_layer.topK_generate(torch.zeros(1, 1, H, device=dev), torch.zeros(1, 1, H, device=dev),
                                      input_ids[:1], attn_mask, None, [], "dynamic")
    tree_candidates, retrieve_indices, tree_mask, tree_position_ids = output
    new_token, checked, newline_tokens = 0, 0, 0
    for step in range(8):
        tree = m.tree
        prev_len = input_ids.shape[1]
        logits, hs, uhs = m.tree_decoding(tree_candidates, attn_mask, None, tree_position_ids, input_ids, retrieve_indices)
        assert isinstance(logits, PO.TreeLogits) and logits.top_k == top_k
        kinds = O.lumina_row_kinds((tree.depth + prev_len + 1), isi)
        assert logits.row_kinds[0].cpu().numpy().tolist() == kinds.tolist()
        padded = torch.cat((tree_candidates, torch.full((1, 1), -1, device=dev, dtype=tree_candidates.dtype)), dim=1)
        candidates = padded[0, retrieve_indices]
        st = random.getstate()
        best, a, sample_p = m.evaluate_posterior(logits=logits, candidates=candidates, do_sample=True, lantern=True,
                                                 lantern_k=100, lantern_delta=0.1)
        u = _replayed_uniforms(st, tree.T)
        o = _lumina_oracle(m, tree, top_k, kinds, u + [0.5], 100, 0.1)
        if o.margin >= MARGIN:
            assert (int(best), a) == (o.best_candidate, o.accept_length)
            checked += 1
        input_ids_new, output, new_token, token = m.update_inference_inputs(
            input_ids=input_ids, attention_mask=attn_mask, candidates=candidates, best_candidate=best, accept_length=a,
            retrieve_indices=retrieve_indices, do_sample=True, new_token=new_token, past_key_values_data=[m.kv["cond"].data],
            current_length_data=cur_len, hidden_states_new=hs, uncond_hidden_states_new=uhs, sample_p=sample_p)
        torch.cuda.synchronize()
        assert input_ids_new.shape == (1, prev_len + a + 1)
        assert int(cur_len[0]) == prev_len + a + 1
        sel = (retrieve_indices[int(best), :a + 1] + prev_len).tolist()
        for bi in (0, 1):
            _check_kv(m.kv["cond"].data, bi, prev_len, sel)
        assert token.shape == (1, 1) and float(sample_p[int(token)]) > 0
        newline_tokens += int(int(token) == O.LUMINA_NEWLINE_TOKEN) + int(
            O.LUMINA_NEWLINE_TOKEN in input_ids_new[0, prev_len:].tolist())
        # the reference keeps input_ids [1, len] after the first parallel step (input_ids[None, 0], :751) and re-doubles
        # nothing: later steps read only shape[1] and row 0
        input_ids = input_ids_new
        attn_mask = torch.ones(2, input_ids.shape[1], dtype=torch.long, device=dev)
        tree_candidates, retrieve_indices, tree_mask, tree_position_ids = output
    assert checked >= 5 and newline_tokens >= 1, "the loop must cross an image-row boundary"
    assert new_token == input_ids.shape[1] - (isi + 3 + n0)


@pytest.mark.parametrize("tree_name,k,lam", [("mc_sim_7b_63", 10, 10.0), ("synth_40_3", 5, 20.0)])
def test_lumina_static_sequential_cfg_loop(tree_name, k, lam):
    dev = torch.device("cuda")
    m = _Lumina(dev, "sequential", 1, 400)
    m.tree_choices = CH.tree(tree_name)
    tb = trees.generate_tree_buffers(m.tree_choices, device=dev)
    random.seed(99)
    torch.manual_seed(99)
    isi = m.image_start_token_id_index
    input_ids = torch.randint(FAM_L.col0, FAM_L.col1, (1, isi + 3 + 5), device=dev)
    cur_len = {"cond": torch.zeros(4, dtype=torch.long), "uncond": torch.zeros(4, dtype=torch.long)}
    sample_token = input_ids[:, -1:].clone()
    tree_logits = m.ea_layer.topK_generate(torch.zeros(1, 1, H, device=dev), torch.zeros(1, 1, H, device=dev), input_ids,
                                           None, None, [], "static")
    retrieve_indices, tree_position_ids = tb["retrieve_indices"], tb["tree_position_ids"]
    parent = np.asarray(tb["parents"], dtype=np.int64)
    new_token, checked = 0, 0
    for step in range(5):
        prev_len = input_ids.shape[1]
        candidates, cart_prob, tree_candidates = trees.generate_candidates(tree_logits, tb["tree_indices"],
                                                                          retrieve_indices, sample_token)
        tree = synth.Tree(parent, tree_position_ids.cpu().numpy(), retrieve_indices.cpu().numpy())
        tree.tokens = tree_candidates[0].cpu().numpy()
        m.tree = tree
        logits, hs, uhs = m.tree_decoding(tree_candidates, None, {"cond": "cond", "uncond": "uncond"},
                                          tree_position_ids, input_ids, retrieve_indices)
        kinds = O.lumina_row_kinds(tree.depth + prev_len + 1, isi)
        st = random.getstate()
        best, a, sample_p = m.evaluate_posterior(
            logits=logits, candidates=candidates, cart_candidates_prob=cart_prob, original_prob=tree_logits[2],
            tree_candidates=tree_candidates, p_indices=tb["p_indices"], b_indices=tb["b_indices"], do_sample=True,
            lantern=True, lantern_k=k, lantern_delta=lam)
        u = _replayed_uniforms(st, tree.T)
        d = m.static_draft
        ob = O.generate_tree_buffers(m.tree_choices)
        static = O.StaticDraft(cart_prob.cpu().numpy(), d.op, ob["p_indices"], ob["b_indices"], tree.tokens)
        o = _lumina_oracle(m, tree, 400, kinds, u + [0.5], k, lam, static=static)
        if o.margin >= MARGIN:
            assert (int(best), a) == (o.best_candidate, o.accept_length)
            checked += 1
        kvd = {"cond": [m.kv["cond"].data], "uncond": [m.kv["uncond"].data]}
        input_ids, tree_logits, new_token, sample_token = m.update_inference_inputs(
            input_ids=input_ids, attention_mask=None, candidates=candidates, best_candidate=best, accept_length=a,
            retrieve_indices=retrieve_indices, do_sample=True, new_token=new_token, past_key_values_data=kvd,
            current_length_data=cur_len, hidden_states_new=hs, uncond_hidden_states_new=uhs, sample_p=sample_p)
        torch.cuda.synchronize()
        assert input_ids.shape == (1, prev_len + a + 1)
        rel = retrieve_indices[int(best), :a + 1].tolist()
        assert int(cur_len["cond"][0]) == prev_len + a + 1 and int(cur_len["uncond"][0]) == prev_len - isi + a + 1
        _check_kv(m.kv["cond"].data, 0, prev_len, [r + prev_len for r in rel])
        _check_kv(m.kv["uncond"].data, 0, prev_len - isi, [r + prev_len - isi for r in rel])
        assert sample_token.shape == (1, 1) and float(sample_p[int(sample_token)]) > 0
        assert isinstance(tree_logits, tuple) and len(tree_logits) == 3
    assert checked >= 3 and new_token == input_ids.shape[1] - (isi + 3 + 5)


# ------------------------------------------------------------------------------------------------------------------
# Anole (input_position_diff, attention mask) and LlamaGen static_tree=True
# ------------------------------------------------------------------------------------------------------------------
class _EaDrafter:
    def __init__(self, owner):
        self.owner, self.calls, self.kwargs = owner, 0, []

    def topK_genrate(self, hidden, input_ids, head, logits_processor, cfg_scale, **kw):
        self.calls += 1
        self.kwargs.append(kw)
        o = self.owner
        assert input_ids.shape[0] == 2 and torch.equal(input_ids[0], input_ids[1])      # .repeat(2, 1), :986
        tree = synth.eagle2_tree(500 + self.calls, 30, 4)
        synth.assign_tokens(500 + self.calls, tree, o.fam.col0, o.fam.col1, root_token=int(input_ids[0, -1]))
        o.tree = tree
        dev = input_ids.device
        return (torch.from_numpy(tree.tokens)[None].to(dev), torch.from_numpy(tree.retrieve_indices).to(dev),
                _mask_of(tree).to(dev), torch.from_numpy(tree.depth).to(dev))

    def topK_genrate_v1(self, hidden, input_ids, head, logits_processor, cfg_scale, **kw):
        self.calls += 1
        self.kwargs.append(kw)
        o = self.owner
        counts = synth.static_group_counts(o.tree_choices)
        d = synth.static_draft(600 + self.calls, counts, o.fam.vocab, o.fam.col0, o.fam.col1, sharp=1.0)
        o.static_draft = d
        dev = input_ids.device
        return (torch.from_numpy(d.ss_token).to(dev), torch.from_numpy(d.ss_prob).to(dev),
                [torch.from_numpy(x).to(dev) for x in d.op])


class _Ea(PO.VerifyMixin):
    def __init__(self, dev, family, ncols):
        self.dev = dev
        self.lantern_family = family
        self.lantern_image_tokens = ncols
        self.fam = O.small_family(O.ANOLE if family == "anole" else O.LLAMAGEN, ncols)
        if family == "anole":
            self.image_token_offset = 4
        self.ea_layer = _EaDrafter(self)
        self.base_model = type("B", (), {"lm_head": None})()
        self.kv = _Kv(dev, 2)
        self.nearest_latents = synth.neighbor_table(0, ncols, 101)
        self.calls = 0
        self.tree = None
        self.last_call = None

    def __call__(self, input_ids=None, output_orig=True, past_key_values=None, position_ids=None, attention_mask=None):
        self.calls += 1
        T = input_ids.shape[1]
        assert input_ids.shape[0] == 2 and input_ids[0].tolist() == self.tree.tokens.tolist()
        self.last_call = dict(position_ids=position_ids, attention_mask=attention_mask)
        self.step_logits = synth.tree_logits(800 + self.calls, self.tree, self.fam.vocab, cfg=True, boost=self.boost)
        pos = position_ids.reshape(-1, T)
        self.kv.write([pos[0], pos[0]])
        tl = torch.from_numpy(np.stack(self.step_logits)).to(self.dev)
        return None, tl, torch.zeros(2, T, H, device=self.dev)


def test_anole_dynamic_loop_with_position_diff_and_mask():
    dev = torch.device("cuda")
    m = _Ea(dev, "anole", 2048)
    m.boost = 11.0
    proc = PO.prepare_logits_processor(temperature=1.0, top_p=1.0, top_k=400)
    random.seed(5)
    torch.manual_seed(5)
    diff = 7                                   # cond prompt is 7 tokens longer than the (left-padded) uncond prompt
    input_ids = torch.randint(m.fam.col0, m.fam.col1, (1, 23), device=dev)
    input_mask = torch.ones(2, 23, dtype=torch.long, device=dev)
    input_mask[1, :diff] = 0
    cur_len = torch.zeros(4, dtype=torch.long)
    draft_tokens, retrieve_indices, tree_mask, tree_pos = m.ea_layer.topK_genrate(
        None, input_ids.repeat(2, 1), None, proc, 3.0, input_position_diff=diff, attention_mask=input_mask)
    new_token, checked = 0, 0
    for step in range(6):
        tree = m.tree
        prev_len = input_ids.shape[1]
        tree_draft = torch.cat([draft_tokens, draft_tokens])
        logits, hidden, _ = m.tree_decoding(tree_draft, None, tree_pos, input_ids, retrieve_indices, 3.0, input_mask, diff)
        # ea_model_anole.py:915-922: two position rows (uncond shifted by the diff), mask padded with ones to the new length
        pos = m.last_call["position_ids"]
        assert pos.shape == (2, tree.T) and torch.equal(pos[0], tree_pos + prev_len) and torch.equal(pos[1], pos[0] - diff)
        am = m.last_call["attention_mask"]
        assert am.shape == (2, prev_len + tree.T) and bool(am[:, input_mask.shape[1]:].all())
        assert torch.equal(am[:, :input_mask.shape[1]], input_mask)
        padded = torch.cat((draft_tokens, torch.full((1, 1), -1, device=dev, dtype=draft_tokens.dtype)), dim=1)
        candidates = padded[0, retrieve_indices]
        st = random.getstate()
        best, a, sample_p = m.evaluate_posterior(logits, candidates, proc, lantern=True, lantern_k=100, lantern_delta=0.1)
        u = _replayed_uniforms(st, tree.T)
        cond, uncond = m.step_logits
        o = O.verify_step(cond, uncond, 3.0, tree.tokens, tree.retrieve_indices, np.asarray(u + [0.5]), m.fam,
                          O.Warp(1.0, 1.0, 400), True, 100, 0.1, m.nearest_latents)
        if o.margin >= MARGIN:
            assert (int(best), a) == (o.best_candidate, o.accept_length)
            checked += 1
        assert float(sample_p[:m.fam.col0].sum()) == 0 and float(sample_p[m.fam.col1:].sum()) == 0   # :931 mask
        out = m.update_inference_inputs(input_ids, candidates, best, a, retrieve_indices, proc, new_token, [m.kv.data],
                                        cur_len, hidden, sample_p, 3.0, diff, attention_mask=input_mask)
        input_ids, draft_tokens, retrieve_indices, tree_mask, tree_pos, new_token, _, token = out
        torch.cuda.synchronize()
        assert m.ea_layer.kwargs[-1] == dict(input_position_diff=diff, attention_mask=input_mask)
        assert input_ids.shape == (1, prev_len + a + 1) and int(cur_len[0]) == prev_len + a + 1
        sel = [int(x) + prev_len for x in tree.retrieve_indices[int(best), :a + 1]]
        _check_kv(m.kv.data, 0, prev_len, sel)
        _check_kv(m.kv.data, 1, prev_len, sel)
        assert token.shape == (1, 1) and m.fam.col0 <= int(token) < m.fam.col1
    assert checked >= 4 and new_token == input_ids.shape[1] - 23


@pytest.mark.parametrize("family", ["llamagen", "anole"])
def test_static_tree_branch_loop(family):
    """generate()'s static_tree branch: generate_candidates -> tree_decoding -> evaluate_posterior_v1 ->
    update_inference_inputs(static_tree=True), which returns the 5-tuple (ea_model_llamagen.py:1109-1125, :988-993)."""
    dev = torch.device("cuda")
    m = _Ea(dev, family, 4096)
    m.boost = 8.5
    m.tree_choices = CH.tree("mc_sim_7b_63")
    tb = trees.generate_tree_buffers(m.tree_choices, device=dev)
    proc = PO.prepare_logits_processor(temperature=1.0, top_p=1.0, top_k=500)
    random.seed(11)
    torch.manual_seed(11)
    input_ids = torch.randint(m.fam.col0, m.fam.col1, (1, 19), device=dev)
    cur_len = torch.zeros(4, dtype=torch.long)
    sample_token = input_ids[:, -1:].clone()
    extra = dict(input_position_diff=3, attention_mask=None) if family == "anole" else {}
    tree_logits = m.ea_layer.topK_genrate_v1(None, input_ids.repeat(2, 1), None, proc, 3.0, **extra)
    parent = np.asarray(tb["parents"], dtype=np.int64)
    ri = tb["retrieve_indices"]
    new_token, checked = 0, 0
    ob = O.generate_tree_buffers(m.tree_choices)
    for step in range(5):
        prev_len = input_ids.shape[1]
        candidates, cart_prob, tree_candidates = trees.generate_candidates(tree_logits, tb["tree_indices"], ri,
                                                                          sample_token, proc)
        tree = synth.Tree(parent, tb["tree_position_ids"].cpu().numpy(), ri.cpu().numpy())
        tree.tokens = tree_candidates[0].cpu().numpy()
        m.tree = tree
        tc2 = torch.cat([tree_candidates, tree_candidates])
        args = (tc2, None, tb["tree_position_ids"], input_ids, ri, 3.0)
        logits, hidden, _ = m.tree_decoding(*args, None, 3) if family == "anole" else m.tree_decoding(*args)
        st = random.getstate()
        best, a, sample_p = m.evaluate_posterior_v1(logits, candidates, proc, cart_prob, tree_logits[2], tb["p_indices"],
                                                    tc2, tb["b_indices"], True, 10, 10.0)
        u = _replayed_uniforms(st, tree.T)
        cond, uncond = m.step_logits
        static = O.StaticDraft(cart_prob.cpu().numpy(), m.static_draft.op, ob["p_indices"], ob["b_indices"], tree.tokens)
        o = O.verify_step(cond, uncond, 3.0, tree.tokens, tree.retrieve_indices, np.asarray(u + [0.5]), m.fam,
                          O.Warp(1.0, 1.0, 500), True, 10, 10.0, m.nearest_latents, static=static)
        if o.margin >= MARGIN:
            assert (int(best), a) == (o.best_candidate, o.accept_length)
            checked += 1
        if family == "anole":
            out = m.update_inference_inputs(input_ids, candidates, best, a, ri, proc, new_token, [m.kv.data], cur_len,
                                            hidden, sample_p, 3.0, 3, attention_mask=None, static_tree=True)
        else:
            out = m.update_inference_inputs(input_ids, candidates, best, a, ri, proc, new_token, [m.kv.data], cur_len,
                                            hidden, sample_p, 3.0, static_tree=True)
        assert len(out) == 5
        input_ids, tree_logits, new_token, hidden_state, sample_token = out
        torch.cuda.synchronize()
        assert hidden_state is None and isinstance(tree_logits, tuple) and len(tree_logits) == 3
        assert input_ids.shape == (1, prev_len + a + 1) and int(cur_len[0]) == prev_len + a + 1
        _check_kv(m.kv.data, 0, prev_len, [int(x) + prev_len for x in tree.retrieve_indices[int(best), :a + 1]])
        assert sample_token.shape == (1, 1) and float(sample_p[int(sample_token)]) > 0
    assert checked >= 3 and new_token == input_ids.shape[1] - 19
