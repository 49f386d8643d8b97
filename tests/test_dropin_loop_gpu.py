"""Drop-in plumbing: the mixin methods installed on a stand-in EaModel run a short speculative-decoding loop
(tree_decoding -> evaluate_posterior -> update_inference_inputs) exactly like the reference's generate() body
(ea_model_llamagen.py:1109-1163), with a random-init stand-in target and a fixed-shape drafter stub.  Checks the
bookkeeping the reference relies on: accepted tokens appended, KV slab compacted to the accepted positions,
current_length advanced, bonus token drawn from sample_p, and the same decisions as the oracle."""
import random

import numpy as np
import pytest
import torch

from lantern_b200 import posterior as PO
from lantern_b200 import synth
from oracle import lantern_oracle as O

pytestmark = pytest.mark.gpu
V, H, S_MAX = 4096, 16, 256


class _Drafter:
    """topK_genrate stub: a fresh EAGLE-2-shaped tree with hashed tokens every call (the neural drafter is out of scope)."""

    def __init__(self):
        self.calls = 0

    def topK_genrate(self, hidden, input_ids, head, logits_processor, cfg_scale, **kw):
        self.calls += 1
        tree = synth.eagle2_tree(900 + self.calls, 26, 4)
        synth.assign_tokens(900 + self.calls, tree, 0, V, root_token=int(input_ids[0, -1]))
        dev = input_ids.device
        T = tree.T
        mask = torch.zeros(1, 1, T, T)
        for i in range(T):
            a = i
            while a >= 0:
                mask[0, 0, i, a] = 1
                a = int(tree.parent[a])
        return (torch.from_numpy(tree.tokens)[None].to(dev), torch.from_numpy(tree.retrieve_indices).to(dev), mask.to(dev),
                torch.from_numpy(tree.depth).to(dev))


class _Model(PO.VerifyMixin):
    """Stand-in for EaModel: `self(...)` returns (outputs, logits [2,T,V], hidden [2,T,H]) and writes the new KV rows."""
    lantern_family = "llamagen"
    lantern_image_tokens = V

    def __init__(self, dev):
        self.dev = dev
        self.ea_layer = _Drafter()
        self.base_model = type("B", (), {"lm_head": None})()
        self.kv = torch.zeros(4, 2, 2, S_MAX, 8, device=dev, dtype=torch.bfloat16)   # [2*layers, batch, heads, S, hd]
        self.calls = 0
        self.nearest_latents = synth.neighbor_table(0, V, 101)
        self.last_logits = None

    def __call__(self, input_ids=None, output_orig=True, past_key_values=None, position_ids=None, attention_mask=None):
        self.calls += 1
        T = input_ids.shape[1]
        tree_tokens = input_ids[0].cpu().numpy()
        base = synth.gauss(self.calls, (T, V), stream=5)
        cond = base + synth.gauss(self.calls, (T, V), stream=6) * np.float32(0.25)
        uncond = base + synth.gauss(self.calls, (T, V), stream=7) * np.float32(0.25)
        logits = torch.from_numpy(np.stack([cond, uncond])).to(self.dev)
        self.last_logits = (cond, uncond)
        # KV rows of the tree tokens: node i is appended at slot len + i (the root's position id is len); value = slot,
        # so compaction is easy to check
        slots = int(position_ids.reshape(-1, T)[0, 0]) + torch.arange(T, device=self.dev)
        self.kv[:, :, :, slots, :] = slots.to(torch.bfloat16)[None, None, None, :, None]
        hidden = torch.zeros(2, T, H, device=self.dev)
        return None, logits, hidden


def test_reference_style_loop_with_dropin_methods():
    dev = torch.device("cuda")
    m = _Model(dev)
    proc = PO.prepare_logits_processor(temperature=1.0, top_p=1.0, top_k=500)
    random.seed(1234)
    torch.manual_seed(1234)
    input_ids = torch.randint(0, V, (1, 17), device=dev)
    cur_len = torch.zeros(8, dtype=torch.long)
    draft_tokens, retrieve_indices, tree_mask, tree_pos = m.ea_layer.topK_genrate(None, input_ids, None, proc, 3.0)
    new_token, n_steps = 0, 6
    for step in range(n_steps):
        prev_len = input_ids.shape[1]
        tree_candidates = draft_tokens.repeat(2, 1)
        logits, hidden, _ = m.tree_decoding(tree_candidates, None, tree_pos, input_ids, retrieve_indices, 3.0)
        assert isinstance(logits, PO.TreeLogits)
        padded = torch.cat((draft_tokens, torch.full((1, 1), -1, device=dev, dtype=draft_tokens.dtype)), dim=1)
        candidates = padded[0, retrieve_indices]
        py_state = random.getstate()
        best, a, sample_p = m.evaluate_posterior(logits, candidates, proc, lantern=True, lantern_k=100, lantern_delta=0.1)
        # oracle on the same uniforms
        after = random.getstate()
        random.setstate(py_state)
        u = [random.random() for _ in range(draft_tokens.shape[1])]
        random.setstate(after)
        cond, uncond = m.last_logits
        fam = O.small_family(O.LLAMAGEN, V)
        o = O.verify_step(cond, uncond, 3.0, draft_tokens[0].cpu().numpy(), retrieve_indices.cpu().numpy(),
                          np.asarray(u + [0.5]), fam, O.Warp(1.0, 1.0, 500), True, 100, 0.1, m.nearest_latents)
        if o.margin >= 1e-5:
            assert int(best) == o.best_candidate and a == o.accept_length
        retrieve_prev = retrieve_indices
        out = m.update_inference_inputs(input_ids, candidates, best, a, retrieve_indices, proc, new_token, [m.kv], cur_len,
                                        hidden, sample_p, 3.0)
        input_ids, draft_tokens, retrieve_indices, tree_mask, tree_pos, new_token, _, token = out
        torch.cuda.synchronize()
        # bookkeeping checks
        assert input_ids.shape[1] == prev_len + a + 1
        assert int(cur_len[0]) == prev_len + a + 1
        kept = m.kv[0, 0, 0, prev_len:prev_len + a + 1, 0].float().cpu().numpy()
        sel_prev = retrieve_prev[int(best), :a + 1].cpu().numpy() + prev_len
        assert kept.tolist() == [float(x) for x in sel_prev]                        # accepted slots, in path order
        assert token.shape == (1, 1) and 0 <= int(token) < V and float(sample_p[int(token)]) > 0
    assert new_token == input_ids.shape[1] - 17 and m.ea_layer.calls == n_steps + 1
