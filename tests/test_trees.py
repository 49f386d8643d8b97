"""Host-side tree logic of the product (lantern_b200.trees) against the reference fixtures."""
import json
import os

import pytest
import torch

from lantern_b200 import choices as CH
from lantern_b200 import trees

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "tree_buffers.json")) as f:
    TREES = json.load(f)


@pytest.mark.parametrize("name", CH.NAMES + CH.SYNTH_NAMES)
def test_generate_tree_buffers_matches_reference(name):
    ref = TREES[name]
    tb = trees.generate_tree_buffers(CH.tree(name), device="cpu")
    assert tb["tree_indices"].tolist() == ref["tree_indices"]
    assert tb["tree_position_ids"].tolist() == ref["tree_position_ids"]
    assert tb["retrieve_indices"].tolist() == ref["retrieve_indices"]
    assert tb["tree_attn_mask"][0, 0].long().tolist() == ref["tree_attn_mask"]
    assert tb["p_indices"] == ref["p_indices"]
    got_b = [[(x.tolist() if isinstance(x, torch.Tensor) else list(x)) for x in row] for row in tb["b_indices"]]
    assert got_b == ref["b_indices"]


@pytest.mark.parametrize("name", CH.NAMES + CH.SYNTH_NAMES)
def test_static_tree_csr(name):
    tb = trees.generate_tree_buffers(CH.tree(name), device="cpu")
    st = tb["static_tree"]
    ri = tb["retrieve_indices"]
    T = tb["tree_indices"].shape[0]
    assert st.sib_off.shape[0] == T + 1
    # CSR siblings == b_indices, op row == level offset + p_indices
    counts = tb["group_counts"]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    assert st.n_q_rows == offs[-1]
    for j in range(ri.shape[0]):
        for i in range(1, ri.shape[1]):
            v = int(ri[j, i])
            if v < 0:
                continue
            sib = st.sib_idx[int(st.sib_off[v]):int(st.sib_off[v + 1])].tolist()
            b = tb["b_indices"][j][i]
            assert sib == (b.tolist() if isinstance(b, torch.Tensor) else list(b))
            assert int(st.node_qrow[v]) == offs[i - 1] + tb["p_indices"][j][i]


def test_generate_candidates_shapes():
    tb = trees.generate_tree_buffers(CH.tree("mc_sim_7b_63"), device="cpu")
    n_groups = tb["static_tree"].n_q_rows
    ss_token = torch.arange(n_groups * 10).view(n_groups, 10) + 100
    ss_prob = torch.rand(n_groups, 10)
    cart, cart_prob, tc = trees.generate_candidates((ss_token, ss_prob, None), tb["tree_indices"],
                                                    tb["retrieve_indices"], torch.tensor([[7]]))
    assert cart.shape == tb["retrieve_indices"].shape and tc.shape == (1, 26)
    assert int(tc[0, 0]) == 7 and float(cart_prob[0, 0]) == 1.0
    assert (cart[tb["retrieve_indices"] == -1] == -1).all()
