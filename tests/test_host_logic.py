"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol, struct layouts
match the header, LogitsWarp follows the HF warper semantics the oracle restates."""
import ctypes as C
import os
import re

import numpy as np
import torch

from lantern_b200 import _abi
from lantern_b200 import posterior as PO
from oracle import lantern_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lantern_b200.h")).read()
    declared = set(re.findall(r"LANTERN_API\s+[\w\s\*]+?\b(lantern_\w+)\s*\(", hdr))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    lib = _abi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lantern_version() >> 16 == 2     # LANTERN_ABI_VERSION (2: caller-owned scratch for lantern_build_neighbors)
    assert lib.lantern_last_error() is not None


def test_struct_layouts_match_header(tmp_path):
    """Compile the header with gcc and compare sizeof / offsetof with the ctypes mirrors."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    fields = {"lantern_accept_cfg": ("AcceptCfg", ["n_items", "item_stride", "family", "cfg_scale", "lantern_delta_m1",
                                                    "syntax_tokens", "newline_token", "n_q_rows", "bonus_uniform_last", "philox_seed", "philox_step"]),
              "lantern_accept_in": ("AcceptIn", ["logits_cond", "nbr_table", "sib_tokens_stride"]),
              "lantern_accept_out": ("AcceptOut", ["accept_length", "sample_p"]),
              "lantern_kv_cfg": ("KvCfg", ["n_slabs", "n_outer", "n_batch", "max_keep"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/lantern_b200.h"', "int main(void){"]
    for cname, (_, fl) in fields.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fl:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).strip().splitlines())
    for cname, (pyname, fl) in fields.items():
        st = getattr(_abi, pyname)
        assert int(got[cname]) == C.sizeof(st), cname
        for f in fl:
            assert int(got[f"{cname}.{f}"]) == getattr(st, f).offset, f"{cname}.{f}"


def test_workspace_query_and_argument_validation_without_gpu():
    lib = _abi.load()
    cfg = _abi.AcceptCfg()
    cfg.n_items, cfg.n_rows = 3, 59
    assert lib.lantern_accept_workspace_bytes(C.byref(cfg)) == (3 * 59 * 32 + 255) & ~255      # 32 B per row, 256-byte granules
    rc = lib.lantern_accept_fused(C.byref(cfg), C.byref(_abi.AcceptIn()), C.byref(_abi.AcceptOut()), None, 0, None)
    assert rc == _abi.E_INVALID and b"lantern_accept_fused" in lib.lantern_last_error()
    out = np.zeros(8, dtype=np.float32)
    lib.lantern_philox_uniforms(1234, 7, 3, 8, out.ctypes.data)
    assert np.array_equal(out, O.philox_uniforms(1234, 7, 3, 8))


def test_logits_warp_matches_oracle_warp():
    rng = np.random.default_rng(0)
    row = (rng.standard_normal(4096) * 2.5).astype(np.float32)
    for t, p, k in [(1.0, 1.0, 0), (0.8, 1.0, 100), (1.3, 0.9, 0), (1.0, 0.5, 300), (1.0, 1.0, 5000)]:
        proc = PO.prepare_logits_processor(temperature=t, top_p=p, top_k=k)
        got = proc(None, torch.from_numpy(row)[None])[0].numpy()
        want = O.Warp(t, p, k)(row)
        assert np.array_equal(np.isneginf(got), np.isneginf(want))
        keep = ~np.isneginf(want)
        assert np.allclose(got[keep], want[keep], rtol=0, atol=0)
    assert len(PO.prepare_logits_processor(temperature=0.0, top_k=100)) == 0


def test_lumina_row_kinds():
    pos = torch.arange(10, 10 + 49 * 48 + 5)
    got = PO.lumina_row_kinds(pos, 7).numpy()
    want = O.lumina_row_kinds(pos.numpy(), 7)
    assert np.array_equal(got, want) and (got == 1).sum() >= 47 and (got == 2).sum() == 1


def test_every_exported_entry_is_documented():
    """INTEGRATION.md names every LANTERN_API symbol of the header (the table a reference maintainer binds against)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "lantern_b200.h")).read()
    names = set(re.findall(r"LANTERN_API\s+[\w\s\*]+?\b(lantern_\w+)\s*\(", header))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    assert names and not [n for n in names if n not in doc]


def test_dropin_signatures_match_reference():
    """the shims keep the reference's call signatures.  tests/golden/signatures.json holds the parameter
    lists of the live reference methods (written by inspect in the build container); every reference parameter must
    exist here under the same name, in the same relative order, and the required positional prefix must be identical."""
    import inspect
    import json
    from lantern_b200 import trees
    here = os.path.dirname(os.path.abspath(__file__))
    ref = json.load(open(os.path.join(here, "golden", "signatures.json")))
    ours = {
        "utils.tree_decoding": PO.tree_decoding, "utils.evaluate_posterior": PO.evaluate_posterior,
        "utils.update_inference_inputs": PO.update_inference_inputs,
        "utils.prepare_logits_processor": PO.prepare_logits_processor,
        "utils.generate_tree_buffers": trees.generate_tree_buffers,
        "lumina.EaLumina_mGPT.tree_decoding": PO.LuminaVerifyMixin.tree_decoding,
        "lumina.EaLumina_mGPT.evaluate_posterior": PO.LuminaVerifyMixin.evaluate_posterior,
        "lumina.EaLumina_mGPT.update_inference_inputs": PO.LuminaVerifyMixin.update_inference_inputs,
    }
    for fam in ("llamagen", "anole"):
        for m in ("tree_decoding", "evaluate_posterior", "evaluate_posterior_v1", "update_inference_inputs"):
            ours[f"{fam}.EaModel.{m}"] = getattr(PO.VerifyMixin, m)
    assert set(ours) == set(ref)
    for name, f in ours.items():
        mine = list(inspect.signature(f).parameters)
        theirs = [p for p, _ in ref[name]]
        required = [p for p, d in ref[name] if d is None]
        assert mine[:len(required)] == required, (name, mine, required)
        pos = [mine.index(p) for p in theirs]                 # raises if a reference parameter is missing
        assert pos == sorted(pos), (name, mine, theirs)


def test_missing_library_fails_loudly(tmp_path):
    """No CPU fallback: with the library path pointing at nothing (``LANTERN_B200_LIB``) the binding raises."""
    import subprocess
    import sys
    code = "from lantern_b200 import _abi; _abi.load()"
    env = dict(os.environ, LANTERN_B200_LIB=str(tmp_path / "absent.so"), PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and "absent.so" in r.stderr
