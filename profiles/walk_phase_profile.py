"""Cycle breakdown of walk_kernel phases for item 0 (needs a build with -DLANTERN_WALK_PROFILE: set
LANTERN_EXTRA_NVCC_FLAGS=-DLANTERN_WALK_PROFILE before python -m lantern_b200.build --force)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import casegen as CG, cuda_runner as R
from lantern_b200 import _abi
lib = _abi.load()
names = ["prologue", "child list", "level distribution", "relaxation", "rejection", "tail distribution", "bonus token"]
fam = sys.argv[1] if len(sys.argv) > 1 else "lumina_mgpt"
tot = np.zeros(16)
n = 0
for seed in range(200, 216):
    b = CG.build(dict(family=fam, seed=seed, depth=5 if fam == "lumina_mgpt" else 4))
    buf = (C.c_ulonglong * 16)()
    R.run_cases([b], want_sample_p=False)
    lib.lantern_debug_walk_profile(buf, 1)
    R.run_cases([b], want_sample_p=False)
    lib.lantern_debug_walk_profile(buf, 1)
    tot += np.array(list(buf), dtype=np.float64)
    n += 1
tot /= n
print(f"{fam}: mean cycles per item (1 item per launch), total {tot[:7].sum():.0f} cycles = {tot[:7].sum()/1.965e3:.1f} us")
for i, nm in enumerate(names):
    print(f"  {nm:20s} {tot[i]:9.0f} cycles  {tot[i]/tot[:7].sum()*100:5.1f} %")
