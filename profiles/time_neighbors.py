"""Time the neighbour-table build (tensor-core path vs all-fp64 path) at the BASELINE configs[4] sizes."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lantern_b200 import codebook
res = {}
for name, N, d, K in [("llamagen_16384x8_k1001", 16384, 8, 1001), ("chameleon_8192x256_k1001", 8192, 256, 1001),
                      ("llamagen_16384x8_full", 16384, 8, 16383)]:
    E = torch.nn.functional.normalize(torch.randn(N, d, device="cuda"), dim=1)
    for mode in ("tensor_core", "exact"):
        if mode == "exact":
            os.environ["LANTERN_NBR_EXACT_ONLY"] = "1"
        else:
            os.environ.pop("LANTERN_NBR_EXACT_ONLY", None)
        _, route = codebook.build_neighbor_table(E, K, return_route=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            codebook.build_neighbor_table(E, K)
        torch.cuda.synchronize()
        res[f"{name}_{mode}_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        res[f"{name}_{mode}_route"] = route.tolist()
    if N * N * 4 < 2**31:
        def ref():
            dist = torch.cdist(E, E); dist.fill_diagonal_(float("inf")); return torch.topk(dist, K, largest=False)
        ref(); ref(); torch.cuda.synchronize()      # warmed: the first call pays cuBLAS / allocator start-up
        t0 = time.perf_counter()
        for _ in range(3):
            ref()
        torch.cuda.synchronize()
        res[f"{name}_torch_gpu_cdist_topk_ms_warmed"] = (time.perf_counter() - t0) / 3 * 1e3
print(json.dumps(res, indent=1))
