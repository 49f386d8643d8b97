import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lantern_b200 import codebook
N, d, K = (int(x) for x in sys.argv[1:4])
E = torch.nn.functional.normalize(torch.randn(N, d, device="cuda"), dim=1)
for _ in range(2):
    codebook.build_neighbor_table(E, K)
torch.cuda.synchronize()
