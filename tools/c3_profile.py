"""evaluation.test at the OpenEA-15K test size (10,500 x 10,500, d = 256) in the default fp32-accurate mode, for ncu captures of the
3xTF32 similarity kernel (three passes: row top-k, column top-k, CSLS rank count)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jmac_b200 import finding as F
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
e1, e2 = bench.planted_pair(10_500, 256, 4.0, g, dev)
for _ in range(2):
    ranks, top1 = F.alignment_ranks(e1, e2, metric="cosine", csls_k=10)
torch.cuda.synchronize()
print("hits@1", float((ranks < 1).float().mean()))
