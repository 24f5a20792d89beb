"""Time po_matrix_knn on a resident float32 matrix (uniform random and a JSD-shaped one)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phyloligo_b200 import phyloselect

rows, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
D = torch.rand((rows, n), device="cuda", dtype=torch.float32)
for tag, M in (("uniform", D), ("squared", D * D)):
    for _ in range(2):
        phyloselect.knn_graph(M, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        phyloselect.knn_graph(M, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("knn %s rows=%d cols=%d k=%d: %.2f ms  %.0f GB/s (one pass of the matrix)" % (tag, rows, n, k, ms, rows * n * 4 / ms / 1e6))
