import sys, time, torch
sys.path.insert(0, "/root/repo")
from phyloligo_b200 import engine
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.rand((100000, 256), device="cuda", generator=g) ** 2
X /= X.sum(dim=1, keepdim=True)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); P, aux, dim = engine.prepare(X, "JSD"); e1.record(); torch.cuda.synchronize()
    print("prepare JSD 100k x 256: device %.3f ms, wall %.3f ms" % (e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
