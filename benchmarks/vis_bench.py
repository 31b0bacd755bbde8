"""Times the visibility pipeline alone (sorted filter vs direct exact kernel) at the bench extents.
`python benchmarks/vis_bench.py [scenes] [views] [points]`; checks bit-equality of the two kernels first."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dropclip_b200.engine import FusionEngine, batch_from_device
from dropclip_b200.scenes import make_scene

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
V = int(sys.argv[2]) if len(sys.argv) > 2 else 73
N = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
dev = torch.device("cuda")
eng = FusionEngine(dev)
uniq = [make_scene(1234 + i, n_views=V, n_points=N, n_objects=21, device="cuda", as_torch=True) for i in range(min(S, 8))]
b = batch_from_device([uniq[i % len(uniq)] for i in range(S)], dev)
torch.cuda.synchronize()
direct, any_d, _ = eng.visibility(b, 0.05, torch.uint8)
rec, rank, any_s = eng.visibility_sorted(b, 0.05)
same = bool(torch.equal(eng.unpack_visibility(b, rec, rank, torch.uint8), direct)) and bool(torch.equal(any_s, any_d))
del direct


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / n


t_sorted = timeit(lambda: eng.visibility_sorted(b, 0.05))
pairs = sum(n * v for n, v in zip(b.n_points, b.n_views))
alg = sum(24 * n + v * n * 5 for n, v in zip(b.n_points, b.n_views))
print(json.dumps({"bit_equal_to_direct": same, "scenes": S, "views": V, "points": N, "sorted_ms": t_sorted,
                  "pairs_per_s": pairs / t_sorted * 1e3, "alg_GBps": alg / t_sorted / 1e6}))
