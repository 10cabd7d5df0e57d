"""Tuning probe: time the all-groups and the per-image metric kernels over batch sizes (run once per cluster-size build)."""
import sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
from polcue import ops, synth
dev = torch.device("cuda", 0)
base = synth.gen_depth_batch(0, 8, 320, 480)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev).view(torch.int32)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    def run(body):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): body()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    tf = min(run(lambda: flush.sum()) for _ in range(2))
    return min(run(lambda: (flush.sum(), fn())) for _ in range(2)) - tf
out = {}
for B in (15, 30, 60, 120, 240, 480, 960, 1920):
    reps_ = (B + 7) // 8
    gt = torch.from_numpy(base[0]).to(dev).repeat(reps_, 1, 1)[:B].contiguous()
    pred = torch.from_numpy(base[1]).to(dev).repeat(reps_, 1, 1)[:B].contiguous()
    inst = torch.from_numpy(base[2]).to(dev).repeat(reps_, 1, 1)[:B].contiguous()
    g = timeit(lambda: ops.depth_errors_groups(gt, pred, inst, 0.1, 2.0, [None] + list(synth.MATERIAL_LEVELS)))
    i = timeit(lambda: ops.depth_errors_per_image(gt, pred, 0.1, 2.0, inst, 40))
    out[B] = (round(g * 1e3, 1), round(i * 1e3, 1))
print(json.dumps(out))
