"""One warm-up and one measured launch of the fused kernel, plain and with the statistics by-product (for ncu captures)."""
import sys, torch
sys.path.insert(0, 'supervised-depth-estimation-from-polarized-images_b200')
from polcue import ops, synth
dev = torch.device('cuda', 0)
m = synth.gen_p_batch_torch(0, 64, device=dev)
out = {}
for stats in (False, False, True, True):
    out = ops.fused_mosaic(m, 1.5, out=out, want_stats=stats)
torch.cuda.synchronize()
print("ok")
