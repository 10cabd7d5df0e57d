import sys, torch, time
sys.path.insert(0,'supervised-depth-estimation-from-polarized-images_b200')
from polcue import ops, synth
dev=torch.device('cuda',0)
m=synth.gen_p_batch_torch(0,64,device=dev)
out={}
for stats in (False, True):
    for _ in range(5): out=ops.fused_mosaic(m,1.5,out=out,want_stats=stats)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): out=ops.fused_mosaic(m,1.5,out=out,want_stats=stats)
    b.record(); torch.cuda.synchronize()
    print("stats" if stats else "plain", a.elapsed_time(b)/50, "ms")
