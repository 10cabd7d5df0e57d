import sys, time, json, torch
sys.path.insert(0,'supervised-depth-estimation-from-polarized-images_b200')
from polcue import ops, synth
dev=torch.device('cuda',0)
B,H,W=64,2048,2448
mosaic=synth.gen_p_batch_torch(0,B,H,W,device=dev)
h_m=ops.host_empty((B,H,W),torch.uint8,dev); h_m.copy_(mosaic)
out={"xolp":ops.host_empty((B,2,H//2,W//2),torch.float32,dev),"normals":ops.host_empty((B,9,H//2,W//2),torch.float32,dev)}
res={}
for chunk in (0,1,2,4,8,16):
    ops.fused_mosaic_host(h_m,1.5,out=out,chunk_frames=chunk)
    torch.cuda.synchronize()
    best=1e9
    for _ in range(4):
        t0=time.perf_counter(); ops.fused_mosaic_host(h_m,1.5,out=out,chunk_frames=chunk); best=min(best,time.perf_counter()-t0)
    res[chunk]=best*1e3
print(json.dumps(res))
