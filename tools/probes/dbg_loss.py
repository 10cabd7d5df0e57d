import sys, numpy as np, torch
sys.path.insert(0,'supervised-depth-estimation-from-polarized-images_b200'); sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from polcue import ops, synth
def mis(t):
    flat=torch.empty(t.numel()+1,dtype=t.dtype,device=t.device); v=flat[1:].view(t.shape); v.copy_(t); return v
h,w=64,128
for holes in (False, True):
    gt,_,_,k=synth.gen_depth_batch(21,1,h,w)
    if holes: gt=synth.add_hole_regions(gt,21)
    else: gt=np.where(gt>0,gt,0.7).astype(np.float32)
    vv,uu=np.mgrid[0:h,0:w].astype(np.float32)
    pred=(np.where(gt>0,gt,0.7)*(1.0+0.05*np.sin(uu/23.0)*np.cos(vv/17.0))).astype(np.float32)
    g=torch.from_numpy(gt).cuda()[:,None]; pr=torch.from_numpy(pred).cuda()[:,None]; kk=torch.from_numpy(k).cuda()
    ng=ops.depth_to_normals(g,kk).double(); npd=ops.depth_to_normals(pr,kk).double()
    cos=(ng*npd).sum(1)/torch.clamp((ng.norm(dim=1)*npd.norm(dim=1)),min=1e-8)
    per_px_ref=(2-cos)[0]
    print("holes",holes,"ref mean",per_px_ref.mean().item())
    for name,gg in (("packed",g),("scalar",mis(g))):
        full=torch.ones_like(g)
        print(" ",name, ops.normals_loss(gg,pr,kk,full).item())
    # per pixel values at a few pixels
    for (y,x) in ((10,10),(0,0),(31,64),(63,127),(20,5),(40,100)):
        m=torch.zeros_like(g); m[0,0,y,x]=1
        a=ops.normals_loss(g,pr,kk,m).item(); b=ops.normals_loss(mis(g),pr,kk,m).item()
        print("   px",(y,x),"packed",a,"scalar",b,"ref",per_px_ref[y,x].item(), "gt",gt[0,y,x])
