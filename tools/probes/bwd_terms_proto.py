"""CPU prototype (numpy float64) of the normals-loss backward in the cancellation-free form used by csrc/normals_loss.cu:
forward m = f_x f_y (gu x gv) from the six window functionals (G, V, A, B, Cu, Cv), their adjoints, and the gather with the
replicate-padding folds -- checked against float64 autograd of the reference formulation (oracle.normals_loss_torch).
  python tools/probes/bwd_terms_proto.py      (prints differences at the 1e-15 level; images with H = 1 or W = 1 are degenerate: m = 0, the clamped branch of the kernels)
"""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import polcue_oracle as O

def fwd_terms(Z, K):
    H, W = Z.shape
    fX, fY, cx, cy = K[0,0], K[1,1], K[0,2], K[1,2]
    Zp = np.pad(Z, 1, mode='edge')
    sh = lambda a, b: Zp[1+a:H+1+a, 1+b:W+1+b]
    x = np.arange(W)[None, :]; y = np.arange(H)[:, None]
    wl, wr = (x > 0)*1.0, (x < W-1)*1.0
    wm, wp = (y > 0)*1.0, (y < H-1)*1.0
    S2 = {b: sh(-1,b) + 2*sh(0,b) + sh(1,b) for b in (-1,0,1)}
    D2 = {b: sh(1,b) - sh(-1,b) for b in (-1,0,1)}
    T = {b: wp*sh(1,b) + wm*sh(-1,b) for b in (-1,0,1)}
    D = {b: wp*sh(1,b) - wm*sh(-1,b) for b in (-1,0,1)}
    G = S2[1] - S2[-1]; V = D2[-1] + 2*D2[0] + D2[1]
    A = wr*S2[1] + wl*S2[-1]; B = T[-1] + 2*T[0] + T[1]
    Cu = D[1] - D[-1]; Cv = wr*D2[1] - wl*D2[-1]
    fx0 = (x - cx)/fX; fy0 = (y - cy)/fY
    mx = fX*(V*Cu - G*B); my = fY*(G*Cv - V*A); mz = -fx0*mx - fy0*my + A*B - Cu*Cv
    return dict(G=G,V=V,A=A,B=B,Cu=Cu,Cv=Cv,fx0=fx0,fy0=fy0,m=np.stack([mx,my,mz]),fX=fX,fY=fY,wl=wl,wr=wr,wm=wm,wp=wp)

def loss_and_grad(Zg, Zq, K, mask):
    H, W = Zq.shape
    tg, tp = fwd_terms(Zg, K), fwd_terms(Zq, K)
    a = tg['m'] / np.linalg.norm(tg['m'], axis=0)
    nm = np.linalg.norm(tp['m'], axis=0); mh = tp['m']/nm
    c = (a*mh).sum(0)
    M = mask.sum()
    loss = ((2-c)*mask).sum()/M
    mbar = -(mask/M) * (a - c*mh)/nm                  # dL/dm
    mzb = mbar[2]; mxp = mbar[0] - tp['fx0']*mzb; myp = mbar[1] - tp['fy0']*mzb
    px, py = tp['fX']*mxp, tp['fY']*myp
    G,V,A,B,Cu,Cv = (tp[k] for k in ('G','V','A','B','Cu','Cv'))
    Gb = py*Cv - px*B; Vb = px*Cu - py*A; Ab = mzb*B - py*V; Bb = mzb*A - px*G; Cub = px*V - mzb*Cv; Cvb = py*G - mzb*Cu
    # gather form: F = 0 outside the image
    pad0 = lambda F: np.pad(F, 1)                      # zeros
    def rows(F, kind):                                 # vertical combination at output row q: F(q-1), F(q), F(q+1)
        P = pad0(F); up, mid, dn = P[0:H,1:W+1], P[1:H+1,1:W+1], P[2:H+2,1:W+1]
        y = np.arange(H)[:, None]; first, last = (y == 0)*1.0, (y == H-1)*1.0
        if kind == 's': return up + dn + (2 + first + last)*mid
        if kind == 'd': return up - dn + (last - first)*mid
        if kind == 'e': return up + dn
        if kind == "d'": return up - dn
    def cols(F, kind):
        P = pad0(F); lf, mid, rt = P[1:H+1,0:W], P[1:H+1,1:W+1], P[1:H+1,2:W+2]
        x = np.arange(W)[None, :]; first, last = (x == 0)*1.0, (x == W-1)*1.0
        if kind == 's': return lf + rt + (2 + first + last)*mid
        if kind == 'd': return lf - rt + (last - first)*mid
        if kind == 'e': return lf + rt
        if kind == "d'": return lf - rt
    Zbar = (cols(rows(Gb,'s'),'d') + cols(rows(Vb,'d'),'s') + cols(rows(Ab,'s'),'e') + cols(rows(Bb,'e'),'s')
            + cols(rows(Cub,"d'"),'d') + cols(rows(Cvb,'d'),"d'"))
    return loss, Zbar

def compare_with_autograd(H, W, seed=0):
    """(loss difference, max gradient difference relative to the largest gradient) between the six-functional form and
    float64 autograd of the reference formulation, on a random smooth surface with a random mask."""
    rng = np.random.default_rng(seed)
    K = np.array([[520.0*W/640, 0, W/2 - 0.3], [0, 515.0*H/480, H/2 + 0.2], [0, 0, 1]])
    yy, xx = np.mgrid[0:H, 0:W]
    Zg = 0.8 + 0.1*np.sin(xx/3.0) + 0.05*np.cos(yy/2.0) + 0.01*rng.standard_normal((H, W))
    Zq = Zg*(1 + 0.05*rng.standard_normal((H, W)))
    mask = (rng.random((H, W)) > 0.2)*1.0
    loss, grad = loss_and_grad(Zg, Zq, K, mask)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))[None, None]
    zq = t(Zq).clone().requires_grad_(True)
    ref = O.normals_loss_torch(t(Zg), zq, torch.from_numpy(K)[None], t(mask))
    ref.backward()
    g_ref = zq.grad[0, 0].numpy()
    return abs(loss - float(ref.detach())), np.abs(grad - g_ref).max() / (np.abs(g_ref).max() + 1e-30)


if __name__ == "__main__":
    for (H, W) in ((7, 9), (2, 2), (12, 16), (3, 4), (2, 9), (9, 2)):
        print((H, W), "loss diff %.1e, gradient diff %.1e of its scale" % compare_with_autograd(H, W))
