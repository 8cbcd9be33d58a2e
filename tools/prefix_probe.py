import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, slamklt
from slamklt import synth
from oracle import oracle as O
ctx = slamklt.Context(0)
for (H, W) in ((376, 1241), (1080, 1920)):
    fr, _ = synth.make_sequence(2003, 1, H=H, W=W)
    img = synth.to_f64(fr)[0]
    levels = 3 if H < 1000 else 5
    op = O.LKPyramid(img, levels, mode="ctor"); op.update(img)
    gp = slamklt.LKPyramid(ctx, img, levels); gp.update(img)
    for l in range(levels + 1):
        for rname, sname in (("Ryy", "Syy"), ("Rxx", "Sxx"), ("Ryx", "Syx")):
            R = gp.plane(l, rname); S = op.plane(l, sname)
            Hl, Wl = S.shape
            ref = np.concatenate([np.zeros((Hl, 1)), np.cumsum(S, axis=1)], axis=1)
            k = min(19, Wl)
            rd = R[:, k:] - R[:, :-k]; rr = ref[:, k:] - ref[:, :-k]
            absS = np.concatenate([np.zeros((Hl, 1)), np.cumsum(np.abs(S), axis=1)], axis=1)
            kc = min(40, Wl)
            cm = np.max(absS[:, kc:] - absS[:, :-kc], axis=1, keepdims=True)
            rowmax = np.max(np.abs(ref), axis=1, keepdims=True)
            err = np.abs(rd - rr)
            e_abs = err - 1e-5 * np.abs(rr)
            kr = min(19, Hl)
            cd = np.concatenate([np.zeros((1, rd.shape[1])), np.cumsum(rd, axis=0)]); cr = np.concatenate([np.zeros((1, rr.shape[1])), np.cumsum(rr, axis=0)])
            gd, gr = cd[kr:] - cd[:-kr], cr[kr:] - cr[:-kr]
            gerr = np.abs(gd - gr)
            print(H, l, rname, "prefix abs err max %.3e  / rowmax %.3e /(rowmax+cm) %.3e | row-sum excess/(rowmax) %.3e /(rowmax+cm) %.3e | G: max-norm %.3e  frac(1e-5,1e-6) %.5f frac(1e-5,1e-7) %.5f" % (
                np.max(np.abs(R - ref)), np.max(np.abs(R - ref) / rowmax), np.max(np.abs(R - ref) / (rowmax + cm)),
                np.max(e_abs / rowmax), np.max(e_abs / (rowmax + cm)),
                gerr.max() / np.abs(gr).max(), np.mean(gerr <= 1e-5 * np.abs(gr) + 1e-6 * np.abs(gr).max()), np.mean(gerr <= 1e-5 * np.abs(gr) + 1e-7 * np.abs(gr).max())))
        for name in ("layer", "Iy", "Ix", "Syy", "Sxx", "Syx"):
            a, b = gp.plane(l, name), op.plane(l, name)
            e = np.abs(a - b)
            print("   ", H, l, name, "max-norm %.3e frac(1e-5,1e-6) %.5f frac(1e-5,1e-7) %.5f" % (e.max() / np.abs(b).max(), np.mean(e <= 1e-5 * np.abs(b) + 1e-6 * np.abs(b).max()), np.mean(e <= 1e-5 * np.abs(b) + 1e-7 * np.abs(b).max())))
