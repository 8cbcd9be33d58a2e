"""Generates tests/golden/oracle_small.npz from the CPU oracle (the reference is Julia and cannot run in this
image, so these are regression pins of the restatement, not outputs of the real reference -- DESIGN.md).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from slamklt import synth  # noqa: E402

fr, aff = synth.make_sequence(4242, 2, H=96, W=128)
f = synth.to_f64(fr)
pts = synth.random_keypoints(7, 48, 96, 128, border=2.0)
p0 = O.LKPyramid(f[0], 2, mode="ctor")
p1 = O.LKPyramid(f[1], 2, mode="ctor")
p1.update(f[1])
new, st, fst = O.fb_tracking(p0, p1, pts, window_size=9, pyramid_levels=2, max_distance=1.0)
kp = O.detect(O.Extractor(120, 8, (3, 4), 35), f[0], pts[:10])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz"),
                    img0=f[0], img1=f[1], frames_u8=fr, pts=pts, new=new, status=st, fstatus=fst, kp=kp,
                    p0_layer2=p0.plane(2, "layer"), p1_Sxx1=p1.plane(1, "Sxx"),
                    p1_layer1=p1.plane(1, "layer"), p1_Iy0=p1.plane(0, "Iy"))
# camera geometry + optical_flow_matching! composition (map_manager.jl:451-564) on the same small pair
gt = synth.true_flow(aff, 0, 1, pts)
sc = synth.matching_scene(11, pts, gt, camera=dict(synth.KITTI_CAMERA, cx=64.0, cy=48.0, fx=120.0, fy=120.0, height=96, width=128),
                          frac_bad=0.2, frac_outside=0.1)
cam = O.Camera(**sc["camera"])
m_pix, m_und, m_pos, m_st = O.optical_flow_matching(p0, p1, pts, sc["is_3d"], sc["world"], None, sc["cw"], cam,
                                                    window_size=9, pyramid_levels=2, max_distance=1.0)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_matching.npz"),
                    is_3d=sc["is_3d"], world=sc["world"], cw=sc["cw"], camera=np.array([sc["camera"][k] for k in
                    ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "height", "width")], dtype=np.float64),
                    proj=O.project_world_distort(cam, sc["cw"], sc["world"]), undist=O.undistort_point(cam, pts),
                    pix=m_pix, und=m_und, pos=m_pos, status=m_st)
print("wrote oracle_matching.npz: status histogram", np.unique(m_st, return_counts=True))
print("wrote oracle_small.npz", st.sum(), "/", len(st), "tracked;", len(kp), "keypoints")
