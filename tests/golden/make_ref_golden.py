"""Generates tests/golden/ref_golden.json from oracle/_ref/libuba_ref.so — the reference's own BundleAdjuster.h /
rotation_utils.cpp / StereoVisualOdometry.cpp compiled in this container (make -C oracle ref; needs /root/reference).
The fixture lets the pinning tests run where the reference tree and the compiled library are absent.

    python tests/golden/make_ref_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ref_binding as rb  # noqa: E402
from uasl_motion_estimation_b200 import capi, synth  # noqa: E402


def main():
    G = json.loads((ROOT / "tests" / "golden" / "residual_golden.json").read_text())
    k = capi.Calib()
    for n, v in G["calib"].items():
        setattr(k, n, v)
    out = {"note": "outputs of the reference's own sources (oracle/_ref), hex floats are exact", "calib": G["calib"], "functor": [], "quat": [], "ba": []}
    hx = lambda a: [float(x).hex() for x in np.asarray(a).reshape(-1)]
    rng = np.random.default_rng(7)
    cases = [dict(M=c["M"], cam=c["cam"], X=c["X"], obs=c["obs"], cam_id=c["cam_id"]) for c in G["cases"]]
    for i in range(40):   # more poses, including tiny rotations on both sides of the theta^2 = eps switch
        M = 4 if i % 2 == 0 else 2
        ang = [1e-9, 1.2e-8, 2e-8, 1e-3, 0.3, 2.5][i % 6]
        r = rng.normal(size=3); r *= ang / np.linalg.norm(r)
        cam = np.concatenate([rng.normal(size=3) * 0.5, r])
        X = np.array([rng.uniform(-8, 8), rng.uniform(-3, 3), rng.uniform(4, 60)])
        obs = np.float32(rng.uniform(20, 1200, size=M)).astype(np.float64)
        cases.append(dict(M=M, cam=cam.tolist(), X=X.tolist(), obs=obs.tolist(), cam_id=int(i % 4 == 3)))
    for c in cases:
        r, Jc, Jp = rb.residual(c["M"], k, c["cam"], c["X"], c["obs"], c["cam_id"])
        out["functor"].append(dict(c, r=hx(r), Jc=hx(Jc), Jp=hx(Jp)))
    for i in range(24):
        q = rng.normal(size=4)
        if i < 4:
            q = np.array([1.0, 0, 0, 0]) + (0 if i == 0 else 1e-9 * rng.normal(size=4))
        rv = rb.log_map(q)
        out["quat"].append(dict(q=hx(q), log=hx(rv), exp_of_log=hx(rb.exp_map(rv))))
    for name, scale, M, fixed in (("c1", 0.02, 4, 2), ("c2", 0.003, 4, 2), ("c1", 0.02, 2, 1)):
        win = synth.config_window(name, scale=scale, M=M)
        o = rb.ba_run(win, first_frame=100, fixed_frames=fixed)
        out["ba"].append(dict(config=name, scale=scale, M=M, fixed_frames=fixed, first_frame=100, status=o["status"], n_obs=o["n_obs"],
                              cam_idx=o["cam_idx"].tolist(), pt_idx=o["pt_idx"].tolist(), cam_id=o["cam_id"].tolist(),
                              cams_init=hx(o["cams_init"]), cams=hx(o["cams"]), pts=hx(o["pts"]), quat_id=hx(o["quat_id"])))
    (ROOT / "tests" / "golden" / "ref_golden.json").write_text(json.dumps(out))
    print("wrote ref_golden.json:", len(out["functor"]), "functor cases,", len(out["quat"]), "quaternions,", len(out["ba"]), "BA runs")


if __name__ == "__main__":
    main()
