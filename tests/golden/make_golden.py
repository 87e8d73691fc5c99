"""Generates tests/golden/residual_golden.json: 40-digit mpmath evaluation of the reference's
residual functors (BundleAdjuster.h:78-94, :113-130, :153-171) and of their Jacobians (mpmath
numerical differentiation at 40 digits) on seeded random inputs.  Independent of both the oracle
and the CUDA path: it is written straight from the reference's formulas.  Run: python make_golden.py"""
import json
import random
from pathlib import Path

import mpmath as mp

mp.mp.dps = 40
K = dict(fx0=718.856, fy0=718.856, cx0=607.1928, cy0=185.2157, fx1=718.856, cx1=607.1928, baseline=0.537, feat_var=0.25)


def rotate(r, p):
    r = [mp.mpf(x) for x in r]; p = [mp.mpf(x) for x in p]
    th2 = sum(x * x for x in r)
    if th2 > mp.mpf(2) ** -52:  # DBL_EPSILON
        th = mp.sqrt(th2); c = mp.cos(th); s = mp.sin(th)
        w = [x / th for x in r]
        wxp = [w[1] * p[2] - w[2] * p[1], w[2] * p[0] - w[0] * p[2], w[0] * p[1] - w[1] * p[0]]
        tmp = (w[0] * p[0] + w[1] * p[1] + w[2] * p[2]) * (1 - c)
        return [p[i] * c + wxp[i] * s + w[i] * tmp for i in range(3)]
    rxp = [r[1] * p[2] - r[2] * p[1], r[2] * p[0] - r[0] * p[2], r[0] * p[1] - r[1] * p[0]]
    return [p[i] + rxp[i] for i in range(3)]


def residual(M, cam_id, x, obs):
    cam, X = x[:6], x[6:]
    p = rotate(cam[3:], X)
    si = 1 / mp.sqrt(mp.mpf(K["feat_var"]))
    if M == 4:
        p = [p[0] + cam[0], p[1] + cam[1], p[2] + cam[2]]
        x1 = K["fx0"] * (p[0] / p[2]) + K["cx0"]
        x2 = K["fx1"] * ((p[0] - K["baseline"]) / p[2]) + K["cx1"]
        y = K["fy0"] * (p[1] / p[2]) + K["cy0"]
        return [si * (x1 - obs[0]), si * (y - obs[1]), si * (x2 - obs[2]), si * (y - obs[3])]
    px = p[0] + cam[0] - (K["baseline"] if cam_id else 0)
    p = [px, p[1] + cam[1], p[2] + cam[2]]
    return [si * (K["fx0"] * (p[0] / p[2]) + K["cx0"] - obs[0]), si * (K["fy0"] * (p[1] / p[2]) + K["cy0"] - obs[1])]


def main():
    rng = random.Random(20261018)
    cases = []
    for i in range(24):
        M = 4 if i % 3 else 2
        cam_id = (i // 3) % 2 if M == 2 else 0
        scale = [0.3, 1e-3, 0.0, 1e-9][i % 4]  # includes the exact-zero and below-epsilon rotation branches
        cam = [rng.uniform(-1, 1), rng.uniform(-0.5, 0.5), rng.uniform(-2, 2)] + [scale * rng.uniform(-1, 1) for _ in range(3)]
        X = [rng.uniform(-8, 8), rng.uniform(-3, 3), rng.uniform(6, 50)]
        obs = [rng.uniform(0, 1241), rng.uniform(0, 376), rng.uniform(0, 1241), rng.uniform(0, 376)][:M]
        obs = [float(mp.mpf(o)) for o in obs]
        x0 = [mp.mpf(v) for v in cam + X]
        r = residual(M, cam_id, x0, obs)
        J = [[mp.diff(lambda *a, m=m: residual(M, cam_id, list(a), obs)[m], tuple(x0), tuple(int(k == j) for k in range(9))) for j in range(9)] for m in range(M)]
        cases.append(dict(M=M, cam_id=cam_id, cam=cam, X=X, obs=obs, r=[float(v) for v in r], J=[[float(v) for v in row] for row in J]))
    out = Path(__file__).resolve().parent / "residual_golden.json"
    out.write_text(json.dumps(dict(calib=K, cases=cases), indent=1))
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
