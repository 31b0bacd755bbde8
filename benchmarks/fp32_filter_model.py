"""CPU model of the fp32 filter of visibility_sorted.cu (numpy float32, non-fused arithmetic).

Checks on a synthetic scene that every (point, view) pair the filter DECIDES agrees with the exact
fp64 evaluation (the oracle's arithmetic), and reports the share of undecided pairs that go to the
exact fp64 queue. Development aid only: run `python benchmarks/fp32_filter_model.py`."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dropclip_b200.scenes import make_scene

EPS = np.float64(2.0 ** -24)
C_ERR = 8.0  # roundings charged per dot product (inputs 2, products/adds <= 5, fp64 slack)
f32 = np.float32


def run(seed=1234, V=16, N=20000, jitter=0.0):
    sc = make_scene(seed, n_views=V, n_points=N, n_objects=21, device="cpu")
    H, W = sc.intrinsic["height"], sc.intrinsic["width"]
    K = np.array([[sc.intrinsic["fx"], 0, sc.intrinsic["cx"]], [0, sc.intrinsic["fy"], sc.intrinsic["cy"]], [0, 0, 1]], np.float64)
    pts = np.asarray(sc.points, np.float64)
    B = np.abs(pts).max(axis=0)
    und = dec = bad = 0
    for v in range(V):
        inv = np.linalg.inv(np.asarray(sc.camera_poses[v])).astype(np.float32).astype(np.float64)
        M = inv[:3].copy(); M[1:] *= -1
        # exact (fp64) reference values
        c = pts @ M[:, :3].T + M[:, 3]
        q = c @ K.T
        with np.errstate(all="ignore"):
            u = q[:, 0] / q[:, 2]; w = q[:, 1] / q[:, 2]
        inside = (u > -1) & (u < W) & (w > -1) & (w < H)
        pu = np.where(inside, np.trunc(u), 0).astype(np.int64); pv = np.where(inside, np.trunc(w), 0).astype(np.int64)
        d = sc.depths[v][pv, pu].astype(np.float64)
        vis = inside & (np.abs(d - q[:, 2]) <= 0.05)
        # fp32 filter
        P = K @ M  # fp64
        Pf = P.astype(f32)
        A = np.stack([np.abs(K[0, 0]) * np.abs(M[0]) + np.abs(K[0, 2]) * np.abs(M[2]),
                      np.abs(K[1, 1]) * np.abs(M[1]) + np.abs(K[1, 2]) * np.abs(M[2]), np.abs(M[2])])  # no cancellation
        S = A[:, :3] @ B + A[:, 3]
        E = (C_ERR * EPS * S * (1 + 2.0 ** -20))
        Ex, Ey, Ez = E
        pf = pts.astype(f32)
        qf = [(Pf[r, 0] * pf[:, 0] + (Pf[r, 1] * pf[:, 1] + (Pf[r, 2] * pf[:, 2] + Pf[r, 3]))).astype(f32) for r in range(3)]
        with np.errstate(all="ignore"):
            r = (f32(1) / qf[2]).astype(f32)
        uf = (qf[0] * r).astype(f32); vf = (qf[1] * r).astype(f32)
        # per-axis constants (visibility_sorted.cu: camera_prep_kernel)
        cu, hu = (W - 1) / 2.0, (W - 1) / 2.0
        cv, hv = (H - 1) / 2.0, (H - 1) / 2.0
        NLu, NLv = hu + 2 + 0.0014 * W, hv + 2 + 0.0014 * H
        Uu, Uv = cu + NLu + 1, cv + NLv + 1
        Au = 1.02 * (Ex + 1.03 * Uu * Ez); Av = 1.02 * (Ey + 1.03 * Uv * Ez)
        H0u = 0.499999 - 1.02 * 5 * EPS * Uu; H0v = 0.499999 - 1.02 * 5 * EPS * Uv
        zmin = max(1024 * Ez, 4 * Ex, 4 * Ey)
        hu_ = (f32(H0u) - f32(Au) * r).astype(f32); hv_ = (f32(H0v) - f32(Av) * r).astype(f32)
        sane = qf[2] >= f32(zmin)
        fu = np.floor(uf); fv = np.floor(vf)
        gu = ((uf - np.abs(fu)) - f32(0.5)).astype(f32); gv = ((vf - np.abs(fv)) - f32(0.5)).astype(f32)
        au = np.abs(fu - f32(cu)); av = np.abs(fv - f32(cv))
        notnear = (au > f32(NLu)) | (av > f32(NLv))
        dec_uv = sane & (notnear | ((np.abs(gu) < hu_) & (np.abs(gv) < hv_)))
        in_f = (au <= f32(hu)) & (av <= f32(hv))
        pu_f = np.clip(fu, 0, W - 1).astype(np.int64); pv_f = np.clip(fv, 0, H - 1).astype(np.int64)
        df = sc.depths[v][pv_f, pu_f]
        delta = np.abs(df - qf[2]).astype(f32)
        thr_lo = f32((0.05 - Ez) * (1 - 4 * EPS)); thr_hi = f32((0.05 + Ez) * (1 + 4 * EPS))
        vis_yes = delta <= thr_lo; vis_no = delta > thr_hi
        decided = dec_uv & (~in_f | vis_yes | vis_no)
        res = dec_uv & in_f & vis_yes
        # checks
        pix_ok = (~dec_uv) | (in_f == inside) & ((~in_f) | ((pu_f == pu) & (pv_f == pv)))
        bad += int((~pix_ok).sum()) + int((decided & (res != vis)).sum())
        und += int((~decided).sum()); dec += int(decided.sum())
    return und, dec, bad


if __name__ == "__main__":
    for seed in (1234, 1235):
        und, dec, bad = run(seed)
        print(f"seed {seed}: undecided {und} ({100.0 * und / (und + dec):.3f} %), decided {dec}, decided-but-wrong {bad}")
