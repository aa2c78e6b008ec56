"""Independent numpy/LAPACK restatement of estimateTransformationRANSAC.m ('projective') used ONLY to pin
oracle/aps_oracle_ransac.c: it follows the MATLAB text with numpy.linalg (svd / solve / cond = LAPACK, the
library MATLAB calls for the same operations).  Minimal samples are an input table, like the oracle's."""
import numpy as np

EPS = np.finfo(np.float64).eps


def normalize_points(pts):  # :574-596
    centroid = pts.mean(axis=0)
    pc = pts - centroid
    scale = 1.0 / np.mean(np.sqrt((pc ** 2).sum(axis=1)))
    T = np.array([[scale, 0, -scale * centroid[0]], [0, scale, -scale * centroid[1]], [0, 0, 1.0]])
    ph = np.c_[pts, np.ones(len(pts))]
    return (T @ ph.T).T[:, :2], T


def estimate_homography(p1, p2):  # :188-225
    a, T1 = normalize_points(p1)
    b, T2 = normalize_points(p2)
    n = len(a)
    x, y, u, v = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    o, z = np.ones(n), np.zeros((n, 3))
    A = np.r_[np.c_[-x, -y, -o, z, x * u, y * u, u], np.c_[z, -x, -y, -o, x * v, y * v, v]]
    _, _, Vt = np.linalg.svd(A)  # full: V is 9 x 9, V(:,end) = last row of Vt
    Hn = Vt[-1].reshape(3, 3)
    return np.linalg.solve(T2, Hn / Hn[2, 2]) @ T1


def check_model(H):  # :518-530
    if not np.all(np.isfinite(H)):
        return False
    with np.errstate(all="ignore"):
        rc = 1.0 / np.linalg.cond(H, 1)
    return bool(rc > EPS and abs(np.linalg.det(H)) > EPS)


def is_degenerate(points):  # :532-572
    if len(points) < 3:
        return True
    s = np.linalg.svd(points - points.mean(axis=0), compute_uv=False)
    with np.errstate(all="ignore"):
        return bool(s[1] / s[0] < 1e-3)


def find_inliers(H, ph1, ph2, thr):  # :444-516
    with np.errstate(all="ignore"):
        t = (H @ ph1.T).T
        t = t / t[:, 2:3]
        iv = np.linalg.solve(H, ph2.T).T
        iv = iv / iv[:, 2:3]
        d1 = ((ph2[:, :2] - t[:, :2]) ** 2).sum(axis=1)
        d2 = ((ph1[:, :2] - iv[:, :2]) ** 2).sum(axis=1)
        err = np.sqrt(d1 + d2)
        err[~np.isfinite(err)] = np.inf
        err[np.abs(t[:, 2]) < EPS] = np.inf
    inl = err < thr
    if inl.sum() >= 4 and is_degenerate(ph1[inl, :2]):
        inl[:] = False
        err[:] = np.inf
    return inl, err


def ransac(p1, p2, max_distance, confidence, max_trials, samples):  # :94-183
    n = len(p1)
    if n < 4:
        return False, None, np.zeros(n, bool), 0
    ph1, ph2 = np.c_[p1, np.ones(n)], np.c_[p2, np.ones(n)]
    best_inl, best_model, best_err = np.zeros(n, bool), None, np.inf
    trial, skip, d, max_skip = 1, 0, 0, max_trials * 10
    while trial <= max_trials and skip < max_skip and d < len(samples):
        s = samples[d]
        d += 1
        try:
            with np.errstate(all="ignore"):
                H = estimate_homography(p1[s], p2[s])
            if not check_model(H):
                skip += 1
                continue
        except np.linalg.LinAlgError:
            skip += 1
            continue
        inl, err = find_inliers(H, ph1, ph2, max_distance)
        ni = int(inl.sum())
        if ni >= 4:
            me = err[inl].mean()
            if ni > best_inl.sum() or (ni == best_inl.sum() and me < best_err):
                best_inl, best_model, best_err = inl, H, me
                ratio = ni / n
                if ratio > 0:
                    with np.errstate(all="ignore"):
                        cand = np.ceil(np.log(1 - confidence / 100) / np.log(1 - ratio ** 4))
                    if cand < max_trials:
                        max_trials = cand
        trial += 1
    if best_inl.sum() >= 4:
        model = estimate_homography(p1[best_inl], p2[best_inl])
        if check_model(model):
            inl, _ = find_inliers(model, ph1, ph2, max_distance)
            if inl.sum() >= 4:
                return True, model, inl, d
        return True, best_model, best_inl, d
    return False, best_model, best_inl, d


def make_pair(rng, n, inlier_frac, noise=0.7, size=1000.0):
    """n correspondences: inlier_frac of them follow a random mild homography (+ gaussian noise), the rest are random."""
    H = np.eye(3) + rng.normal(0, 1, (3, 3)) * np.array([[0.08, 0.08, 60.0], [0.08, 0.08, 60.0], [4e-5, 4e-5, 0.0]])
    p1 = rng.uniform(0, size, (n, 2))
    q = (H @ np.c_[p1, np.ones(n)].T).T
    p2 = q[:, :2] / q[:, 2:3] + rng.normal(0, noise, (n, 2))
    out = rng.random(n) >= inlier_frac
    p2[out] = rng.uniform(0, size, (int(out.sum()), 2))
    return p1, p2, H


def draw_table(rng, n, n_draws):
    """n_draws x 4 distinct zero-based indices (stand-in for randperm(numPoints, 4))."""
    if n < 4:
        return np.zeros((n_draws, 4), np.uint32)
    return np.stack([rng.choice(n, 4, replace=False) for _ in range(n_draws)]).astype(np.uint32)
