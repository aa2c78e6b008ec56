"""Deterministic synthetic descriptor sets for the feature-matching path.

Host-side NumPy only (Generator(PCG64(seed)), seed = 20070000 + config#).  This is the input
producer for tests and bench.py -- it stands where PP/featureMatching/getFeaturePoints.m:32-74
(closed-source extractors) stands in the reference -- and is not part of the matching path.

  float "SIFT-like": |N(0,1)| -> L2 normalise -> clip 0.2 -> renormalise -> round(512 x) clipped to
      255, stored as float32 (integer-valued 0..255 like OpenCV / VLFeat SIFT, getFeaturePoints.m:42-43)
  float "KAZE-like": N(0,1) L2-normalised float32 (real-valued)
  binary "ORB-like": uniform random bytes [N x nb]
  planted overlaps : images on rings; image i receives noisy copies of 30% of its rows from image
      i-1 and 15% from image i-2 (float: + N(0,0.02^2) per dim before quantisation; binary: each bit
      flipped w.p. 0.05); 8 exact duplicates across two images and 8 inside one image.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # id: (kind, n_images, kp_per_image, D_or_nb, ring_size, empty_images, k, ratio, m)
    1: dict(kind="sift", n=6, kp=2048, D=128, ring=6, empty=(3,), k=4, ratio=0.6, m=4,
            name="C1: 6 img x 2048 kp SIFT-128 f32, global k=4 (one image empty)"),
    2: dict(kind="sift", n=20, kp=8192, D=128, ring=20, empty=(), k=4, ratio=0.8, m=4,
            name="C2: 20 img x 8192 kp SIFT-128 f32, global k=4, ratio 0.8"),
    3: dict(kind="sift", n=100, kp=10000, D=128, ring=100, empty=(), k=4, ratio=0.8, m=4,
            name="C3: 100 img x 10000 kp SIFT-128 f32, global k=4"),
    4: dict(kind="orb", n=50, kp=20000, D=32, ring=50, empty=(), k=4, ratio=0.8, m=4,
            name="C4: 50 img x 20000 kp ORB-256bit, BF Hamming k=4"),
    5: dict(kind="kaze", n=300, kp=4096, D=64, ring=30, empty=(), k=4, ratio=0.8, m=4,
            name="C5: 300 img x 4096 kp KAZE-64 f32, pairwise blocks"),
    # not a BASELINE config: C2 with REAL-valued SIFT-like descriptors (stress test of the bf16 error bound)
    6: dict(kind="siftf", n=20, kp=8192, D=128, ring=20, empty=(), k=4, ratio=0.8, m=4,
            name="C2f: 20 img x 8192 kp real-valued SIFT-128 f32, global k=4"),
}


def _unit(x):
    return x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)


def _sift_quantise(v, integer=True):
    v = np.clip(v, 0.0, None)
    v = _unit(v)
    v = np.minimum(v, 0.2)
    v = _unit(v)
    if not integer:   # "siftf": the same descriptor left real-valued (unit norm), as MATLAB's extractFeatures returns it
        return v.astype(np.float32)
    return np.minimum(np.rint(512.0 * v), 255.0).astype(np.float32)


def synth_descriptors(kind, n, kp, D, seed, ring=None, empty=(), duplicates=8):
    """Returns a list of n arrays ([Ni x D] float32, or [Ni x D] uint8 bytes for kind='orb')."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ring = ring or n
    counts = [0 if i in empty else kp for i in range(n)]
    live = [i for i in range(n) if counts[i] > 0]
    n30, n15 = int(0.30 * kp), int(0.15 * kp)
    fresh0 = n30 + n15  # rows [fresh0, kp) of every image are i.i.d. fresh rows
    base = {}
    for i in live:
        if kind == "orb":
            base[i] = rng.integers(0, 256, size=(kp, D), dtype=np.uint8)
        elif kind in ("sift", "siftf"):
            base[i] = _unit(np.abs(rng.standard_normal((kp, D), dtype=np.float32)))
        else:
            base[i] = _unit(rng.standard_normal((kp, D), dtype=np.float32))
    # rings over the live images, in order
    for r0 in range(0, len(live), ring):
        members = live[r0:r0 + ring]
        L = len(members)
        if L < 2:
            continue
        for pos, i in enumerate(members):
            for back, lo, cnt in ((1, 0, n30), (2, n30, n15)):
                if L <= back or cnt == 0 or kp - fresh0 < cnt:
                    continue
                src_img = members[(pos - back) % L]
                if src_img == i:
                    continue
                src_rows = fresh0 + rng.permutation(kp - fresh0)[:cnt]
                src = base[src_img][src_rows]
                if kind == "orb":
                    flips = rng.random((cnt, D * 8)) < 0.05
                    base[i][lo:lo + cnt] = src ^ np.packbits(flips, axis=1)
                else:
                    base[i][lo:lo + cnt] = src + rng.normal(0.0, 0.02, size=src.shape).astype(np.float32)
    out = []
    for i in range(n):
        if counts[i] == 0:
            out.append(np.zeros((0, D), np.uint8 if kind == "orb" else np.float32))
        elif kind == "orb":
            out.append(base[i])
        elif kind in ("sift", "siftf"):
            out.append(_sift_quantise(base[i], integer=(kind == "sift")))
        else:
            out.append(_unit(base[i]).astype(np.float32))
    # exact duplicates: across two images, and inside one image
    if duplicates and len(live) >= 2 and kp >= fresh0 + 4 * duplicates:
        a, b = live[0], live[1]
        out[b][kp - duplicates:kp] = out[a][kp - 2 * duplicates:kp - duplicates]
        c = live[2] if len(live) >= 3 else live[0]
        out[c][kp - 4 * duplicates:kp - 3 * duplicates] = out[c][kp - 3 * duplicates:kp - 2 * duplicates]
    return out


def make_config(cid, n=None, kp=None):
    """Descriptor set + parameters of BASELINE.json config `cid` (optionally down-sized via n / kp)."""
    c = dict(CONFIGS[cid])
    if n is not None:
        c["n"] = n
        c["ring"] = min(c["ring"], n)
        c["empty"] = tuple(e for e in c["empty"] if e < n)
    if kp is not None:
        c["kp"] = kp
    desc = synth_descriptors(c["kind"], c["n"], c["kp"], c["D"], 20070000 + cid, ring=c["ring"], empty=c["empty"])
    return desc, c


def synth_matched_keypoints(n, kp, seed, per_pair=(0.25, 0.10), inlier_frac=(0.7, 0.5), noise=0.6, size=2000.0,
                            stray=12):
    """Synthetic input of the consumer stage (imageMatching.m): keypoints [kp x 2] per image and the n x n
    nested list of putative match rows ([M x 2] float64, 1-based) of images on a ring.  Pair (i, i+1) gets
    per_pair[0]*kp matches of which inlier_frac[0] follow a random mild homography of image i+1 -> image i
    (+ gaussian pixel noise), pair (i, i+2) per_pair[1]*kp with inlier_frac[1]; every other pair gets `stray`
    random rows.  Returns (keypoints, matchesAll, truth) with truth[(i, j)] = H mapping image j points to image i."""
    rng = np.random.Generator(np.random.PCG64(seed))
    keypoints = [rng.uniform(0.0, size, (kp, 2)) for _ in range(n)]
    matches = [[np.zeros((0, 0)) for _ in range(n)] for _ in range(n)]
    truth = {}
    for i in range(n):
        for j in range(i + 1, n):
            gap = min(j - i, n - (j - i))
            if gap in (1, 2) and n > gap:
                m = int(per_pair[gap - 1] * kp)
                H = np.eye(3) + rng.normal(0.0, 1.0, (3, 3)) * np.array([[0.06, 0.06, 80.0], [0.06, 0.06, 80.0],
                                                                           [3e-5, 3e-5, 0.0]])
                ri = rng.permutation(kp)[:m]
                rj = rng.permutation(kp)[:m]
                good = rng.random(m) < inlier_frac[gap - 1]
                # move the matched keypoints of image i onto H * (keypoints of image j) for the inliers
                pj = np.c_[keypoints[j][rj], np.ones(m)]
                q = (H @ pj.T).T
                q = q[:, :2] / q[:, 2:3] + rng.normal(0.0, noise, (m, 2))
                keypoints[i][ri[good]] = q[good]
                matches[i][j] = np.c_[ri + 1, rj + 1].astype(np.float64)
                truth[(i, j)] = H
            elif stray:
                matches[i][j] = np.c_[rng.integers(1, kp + 1, stray), rng.integers(1, kp + 1, stray)].astype(np.float64)
    return keypoints, matches, truth
