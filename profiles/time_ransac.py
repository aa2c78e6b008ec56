"""Times the RANSAC consumer (aps_image_matching_batch) on one GPU next to the oracle on the host cores.
usage: time_ransac.py [n_images] [kp] [--no-cpu]   (C2-like default: 20 images x 8192 keypoints on a ring)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if len(args) > 0 else 20
kp = int(args[1]) if len(args) > 1 else 8192
keypoints, matches, _ = pkg.synth.synth_matched_keypoints(n, kp, seed=77)
inp = {"maxDistance": 5.5, "inliersConfidence": 99.9, "maxIter": 500, "mBrownLowe": 6}
ctx = pkg._lib.default_context()
for it in range(3):
    t0 = time.perf_counter()
    allM, numM, tf = pkg.imageMatching(inp, n, keypoints, matches, seed=1)
    dt = time.perf_counter() - t0
    last = pkg.imageMatching.last
    P = len(last["pairs_lin"])
    total = int(last["pt_ptr"][-1])
    evals = float(np.sum(np.diff(last["pt_ptr"]) * 1000.0))
    print(f"it{it}: {P} candidate pairs, {total} correspondences, 1000 draws/pair evaluated: {dt*1e3:.2f} ms end to end "
          f"({evals/dt:.3e} trial-correspondence evaluations/s), accepted {int(last['accepted'].sum())}, "
          f"draws consumed median {int(np.median(last['draws_used']))}")
if "--no-cpu" not in sys.argv:
    from oracle import oracle as orc
    lin, ptr = last["pairs_lin"], last["pt_ptr"]
    P1 = np.vstack([keypoints[c // n][np.asarray(matches[c % n][c // n], np.int64)[:, 1] - 1] for c in lin])
    P2 = np.vstack([keypoints[c % n][np.asarray(matches[c % n][c // n], np.int64)[:, 0] - 1] for c in lin])
    tab = pkg.ransacSampleTable(ptr, 1000, seed=1)
    t0 = time.perf_counter()
    o = orc.image_matching_batch(ptr, P1, P2, 5.5, 99.9, 500, tab)
    dt = time.perf_counter() - t0
    same = np.array_equal(o["inliers"], last["inliers"]) and np.array_equal(o["accepted"], last["accepted"])
    used = float(np.sum(np.diff(ptr) * o["draws_used"]))
    print(f"oracle (sequential loop, stops at the adaptive bound; {orc.num_threads()} threads over pairs): {dt*1e3:.1f} ms "
          f"({used/dt:.3e} evaluations/s on the draws it consumed); identical result: {same}")
