"""Times the RANSAC consumer (aps_image_matching) on one GPU.  The CPU oracle is timed beside it, on the same input,
by tests/test_gpu_ransac.py::test_c2_like_scale_and_timing (only tests may load oracle/).
usage: time_ransac.py [n_images] [kp]   (C2-like default: 20 images x 8192 keypoints on a ring)"""
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
