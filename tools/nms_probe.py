"""GPU probe: vilco batched_nms vs the golden vectors from the reference extension (tests/golden/nms.npz) and vs the
C oracle on larger random inputs."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mq_oracle as O  # noqa: E402
from oracle import nms_c  # noqa: E402
from vilco_b200.utils.nms import batched_nms  # noqa: E402

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "nms.npz"))
bad = 0
for name in ["n1", "n17", "n300", "n2000", "ties"]:
    segs, scores = torch.from_numpy(g[name + "_segs"]), torch.from_numpy(g[name + "_scores"])
    labels = torch.zeros(len(scores), dtype=torch.int64)
    for sigma, ms in ((0.99, 1e-4), (0.5, 0.01)):
        tag = f"{name}_m2_s{sigma}_t{ms}"
        dets = g[tag + "_dets"]
        k = min(200, len(dets))
        order = np.argsort(-dets[:k, 2], kind="stable")
        s, sc, lb = batched_nms(segs, scores, labels, 0.1, ms, 200, True, True, sigma)
        ok = s.shape[0] == k and (s.numpy() == dets[:k][order][:, :2]).all() and (sc.numpy() == dets[:k][order][:, 2]).all()
        bad += 0 if ok else 1
        print("OK " if ok else "BAD", tag, s.shape[0], k, flush=True)
        if not ok and s.shape[0] == k:
            print("  max score diff", np.abs(sc.numpy() - dets[:k][order][:, 2]).max(), "seg mism", (s.numpy() != dets[:k][order][:, :2]).sum())
s, sc, lb = batched_nms(torch.from_numpy(g["b_segs"]), torch.from_numpy(g["b_scores"]), torch.from_numpy(g["b_labels"]), 0.1, 1e-4, 200, True, True, 0.99)
ok = (s.numpy() == g["b_out_segs"]).all() and (sc.numpy() == g["b_out_scores"]).all() and (lb.numpy() == g["b_out_labels"]).all()
bad += 0 if ok else 1
print("OK " if ok else "BAD", "batched multi-class golden", flush=True)
# larger random: 28k candidates, 110 classes vs the C oracle
rs = np.random.RandomState(5)
for n, K in ((20000, 22), (28000, 110), (50000, 1)):
    centre = rs.uniform(0, 1024, n).astype(np.float32)
    length = np.exp(rs.uniform(np.log(2.0), np.log(400.0), n)).astype(np.float32)
    segs = torch.from_numpy(np.stack([centre - length / 2, centre + length / 2], 1).astype(np.float32))
    scores = torch.from_numpy(rs.beta(0.5, 8, n).astype(np.float32))
    labels = torch.from_numpy(rs.randint(0, K, n).astype(np.int64))
    t0 = time.time()
    rs_, rsc, rl = O.batched_nms(segs, scores, labels, 0.1, 1e-4, 200, True, True, 0.99, softnms_fn=nms_c.softnms_1d)
    t1 = time.time()
    s, sc, lb = batched_nms(segs, scores, labels, 0.1, 1e-4, 200, True, True, 0.99)
    torch.cuda.synchronize()
    t2 = time.time()
    s, sc, lb = batched_nms(segs, scores, labels, 0.1, 1e-4, 200, True, True, 0.99)
    torch.cuda.synchronize()
    t3 = time.time()
    ok = (s.numpy() == rs_.numpy()).all() and (sc.numpy() == rsc.numpy()).all() and (lb.numpy() == rl.numpy()).all()
    bad += 0 if ok else 1
    print("OK " if ok else "BAD", f"random n={n} K={K}: cpu oracle {1e3*(t1-t0):.1f} ms, gpu (incl. copies) {1e3*(t3-t2):.2f} ms", flush=True)
sys.exit(1 if bad else 0)
