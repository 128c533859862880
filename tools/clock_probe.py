#!/usr/bin/env python
"""Is the rollout kernel power/clock limited in steady state?  Per-block kernel time for a long back-to-back run
(with nvidia-smi SM clock and power sampled every 20 ms) vs isolated launches separated by idle gaps."""
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygent_b200.examples import make_ilqr  # noqa: E402

B = 65536
L = make_ilqr("cart_pole", 10, x0=np.zeros((2, 4))).lib
x0 = torch.zeros((4, B), dtype=torch.float64, device="cuda")
x0[1] = np.pi
U = (torch.rand((500, 1, B), dtype=torch.float64, device="cuda") * 2 - 1) * 10
for _ in range(3):
    L.rollout(x0, U, S=4, dt=0.02)
torch.cuda.synchronize()

lines = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "20"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()


def block(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        L.rollout(x0, U, S=4, dt=0.02)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


time.sleep(1.0)
print("isolated launches (200 ms idle between):", ["%.4f" % (time.sleep(0.2) or block(1)) for _ in range(8)], flush=True)
t0 = time.time()
res = []
for i in range(30):
    res.append((time.time() - t0, block(100)))
print("back-to-back blocks of 100 launches: t[s] -> ms/launch")
print("  " + ", ".join("%.2f:%.4f" % r for r in res), flush=True)
t1 = time.time()
time.sleep(0.3)
proc.terminate()
busy = [l for t, l in lines if t0 <= t <= t1]
print("nvidia-smi during the run (sm MHz, W, reasons): first/median/last of %d samples" % len(busy))
if busy:
    mhz = [float(l.split(",")[0]) for l in busy]
    w = [float(l.split(",")[1]) for l in busy]
    print("  sm MHz min/median/max %.0f/%.0f/%.0f; power W median/max %.0f/%.0f; reasons %s" % (
        min(mhz), np.median(mhz), max(mhz), np.median(w), max(w), sorted({l.split(",")[2].strip() for l in busy})))
