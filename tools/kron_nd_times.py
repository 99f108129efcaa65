"""One 2-D ELBO + gradient evaluation at 200 x 200 (k = 3, bench hypers) on a small raster, repeated; meant to be run under
`ncu --metrics gpu__time_duration.sum` to list the per-launch times of the nested-dissection kernels level by level."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scale_cases as SC
from test_gpu_scale import _model_2d

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
case = dict(SC.C4)
if len(sys.argv) > 2:
    case["m"] = (int(sys.argv[2]), int(sys.argv[2]))
if len(sys.argv) > 3:
    case["order"] = int(sys.argv[3])
model, X, y = _model_2d(case)
for _ in range(reps):
    e, g = model.elbo_and_grad()
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    e, g = model.elbo_and_grad()
t1.record(); torch.cuda.synchronize()
print("elbo", e, "ms per elbo_and_grad", t0.elapsed_time(t1) / 5)
