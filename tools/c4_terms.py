"""Prints the four data-dependent terms of the 2-D bound at the C4 fixture (tests/scale_cases.py) from the GPU path, with
full digits, for comparison with oracle/extended_check.py (long double) and the fp64 LAPACK oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scale_cases as SC
from test_gpu_scale import _model_2d

model, X, y = _model_2d(SC.C4)
e, g = model.elbo_and_grad()
t = dict(model.last_terms)
t["elbo"] = e
t["tr_yTy"] = model.tr_yTy
t["grad"] = [float(g[id(p)]) for p in model.trainable_variables]
print(json.dumps({k: (float(v) if not isinstance(v, list) else v) for k, v in t.items()}))
if "--dump" in sys.argv:
    alpha, SigP, S1, S2, info = model.posterior_weights()
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "r2_c4_dump.npz"), Gs=model._Gs.cpu().numpy(), b=model._b.cpu().numpy(),
                        S1=S1.cpu().numpy(), S2=S2.cpu().numpy(), alpha=alpha.cpu().numpy(),
                        SigP_cols=SigP.cpu().numpy()[:, ::97])
