#!/usr/bin/env python
"""Latency of the host-buffer evaluation (vmat_eval) under different host-path variants."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import workload
from vmatwpencode_b200.plan import DosePlan

case, y, w_dose, w_grad = workload("cfg2", 1.0)
Dt = case.D.tocsc()
for name, kw in [("tags + y in launch params", dict(host_path=6)), ("tags + H2D node (default)", dict(host_path=2)), ("zero-copy both", dict(host_path=3)), ("zero-copy y only", dict(host_path=1)),
                 ("zero-copy result only", dict(host_path=2)), ("memcpy nodes", dict(host_path=4)),
                 ("memcpy nodes, no graph", dict(host_path=4, use_graph=False)), ("zero-copy both, no graph", dict(host_path=3, use_graph=False)),
                 ("memcpy nodes, no pdl", dict(host_path=4, pdl=False))]:
    plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, Dt=Dt, **kw)
    plan.set_weights(w_dose, w_grad)
    for _ in range(50):
        plan.eval(y)
    n = 2000
    t0 = time.perf_counter()
    for k in range(n):
        plan.eval(y if k & 1 else y * 1.0000001)
    dt = (time.perf_counter() - t0) / n
    print(json.dumps(dict(variant=name, us_per_eval=round(dt * 1e6, 2))), flush=True)
    plan.close()
