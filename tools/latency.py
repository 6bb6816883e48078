#!/usr/bin/env python
"""Latency of the host-buffer evaluation (vmat_eval) under different host-path variants."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import workload
from vmatwpencode_b200.plan import DosePlan

case, y, w_dose, w_grad = workload("cfg2", 1.0)
Dt = case.D.tocsc()
VARIANTS = [("graph: H2D node, K3 writes results + tags to host memory", dict(host_path=2)),
            ("graph: H2D node, D2H node, stream sync", dict(host_path=8)),
            ("no graph: y in launch params, D2H copy, stream sync", dict(host_path=4)),
            ("no graph: y in launch params, tags", dict(host_path=6)),
            ("no graph: H2D copy, D2H copy, stream sync", dict(host_path=8, use_graph=False))]
for name, kw in [v for _ in range(3) for v in VARIANTS]:
    plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, Dt=Dt, **kw)
    plan.set_weights(w_dose, w_grad)
    for _ in range(50):
        plan.eval(y)
    n = 3000
    t0 = time.perf_counter()
    for k in range(n):
        plan.eval(y if k & 1 else y * 1.0000001)
    dt = (time.perf_counter() - t0) / n
    print(json.dumps(dict(variant=name, us_per_eval=round(dt * 1e6, 2))), flush=True)
    plan.close()
