#!/usr/bin/env python
"""Time of one k-medoid insertion step (all n candidates against all n points, fp64) on the GPU
for voxel-like point sets.  Prints one JSON line per size."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmatwpencode_b200.plan import kmedoid_costs


def main():
    rng = np.random.default_rng(0)
    for n in (16384, 65536, 262144):
        P = np.round(rng.random((n, 3)) * 128)
        cur = np.sqrt(((P - P[0]) ** 2).sum(axis=1))
        for label, cm in (("one medoid", cur), ("64 medoids", None)):
            if cm is None:
                cm = cur.copy()
                for k in rng.integers(0, n, 63):
                    cm = np.minimum(cm, np.sqrt(((P - P[k]) ** 2).sum(axis=1)))
            kmedoid_costs(P[:4096 + 8], cm[:4096 + 8])          # warm-up (context, module load)
            t0 = time.perf_counter()
            cost = kmedoid_costs(P, cm)
            dt = time.perf_counter() - t0
            print(json.dumps(dict(n=n, state=label, seconds=round(dt, 4), pairs_per_s=round(n * n / dt / 1e9, 2),
                                  unit="1e9 distance pairs/s", argmin=int(np.argmin(cost)))))


if __name__ == "__main__":
    main()
