"""compute-sanitizer target for the latency path of predict alone (multi-right-hand-side wavefront kernel, vector_kernels.cuh):
a small fit, then predict_mean_variance for q = 5 and q = 16 against the batched path's results."""
import sys

import numpy as np

sys.path.insert(0, ".")
import friedrich_b200 as F  # noqa: E402
from friedrich_b200.synthetic import make_dataset, make_inputs  # noqa: E402

n, d = 900, 4
X, y = make_dataset(11, n, d)
gp = F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(0.8, 1.0), 0.1, None, X, y)
Xq = make_inputs(12, 40, d)
mb, vb = gp.predict_mean_variance(Xq)          # batched path (q = 40)
for q in (5, 16):
    m, v = gp.predict_mean_variance(Xq[:q])    # one multi-right-hand-side wavefront launch
    print("q", q, "max |dmean|", float(np.max(np.abs(m - mb[:q]))), "max |dvar|", float(np.max(np.abs(v - vb[:q]))))
