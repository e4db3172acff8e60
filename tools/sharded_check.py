"""Multi-GPU parity check, run under torchrun on the GPU box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_check.py [n] [d]
Every rank runs the collective fgp_fit_sharded; rank 0 also runs the single-GPU fgp_fit and (n <= 4096) the CPU oracle.
Prints one JSON line per rank: the sharded factor must equal the single-GPU factor bit for bit on every rank."""
import ctypes as C
import json
import math
import os
import sys

import numpy as np
import torch  # noqa: F401  (first: see csrc/nccl_dyn.cuh)
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from friedrich_b200 import _native as N  # noqa: E402
from friedrich_b200 import sharded  # noqa: E402
from friedrich_b200.kernels import SquaredExp  # noqa: E402
from friedrich_b200.synthetic import make_dataset, make_inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
X, y = make_dataset(0x5EED0004, n, d)
Xq = make_inputs(0x5EED0005, 256, d)
ls = math.sqrt(d / 6.0)
kd = SquaredExp(ls, 1.0).device_desc()
h = N.Handle(local)
sharded.comm_init(h, rank, world, dist)
if rank == 0:
    sharded.fit_sharded(h, X, y, kd, 0.1)
else:
    sharded.fit_sharded(h, (n, d), None, kd, 0.1)  # receives X, y from rank 0
ms = h.last_device_ms()
L = np.zeros((n, n), order="F")
h.check(N.lib().fgp_download_factor(h.ptr, N.dptr(L), n))
mean, var = np.zeros(256), np.zeros(256)
h.check(N.lib().fgp_predict_mean_var(h.ptr, C.byref(kd), N.dptr(N.fcol(Xq)), 256, 256, N.dptr(mean), N.dptr(var)))
digest = float(np.sum(np.tril(L) * np.linspace(1.0, 2.0, n)[:, None]))
# the LML gradient, collectively (scaled and unscaled), timed on the second call
for _ in range(2):
    gscale, ggrads = sharded.lml_gradient_sharded(h, kd, 0.1, 2, scaled=True)
grad_ms = h.last_device_ms()
_, ggrads_u = sharded.lml_gradient_sharded(h, kd, 0.1, 2, scaled=False)
out = {"rank": rank, "world": world, "n": n, "d": d, "fit_device_ms": ms, "lml_scale": gscale, "lml_grads": ggrads, "bcast_mb": N.lib().fgp_comm_last_bytes(h.ptr) / 1e6,
       "digest": digest}
if rank == 0:
    hp = N.Handle(local)
    hp.check(N.lib().fgp_fit(hp.ptr, N.dptr(N.fcol(X)), n, n, d, N.dptr(y), C.byref(kd), 0.1, 0, 0.0))
    Lp = np.zeros((n, n), order="F")
    hp.check(N.lib().fgp_download_factor(hp.ptr, N.dptr(Lp), n))
    out["bitwise_equal_single_gpu"] = bool(np.array_equal(np.tril(L), np.tril(Lp)))
    out["single_gpu_ms"] = hp.last_device_ms()
    g1 = np.zeros(3)
    s1 = C.c_double(1.0)
    for _ in range(2):
        hp.check(N.lib().fgp_lml_gradient(hp.ptr, C.byref(kd), 0.1, 1, C.cast(C.byref(s1), N._dp), N.dptr(g1)))
    out["lml_gradient_single_gpu_ms"] = hp.last_device_ms()
    out["lml_gradient_sharded_ms"] = grad_ms
    out["lml_scale_rel_diff"] = abs(gscale - s1.value) / abs(s1.value)
    out["lml_grads_max_rel_diff"] = float(np.max(np.abs(np.array(ggrads) - g1[:2]) / np.maximum(np.abs(g1[:2]), 1e-300)))
    gu = np.zeros(3)
    hp.check(N.lib().fgp_lml_gradient(hp.ptr, C.byref(kd), 0.1, 0, None, N.dptr(gu)))
    out["lml_unscaled_grads_max_rel_diff"] = float(np.max(np.abs(np.array(ggrads_u) - gu) / np.maximum(np.abs(gu), 1e-300)))
    if n <= 4096:
        from oracle import oracle as O
        ref = O.OracleGaussianProcess(O.ZeroPrior(), O.KernelDesc.make([O.K_SQUARED_EXP], [ls, 1.0]), 0.1, None, X, y)
        out["L_rel_vs_oracle"] = float(np.linalg.norm(np.tril(L) - np.tril(ref.L)) / np.linalg.norm(np.tril(ref.L)))
        rm, rv = ref.predict_mean_variance(Xq)
        out["mean_maxrel_vs_oracle"] = float(np.max(np.abs(mean - rm) / np.maximum(np.abs(rm), 1e-12)))
        out["var_maxrel_vs_oracle"] = float(np.max(np.abs(var - rv) / np.maximum(np.abs(rv), 1e-12)))
digests = [None] * world
dist.all_gather_object(digests, digest)
out["all_ranks_same_factor"] = len(set(digests)) == 1
print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
