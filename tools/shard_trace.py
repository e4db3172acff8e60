"""Per-panel timeline of the sharded fit (FGP_SHARD_TRACE=1 makes csrc/sharded.cu print it on stderr).  Run under torchrun:
    FGP_SHARD_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/shard_trace.py [n] [d]"""
import math
import os
import sys

import torch  # noqa: F401
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from friedrich_b200 import _native as N  # noqa: E402
from friedrich_b200 import sharded  # noqa: E402
from friedrich_b200.kernels import SquaredExp  # noqa: E402
from friedrich_b200.synthetic import make_dataset  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
kd = SquaredExp(math.sqrt(d / 6.0), 1.0).device_desc()
h = N.Handle(local)
sharded.comm_init(h, rank, world, dist)
if rank == 0:
    X, y = make_dataset(0x5EED0004, n, d)
    sharded.fit_sharded(h, X, y, kd, 0.1)
else:
    sharded.fit_sharded(h, (n, d), None, kd, 0.1)
for _ in range(2):
    dist.barrier()
    print(f"=== refit rank {rank}", file=sys.stderr, flush=True)
    sharded.refit_sharded(h, kd, 0.1)
    print(f"rank {rank} refit device ms {h.last_device_ms():.3f}", flush=True)
dist.barrier()
dist.destroy_process_group()
