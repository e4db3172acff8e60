"""Multi-GPU fit: one process per GPU, block-cyclic panels, NCCL panel broadcasts inside libfgp_sm100 (csrc/sharded.cu).

The reference crate is single-threaded, so there is no reference interface to mirror here; this module is the thin host
side of `fgp_comm_*` / `fgp_fit_sharded` (include/fgp.h).  `torch.distributed` (any backend, gloo is enough) is used for
one thing only: shipping rank 0's 128-byte NCCL unique id to the other processes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def shard_plan(n, world, rank):
    """Host-only view of the partition (`fgp_shard_plan`): dict(panel_cols, n_panels, n_owned, flop_share, owned)."""
    pc, npn, own = N._i64(), N._i64(), N._i64()
    share = C.c_double()
    rc = N.lib().fgp_shard_plan(int(n), int(world), int(rank), C.byref(pc), C.byref(npn), C.byref(own), C.byref(share))
    if rc != N.FGP_OK:
        raise N.FgpError(rc, "fgp_shard_plan: bad arguments")
    return {"panel_cols": pc.value, "n_panels": npn.value, "n_owned": own.value, "flop_share": share.value,
            "owned": [p for p in range(npn.value) if p % world == rank]}


def exchange_id(dist, rank, make_id):
    """Rank 0 calls `make_id()` (-> bytes); every rank returns those bytes.  `dist` is an initialised torch.distributed."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def nccl_unique_id():
    buf = (C.c_ubyte * N.FGP_COMM_ID_BYTES)()
    rc = N.lib().fgp_comm_unique_id(buf, N.FGP_COMM_ID_BYTES)
    if rc != N.FGP_OK:
        raise N.FgpError(rc, "ncclGetUniqueId failed")
    return bytes(buf)


def comm_init(handle, rank, world, dist=None):
    """Create the NCCL communicator of `handle` (one handle = one GPU = one rank)."""
    if world > 1:
        if dist is None:
            raise ValueError("world > 1 needs an initialised torch.distributed to exchange the NCCL id")
        uid = exchange_id(dist, rank, nccl_unique_id)
    else:
        uid = bytes(N.FGP_COMM_ID_BYTES)
    buf = (C.c_ubyte * N.FGP_COMM_ID_BYTES).from_buffer_copy(uid)
    handle.check(N.lib().fgp_comm_init_rank(handle.ptr, buf, N.FGP_COMM_ID_BYTES, int(world), int(rank)))


def fit_sharded(handle, X, y_resid, kernel_desc, noise, cholesky_epsilon=None):
    """Collective `fgp_fit_sharded`; X / y_resid may be None on ranks other than 0, but n and d must agree: pass
    X=(n, d) as a shape tuple in that case."""
    if isinstance(X, tuple):
        n, d = X
        xp, yp, ld = None, None, n
    else:
        X = N.fcol(X)
        y = np.ascontiguousarray(y_resid, dtype=np.float64)
        n, d = X.shape
        xp, yp, ld = N.dptr(X), N.dptr(y), n
    has_eps, eps = (0, 0.0) if cholesky_epsilon is None else (1, float(cholesky_epsilon))
    handle.check(N.lib().fgp_fit_sharded(handle.ptr, xp, ld, n, d, yp, C.byref(kernel_desc), float(noise), has_eps, eps))


def refit_sharded(handle, kernel_desc, noise, cholesky_epsilon=None):
    has_eps, eps = (0, 0.0) if cholesky_epsilon is None else (1, float(cholesky_epsilon))
    handle.check(N.lib().fgp_refit_sharded(handle.ptr, C.byref(kernel_desc), float(noise), has_eps, eps))


def lml_gradient_sharded(handle, kernel_desc, noise, nparams, scaled=True):
    """Collective `fgp_lml_gradient_sharded` (every rank calls it after the sharded fit): (scale, grads) as the reference's
    scaled_gradient_marginal_likelihood (optimizer.rs:159-203), or (1.0, grads + [noise gradient]) for scaled=False."""
    grads = np.zeros(nparams + 1)
    scale = C.c_double(1.0)
    handle.check(N.lib().fgp_lml_gradient_sharded(handle.ptr, C.byref(kernel_desc), float(noise), 1 if scaled else 0,
                                                  C.cast(C.byref(scale), N._dp), N.dptr(grads)))
    return scale.value, (list(grads[:nparams]) if scaled else list(grads))
