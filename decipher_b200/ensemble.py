"""Ensemble drivers above the C ABI: member-sharded single-catchment runs and catchment-partitioned
multi-catchment runs (BASELINE.json configs 3 and 5).  One process per GPU; nothing here touches the
data path — each call ends in `dcp_run_ensemble` on this rank's device."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import capi
from .distributed import lpt_assign, shard_range


def run_members_sharded(cat: capi.Catchment, mcpar: np.ndarray, rank: int = 0, world: int = 1, device: Optional[int] = None,
                        **run_kw) -> Dict:
    """Runs this rank's contiguous block of the (npt, 7, M) parameter table; the table itself must be
    the same on every rank (one sequential RANMAR stream), which makes results independent of `world`."""
    lo, hi = shard_range(mcpar.shape[2], rank, world)
    dev = capi.DeviceCatchment(cat, device=device)
    try:
        out = dev.run(np.asfortranarray(mcpar[:, :, lo:hi]), **run_kw)
    finally:
        dev.close()
    out["member_range"] = (lo, hi)
    return out


def run_members_all_devices(cat: capi.Catchment, mcpar: np.ndarray, n_devices: Optional[int] = None, devices: Optional[Sequence[int]] = None,
                            **run_kw) -> Dict:
    """Single-process multi-GPU run through the C ABI (dcp_comm_init + dcp_run_ensemble_multi): one host thread per
    device inside the library, contiguous member blocks of the ONE table `mcpar`, one NCCL gather of the measures at
    the end (BASELINE config 3 without torchrun).  `devices` may repeat a device (several ranks on one GPU: testing)."""
    if devices is None:
        n = n_devices if n_devices is not None else capi.load_library().dcp_device_count()
        devices = list(range(n))
    comm = capi.Communicator(len(devices), devices=list(devices))
    try:
        return comm.run(cat, np.asfortranarray(mcpar), **run_kw)
    finally:
        comm.close()


def run_catchments(catchments: Sequence[capi.Catchment], mcpars: Sequence[np.ndarray], rank: int = 0, world: int = 1,
                   device: Optional[int] = None, **run_kw) -> Dict[int, Dict]:
    """Multi-catchment run: catchments are dealt to ranks longest-processing-time first by
    nac x nstep x members; each rank runs its catchments one after the other, whole
    ensembles at a time.  Returns {catchment index: outputs} for the catchments this rank owns."""
    costs = [c.nac * c.nstep * m.shape[2] for c, m in zip(catchments, mcpars)]
    owner = lpt_assign(costs, world)
    out: Dict[int, Dict] = {}
    for i, (c, m) in enumerate(zip(catchments, mcpars)):
        if owner[i] != rank:
            continue
        dev = capi.DeviceCatchment(c, device=device)
        try:
            out[i] = dev.run(np.asfortranarray(m), **run_kw)
        finally:
            dev.close()
    return out
