"""SM partitioning between the two halves of the pipelined step (CUDA green contexts).

The transformer chain (forward, backward, optimiser of batch i) is a sequence of ~60 dependent kernels of
32-128 CTAs: latency bound, never more than half of the GPU busy.  The frozen PointNet++ encoder of batch
i+1 is throughput work whose persistent kernels own every SM they are given.  On one set of SMs the two
do not overlap — whichever kernel is resident holds the SM (profiles/r1_ab_experiments.md: phases sum to
665 us, the "overlapped" step took 644 us).  Two green contexts split the 148 SMs into two disjoint sets;
streams created in them run concurrently WITHOUT competing for SMs, also inside one captured CUDA graph
(tools/green_probe.py: two GEMM streams on 64 + 84 SMs take exactly the time of the slower one).

Driver API through cuda-python (`cuda.bindings.driver`): cuDeviceGetDevResource, cuDevSmResourceSplitByCount,
cuDevResourceGenerateDesc, cuGreenCtxCreate, cuGreenCtxStreamCreate; the streams are handed to torch as
`torch.cuda.ExternalStream`.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


class SmPartition:
    """`n_first` SMs (rounded by the driver to its granularity) for group 0, the rest for group 1."""

    def __init__(self, device: torch.device, n_first: int):
        from cuda.bindings import driver as drv
        self._drv = drv
        self.device = device
        torch.zeros(1, device=device)                      # primary context exists
        dev = self._chk(drv.cuDeviceGet(device.index or 0))
        res = self._chk(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
        self.total = int(res.sm.smCount)
        err, groups, _, remaining = drv.cuDevSmResourceSplitByCount(1, res, 0, int(n_first))
        if err != drv.CUresult.CUDA_SUCCESS:
            raise RuntimeError(f"cuDevSmResourceSplitByCount failed: {err}")
        self.counts: Tuple[int, int] = (int(groups[0].sm.smCount), int(remaining.sm.smCount))
        self._ctx = []
        for r in (groups[0], remaining):
            desc = self._chk(drv.cuDevResourceGenerateDesc([r], 1))
            self._ctx.append(self._chk(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM)))
        self._streams: List[torch.cuda.Stream] = []

    @staticmethod
    def _chk(r):
        if int(r[0]) != 0:
            raise RuntimeError(f"CUDA driver error {r[0]}")
        return r[1] if len(r) == 2 else r[1:]

    def stream(self, group: int, priority: int = 0) -> torch.cuda.Stream:
        """A new non-blocking stream whose kernels run only on the SMs of `group` (0 or 1)."""
        drv = self._drv
        s = self._chk(drv.cuGreenCtxStreamCreate(self._ctx[group], drv.CUstream_flags.CU_STREAM_NON_BLOCKING, priority))
        ts = torch.cuda.ExternalStream(int(s), device=self.device)
        self._streams.append(ts)
        return ts


def make(device: torch.device, n_first: int) -> Optional[SmPartition]:
    """None (with the reason printed once) when green contexts are not available in this environment."""
    try:
        return SmPartition(device, n_first)
    except Exception as e:  # noqa: BLE001
        import warnings
        warnings.warn(f"msr3d_b200: SM partitioning unavailable ({type(e).__name__}: {e}); the encoder and the "
                      f"transformer chain share all SMs")
        return None
