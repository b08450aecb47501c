"""Device-mesh addressing: multi-level split address -> global device, device groups per split level.

ONE implementation: the C++ CommDevManager (csrc/runtime/dev_mesh.{h,cc}; reference: xla/pjrt/dev_id_util.{h,cc}).
`DeviceMesh` is the runtime's handle on it -- it owns no addressing arithmetic of its own, it binds this rank to a manager built
from (split_nums, share_dev, placement_layout) and adds the torch.distributed process groups of the manager's device groups.
Levels flagged `share_dev` (the micro-batch level) are time-multiplexed on the same device and contribute nothing to the
device id.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch.distributed as dist


@dataclass
class DeviceMesh:
    split_nums: List[int]
    share_dev: List[bool]
    placement_layout: List[int] = field(default_factory=list)  # order of levels from outermost to innermost
    rank: int = 0
    world: int = 1
    _groups: Dict[int, object] = field(default_factory=dict)
    _group_ranks: Dict[int, List[int]] = field(default_factory=dict)

    def __post_init__(self):
        from .. import _C
        if not self.placement_layout:
            self.placement_layout = list(range(len(self.split_nums)))
        self.mgr = _C.CommDevManager()
        self.mgr.build([int(n) for n in self.split_nums], [bool(s) for s in self.share_dev], [int(l) for l in self.placement_layout])
        self.total_devices = self.mgr.total_devices()
        self.device_levels = [l for l in range(len(self.split_nums)) if not self.share_dev[l]]
        assert self.total_devices == self.world or self.world == 1, (self.split_nums, self.world)

    def coords(self, device: Optional[int] = None) -> Dict[int, int]:
        """{level: index of `device` (default: this rank) along that level} for the device levels."""
        c = self.mgr.coords(self.rank if device is None else device)
        return {l: c[l] for l in self.device_levels}

    def device_of(self, ids: Dict[int, int]) -> int:
        return self.mgr.global_device([int(ids.get(l, 0)) for l in range(len(self.split_nums))])

    def stride(self, level: int) -> int:
        """Distance in global device ids between neighbours along `level` (pipeline: rank of the next stage = rank + stride)."""
        g = self.mgr.group_of(0, level).devices
        return g[1] - g[0] if len(g) > 1 else 0

    def group_ranks(self, level: int, device: Optional[int] = None) -> List[int]:
        return list(self.mgr.group_of(self.rank if device is None else device, level).devices)

    def index_in_group(self, level: int) -> int:
        return self.mgr.rank_in_group(self.rank, level)

    def all_groups(self, level: int) -> List[List[int]]:
        return [list(g.devices) for g in self.mgr.all_groups(level)]

    def build_process_groups(self) -> None:
        """Every rank creates every group of every device level in the same order (required by torch.distributed)."""
        if self.world == 1 or not dist.is_initialized():
            return
        for l in self.device_levels:
            for ranks in self.all_groups(l):
                if len(ranks) == self.world:
                    pg = dist.group.WORLD
                else:
                    pg = dist.new_group(ranks)
                if self.rank in ranks:
                    self._groups[l] = pg
                    self._group_ranks[l] = ranks

    def group(self, level: int):
        return self._groups.get(level)
