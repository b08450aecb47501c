"""Device-mesh addressing: multi-level split address -> global device, device groups per split level.

Python mirror of the C++ CommDevManager (csrc/runtime/dev_mesh.h; reference: xla/pjrt/dev_id_util.{h,cc},
SURVEY D1).  Levels flagged `share_dev` (micro-batch level) are time-multiplexed on the same device and
contribute nothing to the device id.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

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
        if not self.placement_layout:
            self.placement_layout = list(range(len(self.split_nums)))
        dev_levels = [l for l in self.placement_layout if not self.share_dev[l]]
        self.base: Dict[int, int] = {}
        b = 1
        for l in reversed(dev_levels):  # innermost level varies fastest
            self.base[l] = b
            b *= self.split_nums[l]
        self.total_devices = b
        assert self.total_devices == self.world or self.world == 1, (self.split_nums, self.world)

    def coords(self, device: Optional[int] = None) -> Dict[int, int]:
        d = self.rank if device is None else device
        return {l: (d // self.base[l]) % self.split_nums[l] for l in self.base}

    def device_of(self, ids: Dict[int, int]) -> int:
        return sum(ids.get(l, 0) * self.base[l] for l in self.base)

    def group_ranks(self, level: int, device: Optional[int] = None) -> List[int]:
        d = self.rank if device is None else device
        c = self.coords(d)
        origin = d - c[level] * self.base[level]
        return [origin + i * self.base[level] for i in range(self.split_nums[level])]

    def index_in_group(self, level: int) -> int:
        return self.coords()[level]

    def all_groups(self, level: int) -> List[List[int]]:
        seen, out = set(), []
        for d in range(self.total_devices):
            g = tuple(self.group_ranks(level, d))
            if g not in seen:
                seen.add(g)
                out.append(list(g))
        return out

    def build_process_groups(self) -> None:
        """Every rank creates every group of every device level in the same order (required by torch.distributed)."""
        if self.world == 1 or not dist.is_initialized():
            return
        for l in self.base:
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
