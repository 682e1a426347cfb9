"""Row-band partition of a frame over N devices and the gather of the bands on rank 0.

This is the reference's own distributed partition (its TCP "network device"): server `id` of `count` renders the raster
rows y with ((y >> 2) - id) % count == 0 into the compacted buffer row 4*((y >> 2) / count) + (y & 3)
(devices/device_singleray/api/swapchain.h:57-70) and the client re-interleaves the returned rows
(devices/device_network/network_device.cpp:235-310). device_cuda renders exactly those rows when created with
cfg "serverID=id,serverCount=count"; here the bands travel device to device over torch.distributed (NCCL on GPUs,
gloo in the CPU tests) instead of TCP. No reduction is involved: the bands are disjoint.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def active_rows(height: int, rank: int, world: int) -> List[int]:
    """Raster rows rendered by `rank`, in buffer order (api/swapchain.h:57-60)."""
    return [y for y in range(height) if ((y >> 2) - rank) % world == 0]


def buffer_row(y: int, world: int) -> int:
    """raster2buffer (api/swapchain.h:68-70)."""
    return 4 * ((y >> 2) // world) + (y & 3)


class BandGather:
    """Gathers the compacted row bands of every rank into the full frame on rank 0."""

    def __init__(self, height: int, stride_bytes: int, rank: int, world: int, device):
        self.height, self.stride, self.rank, self.world, self.device = height, stride_bytes, rank, world, device
        self.rows = [active_rows(height, r, world) for r in range(world)]
        self.max_rows = max(len(r) for r in self.rows)
        self.full: Optional[torch.Tensor] = None
        # The bands are exchanged with ONE all-gather (a single NCCL kernel over NVLink / NVSwitch; 3 MB per 1024^2 RGB8 face) instead of a
        # gather's grouped send/recv pairs, whose latency was 0.26-0.67 ms per face at 8 ranks: every rank receives the block, rank 0
        # re-interleaves it. One extra (dummy) row takes the padding rows of ranks that own fewer than max_rows rows, so that the
        # re-interleave is a single index_copy over the gathered block.
        self.block = torch.empty(world * self.max_rows * stride_bytes, dtype=torch.uint8, device=device)
        if rank == 0:
            self.full = torch.zeros((height + 1) * stride_bytes, dtype=torch.uint8, device=device)
            idx = []
            for r in self.rows:
                idx += r + [height] * (self.max_rows - len(r))
            self.index = torch.tensor(idx, dtype=torch.long, device=device)

    def gather(self, local: torch.Tensor) -> Optional[torch.Tensor]:
        """`local`: this rank's framebuffer bytes (compacted rows first). Returns the full frame on rank 0."""
        send = local[: self.max_rows * self.stride]
        if self.world == 1:
            self.block.copy_(send)
        else:
            dist.all_gather_into_tensor(self.block, send.contiguous())
            if self.rank != 0:
                return None
        self.full.view(self.height + 1, self.stride).index_copy_(0, self.index, self.block.view(-1, self.stride))
        return self.full[: self.height * self.stride]


class CubeBandGather:
    """The same exchange for all faces of a stereo cube map at once: ONE all-gather per cube map, then one index-copy on rank 0
    re-interleaves every face. `faces` frame buffers of `height` rows each; returns faces * height * stride bytes, face-major."""

    def __init__(self, faces: int, height: int, stride_bytes: int, rank: int, world: int, device):
        self.faces, self.height, self.stride, self.rank, self.world, self.device = faces, height, stride_bytes, rank, world, device
        rows = [active_rows(height, r, world) for r in range(world)]
        self.max_rows = max(len(r) for r in rows)
        self.send = torch.empty(faces * self.max_rows * stride_bytes, dtype=torch.uint8, device=device)
        self.block = torch.empty(world * faces * self.max_rows * stride_bytes, dtype=torch.uint8, device=device)
        self.full: Optional[torch.Tensor] = None
        if rank == 0:
            self.full = torch.zeros((faces * height + 1) * stride_bytes, dtype=torch.uint8, device=device)
            idx = []
            for r in rows:
                for f in range(faces):
                    idx += [f * height + y for y in r] + [faces * height] * (self.max_rows - len(r))
            self.index = torch.tensor(idx, dtype=torch.long, device=device)

    def gather(self, locals_: List[torch.Tensor]) -> Optional[torch.Tensor]:
        """`locals_`: this rank's `faces` framebuffers as byte tensors (compacted rows first). Returns the cube map's frames on rank 0."""
        n = self.max_rows * self.stride
        sv = self.send.view(self.faces, n)
        for f, l in enumerate(locals_):
            sv[f].copy_(l[:n])
        if self.world == 1:
            self.block.copy_(self.send)
        else:
            dist.all_gather_into_tensor(self.block, self.send)
            if self.rank != 0:
                return None
        self.full.view(self.faces * self.height + 1, self.stride).index_copy_(0, self.index, self.block.view(-1, self.stride))
        return self.full[: self.faces * self.height * self.stride]
