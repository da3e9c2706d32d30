"""b200det.peer -- the loss normaliser's cross-GPU exchange over NVLink peer memory.

`RetinaLoss(..., sync_normalizer='p2p')` (one process per GPU, all ranks in one NVLink / NVSwitch
domain) replaces "reduce kernel -> torch.distributed.all_reduce (NCCL) -> finish kernel" by ONE
kernel per rank that reduces the rank's partial sums, stores its 4 doubles into every peer's
exchange buffer, waits for the peers' stores and normalises (csrc/exchange.cu).  torch.distributed
is only used once, to hand the 64-byte CUDA IPC handles of the 4 KB exchange buffers around.
"""
import ctypes

import torch

from . import _lib

__all__ = ['PeerExchange']


class PeerExchange:
    """Exchange buffers of one process group, mapped into this process."""

    def __init__(self, group=None, device=None, timeout_cycles=0):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('PeerExchange needs an initialised torch.distributed process group')
        lib = _lib.load()
        self._lib = lib
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f'at most {_lib.MAX_PEERS} ranks can share a peer exchange')
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self._own = ctypes.c_void_p()
        self._mapped = []
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _lib.check(lib.b200det_peer_buffer_create(ctypes.byref(self._own), handle),
                       'b200det_peer_buffer_create')
            # (hostname, handle): every rank must sit on the same node
            import socket
            mine = (socket.gethostname(), self.device.index, handle.raw)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, mine, group=group)
            if any(g[0] != mine[0] for g in gathered):
                self.close()
                raise RuntimeError('peer exchange needs all ranks on one node (NVLink domain)')
            px = _lib.PeerExchange()
            px.rank, px.world = self.rank, self.world
            px.epoch = 0
            px.timeout_cycles = int(timeout_cycles)
            for r, (_, dev_index, raw) in enumerate(gathered):
                if r == self.rank:
                    px.peer[r] = self._own.value
                    continue
                # (no can_device_access_peer(local index) test: the other process's device index
                # means nothing here when CUDA_VISIBLE_DEVICES differs per rank; mapping the IPC
                # handle enables peer access and fails if the two GPUs are not peers)
                ptr = ctypes.c_void_p()
                rc = lib.b200det_peer_buffer_open(raw, ctypes.byref(ptr))
                if rc != 0:
                    self.close()
                    _lib.check(rc, f'b200det_peer_buffer_open (rank {r}: not a peer of this GPU?)')
                self._mapped.append(ptr)
                px.peer[r] = ptr.value
        self.params = px
        # nobody starts an exchange before every rank has mapped every buffer
        dist.barrier(group=group)

    def next(self):
        """The struct for the next exchange.  epoch = 0: the exchange number lives in device memory
        and the kernel advances it, in lock-step on every rank (every rank runs the same sequence
        of exchanges) -- so a captured CUDA graph that is replayed advances it like eager calls."""
        return ctypes.byref(self.params)

    def close(self):
        lib = self._lib
        for ptr in self._mapped:
            lib.b200det_peer_buffer_close(ptr)
        self._mapped = []
        if self._own:
            lib.b200det_peer_buffer_destroy(self._own)
            self._own = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown
            pass
