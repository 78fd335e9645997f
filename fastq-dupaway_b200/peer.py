"""All-to-all over mapped peer memory (one process per GPU of one box).

Every rank owns a receive buffer (cudaMalloc through the C ABI), exports it with CUDA IPC and maps the buffers of all
its peers.  An exchange is then `world` asynchronous device-to-device copies per rank - each one writes this rank's
part straight into its destination's receive buffer over NVLink / NVSwitch with the copy engines - bracketed by two
barriers.  torch.distributed only carries the handles, the size matrix and the barriers.  (NCCL's all-to-all moves the
same bytes with SM-driven send/recv kernels: 35 ms for 2 x 6.4 GB per GPU in the sequence-mode repartition.)"""
from __future__ import annotations

import ctypes as C


class PeerExchange:
    def __init__(self, pkg, dist, rank, world, device, capacity_bytes):
        import torch
        self.torch, self.dist, self.rank, self.world, self.dev = torch, dist, rank, world, device
        self.lib = pkg.load_library()
        self.cap = int(capacity_bytes)
        self.buf = pkg.DeviceBuffer(self.cap, device)
        handle = C.create_string_buffer(64)
        assert self.lib.fqd_ipc_export(device, C.c_void_p(self.buf.ptr), handle) == 0
        tdev = torch.device("cuda", device)
        # control messages (handles, size matrix) travel on the device with NCCL, on the host with gloo (tests)
        self.cdev = torch.device("cpu") if dist.get_backend() == "gloo" else tdev
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(self.cdev)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        self.peer = []
        for r in range(world):
            if r == rank:
                self.peer.append(self.buf.ptr)
                continue
            p = C.c_void_p()
            hb = bytes(allh[r].cpu().numpy().tobytes())
            assert self.lib.fqd_ipc_open(device, hb, C.byref(p)) == 0, "cudaIpcOpenMemHandle failed"
            self.peer.append(p.value)
        self.tdev = tdev
        self.streams = [torch.cuda.Stream(device=tdev) for _ in range(world)]     # one per destination: the copies overlap
        dist.barrier()

    def start(self, send_ptr, send_sizes):
        """send_sizes[p] bytes (contiguous, in rank order, starting at send_ptr) go to rank p.  Returns after the copies
        have been enqueued; the bytes received from every source are known already (finish() tells when they landed)."""
        torch, dist = self.torch, self.dist
        row = torch.tensor(send_sizes, dtype=torch.int64, device=self.cdev)
        mat = [torch.empty_like(row) for _ in range(self.world)]
        dist.all_gather(mat, row)                          # mat[src][dst]; doubles as the "buffers are free" barrier
        sizes = [[int(x) for x in m.tolist()] for m in mat]
        recv_sizes = [sizes[src][self.rank] for src in range(self.world)]
        if any(sum(sizes[s][d] for s in range(self.world)) > self.cap for d in range(self.world)):
            raise MemoryError("peer exchange: a receive buffer is too small")
        torch.cuda.synchronize(self.tdev)                 # the send buffer is complete
        for k in range(self.world):
            dst = (self.rank + 1 + k) % self.world         # peers first, round-robin (spreads the link load), myself last
            src_off = sum(send_sizes[:dst])
            dst_off = sum(sizes[s][dst] for s in range(self.rank))
            n = send_sizes[dst]
            assert self.lib.fqd_peer_copy_async(self.dev, C.c_void_p(self.peer[dst] + dst_off), C.c_void_p(send_ptr + src_off), n,
                                                C.c_void_p(self.streams[k].cuda_stream)) == 0
        return self.buf.ptr, recv_sizes

    def finish(self):
        for st in self.streams:
            st.synchronize()
        self.dist.barrier()                                # every peer's writes into my buffer have landed

    def exchange(self, send_ptr, send_sizes):
        """start + finish.  Returns (device pointer of this rank's receive buffer, [bytes received from every source])."""
        r = self.start(send_ptr, send_sizes)
        self.finish()
        return r

    def close(self):
        for r, p in enumerate(self.peer):
            if r != self.rank:
                self.lib.fqd_ipc_close(self.dev, C.c_void_p(p))
        self.buf.free()
