"""Multi-GPU frame partitioning (no counterpart in the reference): one process per GPU,
scene replicated, the frame split by interleaved 32-pixel tile rows (tile row ty belongs to rank ty % world, which
balances the centre-heavy cost of polytope scenes).  Two ways to put the frame together:

  PeerFrameRenderer     the frame buffer lives on rank 0's GPU and every rank's packing epilogue stores its rows straight
                        into it over NVLink (CUDA IPC mapping, ntr_frame_*); the only collective is a 4-byte NCCL
                        all-reduce that orders rank 0's stream behind everybody's kernels.  What bench.py times.
  DistributedRenderer   every rank renders a compact strip (ntr_render_device(..., compact=1)), the strips are gathered
                        with one NCCL all-gather and un-interleaved with one index_select (round 1; NTR_BENCH_GATHER=nccl).

A single process that owns several GPUs uses ntr_group_* instead (backend.DeviceGroup): same partition, same peer stores.
torch / torch.distributed are plumbing only (device buffers, NCCL)."""
import torch
import torch.distributed as dist

TILE = 32       # RENDER_CHUNK_SIZE (reference src/render.cpp:43)


def tile_rows(height, rank, world):
    return [ty for ty in range((height + TILE - 1) // TILE) if ty % world == rank]


def max_rows_per_rank(height, world):
    return ((height + TILE - 1) // TILE + world - 1) // world


def strip_bytes(height, pitch, world):
    """Size of every rank's strip buffer (padded to the largest share so the all-gather is regular)."""
    return max_rows_per_rank(height, world) * TILE * pitch


def row_map(height, world, device='cpu'):
    """For every frame row y: its row index in the gathered buffer viewed as [world * max_rows * 32, pitch]."""
    mr = max_rows_per_rank(height, world)
    y = torch.arange(height, dtype=torch.int64)
    ty = y // TILE
    return ((ty % world) * (mr * TILE) + (ty // world) * TILE + (y % TILE)).to(device)


def extract_strip(frame, height, pitch, rank, world):
    """The compact strip rank `rank` would render, cut out of a full frame (testing / CPU emulation)."""
    f = frame.reshape(height, pitch)
    out = torch.zeros(max_rows_per_rank(height, world) * TILE, pitch, dtype=frame.dtype, device=frame.device)
    for k, ty in enumerate(tile_rows(height, rank, world)):
        n = min(TILE, height - ty * TILE)
        out[k * TILE:k * TILE + n] = f[ty * TILE:ty * TILE + n]
    return out.reshape(-1)


def compose(gathered, height, pitch, world, rmap=None):
    """gathered: uint8 [world * strip_bytes] -> frame [height * pitch] (one gather kernel)."""
    if rmap is None:
        rmap = row_map(height, world, gathered.device)
    return gathered.reshape(-1, pitch).index_select(0, rmap).reshape(-1)


class DistributedRenderer:
    """Renders frames of one replicated DeviceScene across the ranks of a process group."""
    def __init__(self, scene, fmt, group=None):
        self.ds, self.fmt, self.group = scene, fmt, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        # diagnostic: render only the share rank 0 of N would get, on one GPU (profiling the multi-GPU kernel shape
        # without running ncu on a multi-rank job)
        import os
        self.fake_world = int(os.environ.get('NTR_FAKE_WORLD', '0')) if self.world == 1 else 0
        dev = torch.device('cuda', torch.cuda.current_device())
        # a dedicated (non-null) stream: the C ABI treats a NULL stream as "the scene's own stream"
        self.stream = torch.cuda.Stream(device=dev)
        nb = strip_bytes(fmt.height, fmt.pitch, self.world)
        self.strip = torch.zeros(nb, dtype=torch.uint8, device=dev)
        self.gathered = torch.zeros(nb * self.world, dtype=torch.uint8, device=dev) if self.world > 1 else self.strip
        self.rmap = row_map(fmt.height, self.world, dev)
        self.host = torch.zeros(fmt.height * fmt.pitch, dtype=torch.uint8).pin_memory() if self.rank == 0 else None

    def render_strip(self):
        """Enqueue this rank's tile rows on self.stream (asynchronous for scenes without reflective materials)."""
        self.ds.render_device(self.fmt, self.strip.data_ptr(), self.strip.numel(), self.stream.cuda_stream,
                              self.rank, self.fake_world or self.world, True)

    def gather(self):
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(self.gathered, self.strip, group=self.group)

    def frame_on_device(self):
        with torch.cuda.stream(self.stream):
            return compose(self.gathered, self.fmt.height, self.fmt.pitch, self.world, self.rmap)

    def render_to_host(self):
        """Whole frame into pinned host memory on rank 0 (returns the pinned tensor there, None elsewhere)."""
        self.render_strip()
        self.gather()
        if self.rank == 0:
            with torch.cuda.stream(self.stream):
                self.host.copy_(self.frame_on_device(), non_blocking=True)
        self.stream.synchronize()
        return self.host


class PeerFrameRenderer:
    """One frame buffer on rank 0's GPU, written by every rank's kernels over NVLink (no gather, no compose)."""
    def __init__(self, scene, fmt, group=None):
        from .backend import SharedFrame
        self.ds, self.fmt, self.group = scene, fmt, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.cuda.current_device()
        dev = torch.device('cuda', self.device)
        self.stream = torch.cuda.Stream(device=dev)
        self.nbytes = fmt.pitch * fmt.height
        box = [None]
        if self.rank == 0:
            self.frame = SharedFrame(self.device, self.nbytes)
            box[0] = self.frame.export()
        if self.world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        if self.rank != 0:
            self.frame = SharedFrame(self.device, handle=box[0])
        self.token = torch.zeros(1, dtype=torch.float32, device=dev)
        self.host = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory() if self.rank == 0 else None
        if self.world > 1:
            dist.barrier(group=group)

    def render(self):
        """Enqueue this rank's tile rows on self.stream; the pixels land in rank 0's frame at their frame position."""
        self.ds.render_device(self.fmt, self.frame.ptr, self.nbytes, self.stream.cuda_stream, self.rank, self.world, False)

    def fence(self):
        """Order self.stream behind the kernels of every rank (a 4-byte all-reduce: the frame's only collective)."""
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.token, group=self.group)

    def render_to_host(self):
        """Whole frame into pinned host memory on rank 0 (returns the pinned tensor there, None elsewhere)."""
        self.render()
        self.fence()
        if self.rank == 0:
            self.frame.download(self.fmt, self.host.numpy(), self.stream.cuda_stream)
        self.stream.synchronize()
        return self.host

    def close(self):
        self.stream.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)      # nobody unmaps the frame while a peer may still store into it
        if self.rank != 0:
            self.frame.close()
        if self.world > 1:
            dist.barrier(group=self.group)      # ... and the owner frees it only after every mapping is gone
        if self.rank == 0:
            self.frame.close()
