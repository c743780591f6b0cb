"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo process groups gather per-rank tile-row
strips and un-interleave them into the frame (ntracer_b200.dist)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ntracer_b200 import dist as ntd


def test_partition_math():
    for h in (1, 31, 32, 33, 480, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = [ntd.tile_rows(h, r, world) for r in range(world)]
            allr = sorted(sum(rows, []))
            assert allr == list(range((h + 31) // 32))
            assert max(len(r) for r in rows) == ntd.max_rows_per_rank(h, world)
            assert max(len(r) for r in rows) - min(len(r) for r in rows) <= 1
            m = ntd.row_map(h, world)
            assert len(set(m.tolist())) == h


def _worker(rank, world, port, h, pitch, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1234)
    frame = torch.randint(0, 256, (h * pitch,), dtype=torch.uint8, generator=g)       # same on every rank
    strip = ntd.extract_strip(frame, h, pitch, rank, world)                           # what this rank "rendered"
    assert strip.numel() == ntd.strip_bytes(h, pitch, world)
    gathered = [torch.zeros_like(strip) for _ in range(world)]
    dist.all_gather(gathered, strip)
    out = ntd.compose(torch.cat(gathered), h, pitch, world)
    ok = bool(torch.equal(out, frame))
    open(os.path.join(out_dir, 'r%d' % rank), 'w').write('ok' if ok else 'bad')
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,h,pitch', [(2, 150, 601), (3, 1080, 96), (2, 33, 7)])
def test_gloo_gather_and_compose(tmp_path, world, h, pitch):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, h, pitch, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), 'r%d' % r)).read() == 'ok'
