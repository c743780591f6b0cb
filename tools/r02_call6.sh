# Round 2, GPU call 6 (2 GPUs): the multi-GPU paths on real peers -- ntr_group_* (one process), shared frame over CUDA IPC
# (one process per GPU, torchrun), against the NCCL all-gather of round 1.
set -x
mkdir -p gpurun_out/r02f
O=gpurun_out/r02f
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_cpp_host.py -m gpu -q 2>&1 | tail -8 > $O/tests.txt
cat $O/tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2_peer.json 2> $O/bench_n2_peer.err
tail -c 2500 $O/bench_n2_peer.json; tail -5 $O/bench_n2_peer.err
NTR_BENCH_GATHER=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-secondary > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
tail -c 1200 $O/bench_n2_nccl.json
timeout 600 python - > $O/group.txt 2>&1 <<'PY'
import sys, time, statistics
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from ntracer_b200 import _capi
from ntracer_b200.backend import DeviceGroup, DeviceScene
for name in ('c4', 'c4b', 'c2'):
    fixture, w, h, desc = bench.CONFIGS[name]
    sc, g = bench.load_fixture(fixture)
    fmt = _capi.make_image_format(w, h, _capi.RGB8)
    host = torch.zeros(fmt.pitch * h, dtype=torch.uint8).pin_memory().numpy()
    for n in (1, 2):
        with DeviceGroup(sc, n) as grp:
            ms, wall = [], []
            for i in range(6):
                t = time.perf_counter(); grp.render(fmt, host); wall.append((time.perf_counter() - t) * 1e3); ms.append(grp.last_kernel_ms())
            print(name, 'group of', n, 'device ms', round(statistics.median(ms[2:]), 3), 'e2e ms', round(statistics.median(wall[2:]), 3), flush=True)
PY
cat $O/group.txt
