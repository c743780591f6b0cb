# Round 2, last multi-GPU call (4 GPUs of one box): the multi-GPU tests and the bench command of the driver at N = 4 and 2
# on the final code (frames sharded over 4 or more GPUs run the wide build of the general kernels).
set -x
mkdir -p gpurun_out/r02n4
O=gpurun_out/r02n4
nvidia-smi -L | tee $O/gpus.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -4 | tee $O/tests_multi_gpu.txt
for n in 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 2>$O/bench_c4_n$n.err | grep '^{' > $O/bench_c4_n$n.json
  python -c "
import json; d=json.load(open('$O/bench_c4_n$n.json')); print(d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'Mrays/s', round(d['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['clocks'])"
done
