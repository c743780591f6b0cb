# Round 2, GPU call 16 (8 GPUs): N-device frames against the 1-device frame, and the bench line at N = 8, 4, 1 on one box.
set -x
mkdir -p gpurun_out/r02p
O=gpurun_out/r02p
nvidia-smi -L | head -8 > $O/gpus.txt
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -5 | tee $O/test_multi_gpu.txt
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err
  tail -c 1500 $O/bench_n$n.json
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --stream-frames 0 > $O/bench_n1.json 2> $O/bench_n1.err
tail -c 1500 $O/bench_n1.json
