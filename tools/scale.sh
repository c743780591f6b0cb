# 1 -> 8 GPU scaling of bench.py (same launcher as the driver)
for cfg in c2 c4; do
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --config $cfg --no-cpu-baseline > gpurun_out/scale_${cfg}_$n.json 2>gpurun_out/scale_${cfg}_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 --config $cfg 2>gpurun_out/scale_${cfg}_$n.err | grep '^{' > gpurun_out/scale_${cfg}_$n.json; fi
  python -c "
import json; d=json.load(open('gpurun_out/scale_${cfg}_$n.json')); print('$cfg', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'Mrays/s', round(d['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],4))"
done; done
