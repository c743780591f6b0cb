# Round 2, GPU call 23: the primary pass of opaque scenes as a tracing launch + a shading launch (NTR_SPLIT_SHADE=1).
set -x
mkdir -p gpurun_out/r02w
O=gpurun_out/r02w
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_trace" 2>&1 | tail -5
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 120 python tools/quick.py $c $EXTRA --frames 9 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for v in def split; do
  if [ $v = def ]; then E=A=1; else E=NTR_SPLIT_SHADE=1; fi
  EXTRA= ; for c in c2 c4o c5s c1; do run $v $c $E; done
  EXTRA="--world 8"; for c in c2 c4o; do run ${v}_w8 $c $E; done
  EXTRA="--world 2"; run ${v}_w2 c2 $E
done
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
NTR_SPLIT_SHADE=1 timeout 300 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > $O/bench_c2_split.json 2> $O/bench_c2_split.err; tail -c 300 $O/bench_c2_split.json
timeout 300 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > $O/bench_c2_def.json 2> $O/bench_c2_def.err; tail -c 300 $O/bench_c2_def.json
