# Round 2, GPU call 27: the 96-register build picked per pass by measurement (auto) against never (NTR_WIDE=0) / always (1).
set -x
mkdir -p gpurun_out/r02za
O=gpurun_out/r02za
NTR_WIDE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "polytope or mixed or random or interleaved" 2>&1 | tail -3
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 120 python tools/quick.py $c $EXTRA --frames 9 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for m in auto 0 1; do
  if [ $m = auto ]; then E=A=1; else E=NTR_WIDE=$m; fi
  EXTRA= ; run w$m c4 $E; run w$m c4b $E
  EXTRA="--world 8"; run w${m}_w8 c4 $E; run w${m}_w8 c4b $E
  EXTRA="--world 4"; run w${m}_w4 c4 $E
  EXTRA="--world 2"; run w${m}_w2 c4 $E
done
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in $O/q_c4*_wauto*.err; do echo $f; grep "wide builds" $f | tail -1; grep "pass ms" $f | tail -1; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c4_w3.json 2> $O/bench_c4_w3.err; tail -c 400 $O/bench_c4_w3.json
