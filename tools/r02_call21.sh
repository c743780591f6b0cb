# Round 2, GPU call 21: full GPU suite (no -x), the default bench line and the reference arm on the settled code,
# A/B of the single-copy edge part (NTR_EDGE_LOOP=1).
set -x
mkdir -p gpurun_out/r02u
O=gpurun_out/r02u
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/tests.txt; cat $O/tests.txt
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 120 python tools/quick.py $c $EXTRA --frames 7 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for v in def el; do
  if [ $v = def ]; then L=; else L=NTR_B200_LIB=$PWD/variants/libntr_$v.so; fi
  EXTRA= ; for c in c2 c3 c4 c4b c4o c5s; do run $v $c A=1 $L; done
  EXTRA="--world 8"; for c in c4 c2; do run ${v}_w8 $c A=1 $L; done
done
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.json
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 400 $O/bench_reference.json
