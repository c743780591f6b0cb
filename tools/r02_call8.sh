# Round 2, GPU call 8: full GPU suite on the settled code, the 10-D kernels with the shared-memory axis tables, config 5.
set -x
mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/tests.txt
cat $O/tests.txt
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 600 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
EXTRA= ; for c in c1 c2 c3 c4 c4b c4o c5s; do run def $c A=1; done
EXTRA="--frames 2"; run def c5 A=1; EXTRA=
[ -f variants/libntr_noaxis.so ] && { EXTRA= ; NTR_B200_LIB=$PWD/variants/libntr_noaxis.so; export NTR_B200_LIB; run noaxis c5s A=1; EXTRA="--frames 2"; run noaxis c5 A=1; unset NTR_B200_LIB; EXTRA= ; }
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
timeout 600 python bench.py --config c5s --steps 5 --warmup 3 > $O/bench_c5s.json 2> $O/bench_c5s.err; tail -c 1500 $O/bench_c5s.json
timeout 900 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 1500 $O/bench_c5.json
