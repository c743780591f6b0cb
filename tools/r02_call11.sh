# Round 2, GPU call 11: tuning of the cooperative-leaf thresholds on the settled code (variants), config 4 and its 1/8 share.
set -x
mkdir -p gpurun_out/r02k
O=gpurun_out/r02k
run() { local name=$1 lib=$2 c=$3; shift 3; env NTR_PASS_TIMING=1 NTR_B200_LIB=$PWD/variants/libntr_$lib.so "$@" timeout 300 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for v in base md8 md24 lm24 lm96 cc1 cc3; do
  [ -f variants/libntr_$v.so ] || continue
  EXTRA= ; run $v $v c4 A=1; EXTRA="--world 8"; run ${v}_w8 $v c4 A=1; EXTRA=
done
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
