# Round 2, GPU call 25: register budgets of the 3..5-D kernels again on the settled code (cN = NTR_MIN_CTAS=N: 8 -> 64
# registers, 7 -> 72, 6 -> 80, 5 -> 96), and the shares of config 4 that each of 8 ranks renders.
set -x
mkdir -p gpurun_out/r02y
O=gpurun_out/r02y
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 120 python tools/quick.py $c $EXTRA --frames 7 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for v in def c7 c6 c5; do
  if [ $v = def ]; then L=A=1; else L=NTR_B200_LIB=$PWD/variants/libntr_$v.so; fi
  EXTRA= ; for c in c4 c4b c2 c4o; do run $v $c $L; done
  EXTRA="--world 8"; run ${v}_w8 c4 $L
done
for r in 0 1 2 3 4 5 6 7; do EXTRA="--world 8 --rank $r"; run r${r}of8 c4 A=1; done
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in $O/q_c4_def.err $O/q_c4_c7.err $O/q_c4_c6.err $O/q_c4_c5.err $O/q_c4_r*of8.err; do echo $f; grep "pass ms" $f | tail -1; done
