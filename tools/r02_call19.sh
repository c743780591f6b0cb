# Round 2, GPU call 19: software prefetch of the batch records NTR_PREFETCH_AHEAD items ahead in the scans of big leaves.
set -x
mkdir -p gpurun_out/r02s
O=gpurun_out/r02s
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 60 python tools/quick.py $c $EXTRA --frames 5 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for v in def pf2 pf4 pf8; do
  if [ $v = def ]; then L=; else L=NTR_B200_LIB=$PWD/variants/libntr_$v.so; fi
  EXTRA= ; for c in c4 c4b c4o c2 c3 c5s; do run $v $c A=1 $L; done
  EXTRA="--world 8"; for c in c4 c4b c4o c2; do run ${v}_w8 $c A=1 $L; done
done
EXTRA="--world 8"; run pf4s_w8 c4 NTR_B200_LIB=$PWD/variants/libntr_pf4s.so
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in $O/q_c4_*.err; do echo $f; grep "pass ms" $f | tail -1; grep "fetch stats" $f | tail -5 | cut -c1-300; done
