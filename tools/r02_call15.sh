# Round 2, GPU call 15: merged bounce launch, fourth version (control words on separate cache lines, idle warps poll slowly
# and leave when the rays in flight cannot feed them).
set -x
mkdir -p gpurun_out/r02o
O=gpurun_out/r02o
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 60 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for m in 1 2 0; do
  EXTRA= ; for c in c4 c4b c3; do run m$m $c NTR_MERGE_FROM=$m; done
  EXTRA="--world 8"; run m${m}_w8 c4 NTR_MERGE_FROM=$m; run m${m}_w8 c4b NTR_MERGE_FROM=$m
  EXTRA="--world 4"; run m${m}_w4 c4 NTR_MERGE_FROM=$m
done
for v in nap6 nap100 idle8; do
  EXTRA= ; run m1_$v c4 NTR_MERGE_FROM=1 NTR_B200_LIB=$PWD/variants/libntr_$v.so
  EXTRA="--world 8"; run m1_w8_$v c4 NTR_MERGE_FROM=1 NTR_B200_LIB=$PWD/variants/libntr_$v.so
done
EXTRA= ; run def c2 A=1
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in q_c4_m1 q_c4_m1_w8 q_c4_m0_w8 q_c4_m2_w8; do tail -1 $O/$f.err; done
NTR_MERGE_FROM=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_stream.py tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -4
