# Round 2, GPU call 12: one launch for every bounce depth (merged queue, NTR_MERGE_FROM) against a sorted launch per depth.
set -x
mkdir -p gpurun_out/r02l
O=gpurun_out/r02l
NTR_MERGE_FROM=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_stream.py tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -6 > $O/tests_merge1.txt
cat $O/tests_merge1.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_stream.py tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -6 > $O/tests_default.txt
cat $O/tests_default.txt
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 120 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for m in 0 1 2 3; do
  EXTRA= ; for c in c4 c4b c3; do run m$m $c NTR_MERGE_FROM=$m; done
  EXTRA="--world 8"; run m${m}_w8 c4 NTR_MERGE_FROM=$m; run m${m}_w8 c4b NTR_MERGE_FROM=$m
  EXTRA="--world 2"; run m${m}_w2 c4 NTR_MERGE_FROM=$m; EXTRA="--world 4"; run m${m}_w4 c4 NTR_MERGE_FROM=$m
done
EXTRA= ; run def c2 A=1; run def c4o A=1; run m1 c4o NTR_MERGE_FROM=1; run m1w0 c4 NTR_MERGE_FROM=1 NTR_WARP=0; EXTRA="--world 8"; run m1w0_w8 c4 NTR_MERGE_FROM=1 NTR_WARP=0; EXTRA=
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in q_c4_m1 q_c4_m1_w8 q_c4_m0_w8; do tail -1 $O/$f.err; done
