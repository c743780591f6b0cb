# Round 2, GPU call 32: single-simplex test above 8 dimensions with aligned float4 loads and the next edge requested ahead
# (new = default library) against the scalar loads behind the early-exit branches (base = variants/libntr_base.so).
set -x
mkdir -p gpurun_out/r02zf
O=gpurun_out/r02zf
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3 | tee $O/tests.txt
run() { local name=$1 c=$2; shift 2; env "$@" timeout 120 python tools/quick.py $c --frames 4 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
BASE=NTR_B200_LIB=$PWD/variants/libntr_base.so
run new c5 A=1; run base c5 $BASE
run new c5s A=1; run base c5s $BASE
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
