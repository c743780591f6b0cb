# Round 2, GPU call 33: config 5 with the row-major tile order (NTR_TILE_SCHED=0) against the cost-sorted one it gets by
# default (max leaf >= 256), and with the warps of a CTA fetching four adjacent blocks together (NTR_CTA_FETCH=1).
set -x
mkdir -p gpurun_out/r02zg
O=gpurun_out/r02zg
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3 | tee $O/tests_default.txt
NTR_CTA_FETCH=1 timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_facade.py -m gpu -q 2>&1 | tail -3 | tee $O/tests_cta_fetch.txt
run() { local name=$1 c=$2; shift 2; env "$@" timeout 100 python tools/quick.py $c --frames 4 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
run def c5 A=1
run rowmajor c5 NTR_TILE_SCHED=0
run cta c5 NTR_CTA_FETCH=1
run cta_rowmajor c5 NTR_CTA_FETCH=1 NTR_TILE_SCHED=0
run cta c2 NTR_CTA_FETCH=1
run def c2 A=1
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
