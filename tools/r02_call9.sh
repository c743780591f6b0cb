# Round 2, GPU call 9: exact mailbox for opaque scenes (config 5 / 4-opaque), then the ncu evidence on the settled kernels.
set -x
mkdir -p gpurun_out/r02i
O=gpurun_out/r02i
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -6 > $O/tests.txt
cat $O/tests.txt
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 600 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
EXTRA= ; for c in c2 c3 c4 c4b c4o c5s; do run def $c A=1; done
run nomb c5s NTR_EXACT_MAILBOX=0; run nomb c4o NTR_EXACT_MAILBOX=0
EXTRA="--frames 2"; run def c5 A=1; run nomb c5 NTR_EXACT_MAILBOX=0; EXTRA=
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
# ---- ncu: launch lists (shares of the step), then --set full of the dominant kernels ----
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python tools/quick.py c4 --frames 2 > $O/ncu_l_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv python tools/quick.py c2 --frames 2 > $O/ncu_l_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 15 -c 5 -o $O/prof_c4 python tools/quick.py c4 --frames 1 > $O/ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 3 -c 1 -o $O/prof_c2 python tools/quick.py c2 --frames 1 > $O/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 3 -c 1 -o $O/prof_c5s python tools/quick.py c5s --frames 1 > $O/ncu_c5s.log 2>&1
ls -la $O/*.ncu-rep
