# Round 2, GPU call 10: quarter-size mailbox table (keys = record index >> 2 for trees of aligned batches); settled code.
set -x
mkdir -p gpurun_out/r02j
O=gpurun_out/r02j
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/tests.txt
cat $O/tests.txt
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 600 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
EXTRA= ; for c in c1 c2 c3 c4 c4b c4o c5s; do run def $c A=1; done
EXTRA="--world 8"; run def_w8 c4 A=1; run def_w8 c4b A=1; EXTRA="--world 2"; run def_w2 c4 A=1; EXTRA="--world 4"; run def_w4 c4 A=1; EXTRA=
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
tail -1 $O/q_c4_def.err; tail -1 $O/q_c4_def_w8.err
