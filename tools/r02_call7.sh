# Round 2, GPU call 7: sparse scan of big leaves (word-grouped item tables) on top of the exact mailbox.
set -x
mkdir -p gpurun_out/r02g
O=gpurun_out/r02g
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_stream.py -m gpu -q 2>&1 | tail -6 > $O/tests.txt
cat $O/tests.txt
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 300 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
EXTRA= ; for c in c2 c3 c4 c4b c4o; do run def $c A=1; done
for w in 0 1 2; do EXTRA= ; run w$w c4 NTR_WARP=$w; EXTRA="--world 8"; run w${w}_w8 c4 NTR_WARP=$w; done
EXTRA="--world 8"; run def_w8 c4b A=1; EXTRA="--world 2"; run def_w2 c4 A=1; EXTRA="--world 4"; run def_w4 c4 A=1; EXTRA=
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in q_c4_w1 q_c4_w1_w8 q_c4_w0_w8; do tail -1 $O/$f.err; done
