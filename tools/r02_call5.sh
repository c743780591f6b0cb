# Round 2, GPU call 5: exact mailbox + run-time choice of the per-lane / warp kernels; full GPU suite; default bench.py.
set -x
mkdir -p gpurun_out/r02e
O=gpurun_out/r02e
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/tests.txt
cat $O/tests.txt
run() { # name config env...
  local name=$1 c=$2; shift 2
  env NTR_PASS_TIMING=1 "$@" timeout 300 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err
}
EXTRA=
for c in c1 c2 c3 c4b c4o c5s; do run def $c A=1; done
for w in 0 1; do for m in 0 1; do
  EXTRA= ; run w${w}m${m} c4 NTR_WARP=$w NTR_EXACT_MAILBOX=$m; run w${w}m${m} c4b NTR_WARP=$w NTR_EXACT_MAILBOX=$m
  EXTRA="--world 8"; run w${w}m${m}_w8 c4 NTR_WARP=$w NTR_EXACT_MAILBOX=$m; run w${w}m${m}_w8 c4b NTR_WARP=$w NTR_EXACT_MAILBOX=$m
done; done
EXTRA= ; run w1_c4o c4o NTR_WARP=1; run w0_c4o c4o NTR_WARP=0; run nohf c4 NTR_HEAVY_FIRST=0
EXTRA="--world 2"; run def_w2 c4 A=1; EXTRA="--world 4"; run def_w4 c4 A=1; EXTRA=
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
tail -c 3000 $O/bench_default.json
timeout 600 python bench.py --impl reference --steps 3 > $O/bench_ref.json 2> $O/bench_ref.err
tail -c 1500 $O/bench_ref.json
