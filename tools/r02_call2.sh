# Round 2, GPU call 2: the warp-synchronous path (trace_warp.cuh) -- parity suite, then timings of its variants.
set -x
mkdir -p gpurun_out/r02b
O=gpurun_out/r02b
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.txt
cat $O/tests.txt
for c in c2 c3 c4 c4o c4b c5s c1; do NTR_PASS_TIMING=1 timeout 300 python tools/quick.py $c > $O/q_$c.json 2> $O/q_$c.err; done
for c in c4 c4b c2; do NTR_PASS_TIMING=1 timeout 300 python tools/quick.py $c --world 8 > $O/q_${c}_w8.json 2> $O/q_${c}_w8.err; done
for v in w80 w96 lm96 lm32 oh8 nocoop; do
  [ -f variants/libntr_$v.so ] || continue
  for c in c2 c4 c4o c4b; do NTR_PASS_TIMING=1 NTR_B200_LIB=$PWD/variants/libntr_$v.so timeout 300 python tools/quick.py $c > $O/q_${c}_$v.json 2> $O/q_${c}_$v.err; done
  NTR_PASS_TIMING=1 NTR_B200_LIB=$PWD/variants/libntr_$v.so timeout 300 python tools/quick.py c4 --world 8 > $O/q_c4_w8_$v.json 2> $O/q_c4_w8_$v.err
done
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', d['ms_median'], d['ms_min'], d['frame_md5'][:8])" 2>/dev/null; done
grep -h "pass ms" $O/q_c4.err | tail -1
grep -h "pass ms" $O/q_c4_w8.err | tail -1
