# Round 2, GPU call 1: full GPU test suite (no -x), per-pass timings of the 4K frames, the round-1 cooperative variant
# (never run before), and the two ncu --set full captures round 1 did not get to.
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/tests1.txt
for c in c2 c3 c4 c4o c4b; do NTR_PASS_TIMING=1 timeout 300 python tools/quick.py $c > $O/q_$c.json 2> $O/q_$c.err; done
NTR_PASS_TIMING=1 timeout 300 python tools/quick.py c4 --world 8 > $O/q_c4_w8.json 2> $O/q_c4_w8.err
NTR_PASS_TIMING=1 timeout 300 python tools/quick.py c4b --world 8 > $O/q_c4b_w8.json 2> $O/q_c4b_w8.err
for v in coop; do
  [ -f variants/libntr_$v.so ] || continue
  for c in c2 c4 c4o c4b; do NTR_PASS_TIMING=1 NTR_B200_LIB=$PWD/variants/libntr_$v.so timeout 300 python tools/quick.py $c > $O/q_${c}_$v.json 2> $O/q_${c}_$v.err; done
  NTR_PASS_TIMING=1 NTR_B200_LIB=$PWD/variants/libntr_$v.so timeout 300 python tools/quick.py c4 --world 8 > $O/q_c4_w8_$v.json 2> $O/q_c4_w8_$v.err
  NTR_B200_LIB=$PWD/variants/libntr_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > $O/tests_$v.txt
done
for v in r80 r96; do
  [ -f variants/libntr_$v.so ] || continue
  for c in c2 c4 c4o; do NTR_B200_LIB=$PWD/variants/libntr_$v.so timeout 300 python tools/quick.py $c > $O/q_${c}_$v.json 2> $O/q_${c}_$v.err; done
done
# ncu --set full: general-variant kernel on the 4K frame (primary pass + first bounce pass of the 4th frame), 10-D kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 15 -c 2 -o $O/prof_c4 python tools/quick.py c4 --frames 1 > $O/ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 3 -c 1 -o $O/prof_c5s python tools/quick.py c5s --frames 1 > $O/ncu_c5s.log 2>&1
timeout 300 python tools/quick.py c5s > $O/q_c5s.json 2> $O/q_c5s.err
for v in sm64 sm256; do
  [ -f variants/libntr_$v.so ] || continue
  NTR_B200_LIB=$PWD/variants/libntr_$v.so timeout 300 python tools/quick.py c5s > $O/q_c5s_$v.json 2> $O/q_c5s_$v.err
done
cat $O/tests1.txt
for f in $O/q_*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['ms_median'], d['ms_min'])" 2>/dev/null; done
grep -h "pass ms" $O/q_c4.err | tail -2
