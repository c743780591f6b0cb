# Round 2, GPU call 3: which parts of the warp-synchronous path pay (variants), adaptive fetch, ncu of the new kernels.
set -x
mkdir -p gpurun_out/r02c
O=gpurun_out/r02c
run() { # name lib config extra-args env...
  local name=$1 lib=$2 c=$3; shift 3
  if [ "$lib" = default ]; then env NTR_PASS_TIMING=1 "$@" timeout 300 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err
  else env NTR_PASS_TIMING=1 NTR_B200_LIB=$PWD/variants/libntr_$lib.so "$@" timeout 300 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; fi
}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_stream.py -m gpu -q -x 2>&1 | tail -5 > $O/tests.txt
cat $O/tests.txt
for lib in default v0 v1 v1b v2b v2c v1r v1f; do
  EXTRA= ; for c in c2 c4 c4o c4b; do run $lib $lib $c A=1; done
  EXTRA="--world 8"; run ${lib}_w8 $lib c4 A=1; run ${lib}_w8 $lib c4b A=1; EXTRA=
done
for lib in default v1 v0; do
  EXTRA= ; run ${lib}_noaf $lib c4 NTR_ADAPTIVE_FETCH=0; run ${lib}_nohf $lib c4 NTR_HEAVY_FIRST=0
  EXTRA="--world 8"; run ${lib}_noaf_w8 $lib c4 NTR_ADAPTIVE_FETCH=0; run ${lib}_nohf_w8 $lib c4 NTR_HEAVY_FIRST=0; EXTRA=
done
EXTRA= ; run default default c3 A=1; run default default c5s A=1; run v0 v0 c3 A=1; run v0 v0 c5s A=1; run v1 v1 c3 A=1; run v1 v1 c5s A=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 3 -c 1 -o $O/prof_c2_warp python tools/quick.py c2 --frames 1 > $O/ncu_c2.log 2>&1
NTR_B200_LIB=$PWD/variants/libntr_v0.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 3 -c 1 -o $O/prof_c2_v0 python tools/quick.py c2 --frames 1 > $O/ncu_c2_v0.log 2>&1
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
