# Round 2, GPU call 4: per-lane path vs warp path after the abort-poll fix, fetch sizes per cost ring, fetch-duration
# histograms of the 1/8-frame passes (what bounds 8-GPU scaling), the new multi-GPU / shared-frame tests on one GPU.
set -x
mkdir -p gpurun_out/r02d
O=gpurun_out/r02d
run() { # name lib config env...
  local name=$1 lib=$2 c=$3; shift 3
  env NTR_PASS_TIMING=1 NTR_B200_LIB=$PWD/variants/libntr_$lib.so "$@" timeout 300 python tools/quick.py $c $EXTRA > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err
}
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > $O/tests.txt
cat $O/tests.txt
for lib in p0 w1 w2 w1r; do
  EXTRA= ; for c in c2 c3 c4 c4o c4b c5s; do run $lib $lib $c A=1; done
  EXTRA="--world 8"; run ${lib}_w8 $lib c4 A=1; run ${lib}_w8 $lib c4b A=1; EXTRA=
done
for lib in p0 w1; do
  EXTRA= ; run ${lib}_noaf $lib c4 NTR_ADAPTIVE_FETCH=0; run ${lib}_nohf $lib c4 NTR_HEAVY_FIRST=0
  run ${lib}_f124 $lib c4 NTR_FETCH_SIZES=1,2,4; run ${lib}_f248 $lib c4 NTR_FETCH_SIZES=2,4,8; run ${lib}_f81632 $lib c4 NTR_FETCH_SIZES=8,16,32
  EXTRA="--world 8"; run ${lib}_noaf_w8 $lib c4 NTR_ADAPTIVE_FETCH=0; run ${lib}_nohf_w8 $lib c4 NTR_HEAVY_FIRST=0
  run ${lib}_f124_w8 $lib c4 NTR_FETCH_SIZES=1,2,4; run ${lib}_f248_w8 $lib c4 NTR_FETCH_SIZES=2,4,8; run ${lib}_f81632_w8 $lib c4 NTR_FETCH_SIZES=8,16,32; EXTRA=
done
EXTRA="--world 8 --frames 1"
for lib in p0s w1s; do run ${lib}_w8 $lib c4 A=1; run ${lib}_noaf_w8 $lib c4 NTR_ADAPTIVE_FETCH=0; run ${lib}_f124_w8 $lib c4 NTR_FETCH_SIZES=1,2,4; done
EXTRA="--frames 1"; run p0s p0s c4 A=1; run p0s p0s c4b A=1
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
