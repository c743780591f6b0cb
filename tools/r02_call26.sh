# Round 2, GPU call 26: the 96-register build (NTR_F_WIDE) for passes below NTR_WIDE_BELOW rays, picked per pass.
set -x
mkdir -p gpurun_out/r02z
O=gpurun_out/r02z
NTR_WIDE_BELOW=100000000 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_stream.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "polytope or mixed or random" 2>&1 | tail -3
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 120 python tools/quick.py $c $EXTRA --frames 7 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for wb in 0 700000 1100000 1400000 100000000; do
  EXTRA= ; run wb$wb c4 NTR_WIDE_BELOW=$wb; run wb$wb c4b NTR_WIDE_BELOW=$wb
  EXTRA="--world 8"; run wb${wb}_w8 c4 NTR_WIDE_BELOW=$wb; run wb${wb}_w8 c4b NTR_WIDE_BELOW=$wb
  EXTRA="--world 4"; run wb${wb}_w4 c4 NTR_WIDE_BELOW=$wb
  EXTRA="--world 2"; run wb${wb}_w2 c4 NTR_WIDE_BELOW=$wb
done
EXTRA= ; run def c2 A=1; run def c3 A=1
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in $O/q_c4_wb*.err; do echo $f; grep "pass ms" $f | tail -1; done
