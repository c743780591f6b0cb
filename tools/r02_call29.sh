# Round 2, GPU call 29: leaf_general without the executed re-test of the first opaque hit (new) against with it (base).
# value (new = default library) against the descriptor behind a pointer (base).
set -x
mkdir -p gpurun_out/r02zc
O=gpurun_out/r02zc
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 120 python tools/quick.py $c $EXTRA --frames 9 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for v in new base new2 base2; do
  case $v in new*) L=A=1;; base*) L=NTR_B200_LIB=$PWD/variants/libntr_base.so;; esac
  EXTRA= ; run $v c4 $L; run $v c4b $L; run $v c3 $L
  EXTRA="--world 8"; run ${v}_w8 c4 $L
done
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in $O/q_c4_new.err $O/q_c4_base.err; do echo $f; grep "pass ms" $f | tail -1; done
