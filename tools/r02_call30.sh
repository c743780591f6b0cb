# Round 2, GPU call 30: tag mailbox of opaque single-simplex scenes (config 5): slots per thread, plain loop against the
# software-pipelined loop (variants/libntr_pipe.so), tree depth 17 against 20; regression check of the other opaque configs
# against the library without it (variants/libntr_base.so = HEAD before the change).
set -x
mkdir -p gpurun_out/r02zd
O=gpurun_out/r02zd
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tag_mailbox or extension_is_loaded or batched_soup" 2>&1 | tail -3 | tee $O/tests.txt
run() { local name=$1 c=$2; shift 2; env "$@" timeout 300 python tools/quick.py $c --frames 5 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
BASE=NTR_B200_LIB=$PWD/variants/libntr_base.so
PIPE=NTR_B200_LIB=$PWD/variants/libntr_pipe.so
run off c5 NTR_TAG_MAILBOX=0
for s in 256 512 1024 2048 4096; do run t$s c5 NTR_TAG_MAILBOX=$s; done
for s in 1024 2048; do run p$s c5 NTR_TAG_MAILBOX=$s $PIPE; done
run d20_off c5 NTR_TAG_MAILBOX=0 NTR_BENCH_SOUP_DEPTH=20
run d20_t2048 c5 NTR_TAG_MAILBOX=2048 NTR_BENCH_SOUP_DEPTH=20
run d20_p2048 c5 NTR_TAG_MAILBOX=2048 NTR_BENCH_SOUP_DEPTH=20 $PIPE
for c in c2 c5s c4o c1; do run new $c A=1; run base $c $BASE; run new2 $c A=1; run base2 $c $BASE; done
run forced c5s NTR_TAG_MAILBOX=1024
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
