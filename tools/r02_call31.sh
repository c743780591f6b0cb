# Round 2, GPU call 31: tag mailbox of opaque single-simplex scenes, second form: one table per WARP in dynamic shared
# memory (entries {index, lane mask}, cleared at every fetch); entries per warp; with a software prefetch of the record
# 2 / 4 items ahead (variants/libntr_pf2.so, libntr_pf4.so); tree depth 17 against 20; regression check of the other
# opaque configs against the library without any of it (variants/libntr_base.so); ncu --set full of config 5's kernel
# with and without the table.
set -x
mkdir -p gpurun_out/r02ze
O=gpurun_out/r02ze
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tag_mailbox or extension_is_loaded or batched_soup" 2>&1 | tail -3 | tee $O/tests.txt
run() { local name=$1 c=$2; shift 2; env "$@" timeout 300 python tools/quick.py $c --frames 5 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
BASE=NTR_B200_LIB=$PWD/variants/libntr_base.so
PF2=NTR_B200_LIB=$PWD/variants/libntr_pf2.so
PF4=NTR_B200_LIB=$PWD/variants/libntr_pf4.so
run off c5 NTR_TAG_MAILBOX=0
for s in 128 256 512 1024 2048; do run t$s c5 NTR_TAG_MAILBOX=$s; done
run pf2_off c5 NTR_TAG_MAILBOX=0 $PF2
run pf4_off c5 NTR_TAG_MAILBOX=0 $PF4
run pf2_t512 c5 NTR_TAG_MAILBOX=512 $PF2
run pf4_t512 c5 NTR_TAG_MAILBOX=512 $PF4
run d20_off c5 NTR_TAG_MAILBOX=0 NTR_BENCH_SOUP_DEPTH=20
run d20_t512 c5 NTR_TAG_MAILBOX=512 NTR_BENCH_SOUP_DEPTH=20
run d20_t1024 c5 NTR_TAG_MAILBOX=1024 NTR_BENCH_SOUP_DEPTH=20
for c in c2 c5s c4o; do run new $c A=1; run base $c $BASE; run new2 $c A=1; run base2 $c $BASE; done
run off c5s NTR_TAG_MAILBOX=0
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 3 -c 1 -o $O/prof_c5_tags python tools/quick.py c5 --frames 1 > $O/ncu_c5_tags.log 2>&1
NTR_TAG_MAILBOX=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 3 -c 1 -o $O/prof_c5_off python tools/quick.py c5 --frames 1 > $O/ncu_c5_off.log 2>&1
ls -la $O | tail -5
