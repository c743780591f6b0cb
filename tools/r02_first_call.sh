# First GPU call of the next round: what the last call of round 1 could not finish, plus the A/B runs prepared at the
# end of round 1.  Build the variants first (here, no GPU needed):
#   tools/build_variants.sh sm64="-DNTR_SINGLE_MAILBOX=64" sm256="-DNTR_SINGLE_MAILBOX=256"
# then:  gpurun --timeout 1500 -- 'bash tools/r02_first_call.sh'
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_tests.txt
# config 5 A/B: single-simplex mailbox sizes, deeper tree
python bench.py --config c5s --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5s_base.json 2>gpurun_out/r02_c5s_base.err
python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_base.json 2>gpurun_out/r02_c5_base.err
for v in sm64 sm256; do
  [ -f variants/libntr_$v.so ] || continue
  NTR_B200_LIB=$PWD/variants/libntr_$v.so python bench.py --config c5s --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5s_$v.json 2>gpurun_out/r02_c5s_$v.err
  NTR_B200_LIB=$PWD/variants/libntr_$v.so python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_$v.json 2>gpurun_out/r02_c5_$v.err
done
NTR_BENCH_SOUP_DEPTH=20 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_depth20.json 2>gpurun_out/r02_c5_depth20.err
# longest-first tile hand-out on one GPU for the heavy-tailed scene (model: up to -29 %)
for c in c4 c4o; do
  python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_${c}_base.json 2>gpurun_out/r02_${c}_base.err
  NTR_TILE_SCHED=1 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_${c}_lpt.json 2>gpurun_out/r02_${c}_lpt.err
  NTR_TILE_SCHED=1 NTR_HEAVY_FIRST=1 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_${c}_lpt_heavy.json 2>gpurun_out/r02_${c}_lpt_heavy.err
done
# warp-cooperative big leaves (general + opaque variants), never run so far: parity first, then the numbers
#   tools/build_variants.sh coop="-DNTR_COOP_LEAVES=1 -DNTR_MIN_CTAS=6"
if [ -f variants/libntr_coop.so ]; then
  NTR_B200_LIB=$PWD/variants/libntr_coop.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_coop_tests.txt
  for c in c4 c4o c2; do NTR_B200_LIB=$PWD/variants/libntr_coop.so timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_${c}_coop.json 2>gpurun_out/r02_${c}_coop.err; done
fi
# the two ncu captures that were queued at the end of round 1
python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/plain_c4.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 5 -c 2 -o gpurun_out/r02_prof_c4 python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/ncu_c4.log 2>&1
python bench.py --config c5s --steps 1 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/plain_c5s.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 2 -c 1 -o gpurun_out/r02_prof_c5s python bench.py --config c5s --steps 1 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/ncu_c5s.log 2>&1
# config 4 under the symbol BASELINE.json spells ({5/2,5,3}, fixture added after the last GPU call of round 1), and
# full lines (cpu baselines included) for the configs whose round-1 profiles carry fields from two runs
for c in c4b c3 c5s; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r02_$c.json 2>gpurun_out/r02_$c.err; done
cat gpurun_out/r02_tests.txt
