# Round 2, last single-GPU call: the whole GPU suite, smoke(), both bench arms on the default workload, config 5 with the
# oracle-sample counts, launch list of the bench command.
set -x
mkdir -p gpurun_out/r02final
O=gpurun_out/r02final
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/tests.txt; cat $O/tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 600 python bench.py > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 600 $O/bench_c4.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_c4.json 2> $O/bench_ref_c4.err; tail -c 400 $O/bench_ref_c4.json
timeout 500 python bench.py --config c5 --steps 2 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 300 $O/bench_c5.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stream-frames 0 > $O/ncu_launches.log 2>&1
ls -la $O
