# Round 2, GPU call 24: what the driver runs at round end, on the settled code -- GPU suite with -x, smoke(), default
# bench line -- plus the ncu evidence of the same commands (launch list of bench.py, --set full of one config-4 frame).
set -x
mkdir -p gpurun_out/r02x
O=gpurun_out/r02x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $O/tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $O/smoke.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --stream-frames 0 > $O/ncu_l_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 15 -c 5 -o $O/prof_c4 python tools/quick.py c4 --frames 1 > $O/ncu_c4.log 2>&1
ls -la $O/*.ncu-rep
