# consolidated single-GPU measurement run (profiles/r01_*)
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1
python bench.py --impl reference --steps 3 > gpurun_out/final_ref_c2.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > gpurun_out/final_c2.json 2>gpurun_out/final_c2.err
for c in c1 c3 c4 c4o c5s; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/final_$c.json 2>gpurun_out/final_$c.err; done
python bench.py --config c5 --steps 2 --warmup 3 > gpurun_out/final_c5.json 2>gpurun_out/final_c5.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/final_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_pass -s 4 -c 1 -o gpurun_out/final_prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/final_launches_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out | tail -30
