# consolidated single-GPU measurement run (profiles/r01_*)
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1
python bench.py --impl reference --steps 3 > gpurun_out/final_ref_c2.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > gpurun_out/final_c2.json 2>gpurun_out/final_c2.err
for c in c1 c3 c5s; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/final_$c.json 2>gpurun_out/final_$c.err; done
python bench.py --config c5 --steps 2 --warmup 3 > gpurun_out/final_c5.json 2>gpurun_out/final_c5.err
if [ -n "$WITH_C4" ]; then for c in c4 c4o; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/final_$c.json 2>gpurun_out/final_$c.err; done; fi
# ncu --set full of the general-variant kernel (config 4: primary pass + first bounce pass of the second frame) and of the 10-D kernel
python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_pass -s 5 -c 2 -o gpurun_out/final_prof_c4 python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/ncu_c4.log 2>&1
python bench.py --config c5s --steps 1 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/plain_c5s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_pass -s 2 -c 1 -o gpurun_out/final_prof_c5s python bench.py --config c5s --steps 1 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/ncu_c5s.log 2>&1
cat gpurun_out/final_tests.txt gpurun_out/final_smoke.txt
ls -la gpurun_out | tail -20
