set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/gpu_tests.txt
cat gpurun_out/gpu_tests.txt
python bench.py --config c5s --steps 5 --warmup 3 --stream-frames 0 > gpurun_out/d10_c5s.json 2>gpurun_out/d10_c5s.err
NTR_FORCE_GENERIC=1 python bench.py --config c5s --steps 5 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/gen_c5s.json 2>gpurun_out/gen_c5s.err
python bench.py --config c5 --steps 2 --warmup 3 --stream-frames 0 > gpurun_out/d10_c5.json 2>gpurun_out/d10_c5.err
