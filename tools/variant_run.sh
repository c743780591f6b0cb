set -x
timeout 600 python -m pytest tests/test_stream.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/stream_tests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/v_base_c2.json 2>gpurun_out/v_base_c2.err
python bench.py --config c4o --steps 5 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/v_base_c4o.json 2>gpurun_out/v_base_c4o.err
NTR_B200_LIB=$PWD/variants/libntr_smemaxis.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --stream-frames 0 > gpurun_out/v_smem_c2.json 2>gpurun_out/v_smem_c2.err
NTR_B200_LIB=$PWD/variants/libntr_smemaxis.so python bench.py --config c4o --steps 5 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/v_smem_c4o.json 2>gpurun_out/v_smem_c4o.err
NTR_B200_LIB=$PWD/variants/libntr_smemaxis.so python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/v_smem_c3.json 2>gpurun_out/v_smem_c3.err
python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/v_base_c3.json 2>gpurun_out/v_base_c3.err
cat gpurun_out/stream_tests.txt
for f in gpurun_out/v_*.json; do python - $f <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d.get('stream'))
except Exception as e: print(sys.argv[1],'ERR',e)
P
done
