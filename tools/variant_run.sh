set -x
python bench.py --config c5s --steps 5 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/hi6_c5s.json 2>gpurun_out/hi6_c5s.err
for v in 5 8; do
NTR_B200_LIB=$PWD/variants/libntr_hi$v.so python bench.py --config c5s --steps 5 --warmup 3 --no-cpu-baseline --stream-frames 0 > gpurun_out/hi${v}_c5s.json 2>gpurun_out/hi${v}_c5s.err
done
python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline --stream-frames 0 > gpurun_out/mid8_c3.json 2>gpurun_out/mid8_c3.err
NTR_B200_LIB=$PWD/variants/libntr_mid6.so python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline --stream-frames 0 > gpurun_out/mid6_c3.json 2>gpurun_out/mid6_c3.err
python bench.py --config c5 --steps 2 --warmup 3 --stream-frames 0 > gpurun_out/hi6_c5.json 2>gpurun_out/hi6_c5.err
