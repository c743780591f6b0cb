# Round 2, GPU call 17: primary pass split in two launches -- tiles costing several mean tiles in quarter blocks by the
# warp-synchronous kernel, the rest as before on a second stream (NTR_HEAVY_TILES = cost factor; 0 = off).
set -x
mkdir -p gpurun_out/r02q
O=gpurun_out/r02q
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 60 python tools/quick.py $c $EXTRA --frames 7 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_primary or interleaved" 2>&1 | tail -4
for hf in 0 2 3 5 8; do
  EXTRA="--world 8"; run h${hf}_w8 c4 NTR_HEAVY_TILES=$hf; run h${hf}_w8 c4b NTR_HEAVY_TILES=$hf
  EXTRA="--world 4"; run h${hf}_w4 c4 NTR_HEAVY_TILES=$hf
  EXTRA="--world 2"; run h${hf}_w2 c4 NTR_HEAVY_TILES=$hf
  EXTRA= ; run h${hf}_sched c4 NTR_HEAVY_TILES=$hf NTR_TILE_SCHED=1
done
EXTRA= ; run def c4 A=1; run def c2 A=1; run def c3 A=1
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in $O/q_c4_h*_w8.err $O/q_c4_h*_sched.err $O/q_c4_def.err; do echo $f; tail -1 $f; done
