# Round 2, GPU call 20: split primary pass with finer units in the heavy tiles (NTR_HEAVY_UNIT pixels of a block per warp).
set -x
mkdir -p gpurun_out/r02t
O=gpurun_out/r02t
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 60 python tools/quick.py $c $EXTRA --frames 6 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_primary" 2>&1 | tail -3
EXTRA="--world 8"; run h0_w8 c4 A=1
EXTRA= ; run h0_sched c4 NTR_TILE_SCHED=1
for hf in 6 10 16 24; do for u in 4 2 1; do
  EXTRA="--world 8"; run h${hf}u${u}_w8 c4 NTR_HEAVY_TILES=$hf NTR_HEAVY_UNIT=$u
done; done
for hf in 10 24; do for u in 2 1; do
  EXTRA= ; run h${hf}u${u}_sched c4 NTR_HEAVY_TILES=$hf NTR_HEAVY_UNIT=$u NTR_TILE_SCHED=1
  EXTRA="--world 4"; run h${hf}u${u}_w4 c4 NTR_HEAVY_TILES=$hf NTR_HEAVY_UNIT=$u
done; done
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null; done
for f in $O/q_c4_*.err; do echo $f; grep "pass ms" $f | tail -1; done
