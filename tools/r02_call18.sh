# Round 2, GPU call 18: where the heaviest blocks of a 1/8 share spend their time -- shadows on/off, transparency on/off,
# per-fetch durations (NTR_FETCH_STATS build).
set -x
mkdir -p gpurun_out/r02r
O=gpurun_out/r02r
L=$PWD/variants/libntr_stats.so
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=1 "$@" timeout 60 python tools/quick.py $c $EXTRA --frames 3 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
for c in c4 c4o; do
  EXTRA="--world 8"; run w8 $c A=1; run w8_stats $c NTR_B200_LIB=$L
  EXTRA="--world 8 --param 1=0"; run w8_noshadow $c A=1; run w8_noshadow_stats $c NTR_B200_LIB=$L
  EXTRA="--world 8 --param 3=0"; run w8_depth0 $c A=1; run w8_depth0_stats $c NTR_B200_LIB=$L
  EXTRA="--world 8 --param 3=0 --param 1=0"; run w8_depth0_noshadow $c A=1
  EXTRA="--param 1=0"; run noshadow $c A=1
done
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['counters'])" 2>/dev/null; done
for f in $O/q_*.err; do echo $f; grep "pass ms" $f | tail -1; grep "fetch stats" $f | tail -5 | cut -c1-400; done
