# Round 2, GPU call 22: several chains of passes on ONE GPU (a group that lists the device more than once): each chain
# renders its interleaved tile rows with queues and launches of its own, so one chain's tail runs beside another's bulk.
set -x
mkdir -p gpurun_out/r02v
O=gpurun_out/r02v
run() { local name=$1 c=$2; shift 2; env NTR_PASS_TIMING=0 "$@" timeout 120 python tools/quick.py $c $EXTRA --frames 7 > $O/q_${c}_$name.json 2> $O/q_${c}_$name.err; }
EXTRA= ; for c in c4 c4b c2 c3 c4o c5s; do run def $c A=1; done
for ch in 0 0,0 0,0,0 0,0,0,0 0,0,0,0,0,0 0,0,0,0,0,0,0,0; do
  n=$(echo $ch | tr -cd , | wc -c); n=$((n+1))
  EXTRA="--chains $ch"; for c in c4 c4b c2 c3; do run ch$n $c A=1; done
done
EXTRA="--chains 0,0"; run ch2 c4o A=1; run ch2 c5s A=1
EXTRA="--chains 0,0,0,0"; run ch4 c4o A=1; run ch4 c5s A=1
for f in $O/q_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['ms_median'],3), round(d['ms_min'],3), d['frame_md5'][:8])" 2>/dev/null || { echo "$f FAILED"; tail -3 ${f%.json}.err; }; done
