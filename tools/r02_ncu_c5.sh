# Round 2: ncu --set full of config 5's kernel (1 M ten-dimensional simplexes) on a 960x540 frame of the same view -- a 4K
# frame of it (1.25 s per launch, 74 replay passes) does not fit a GPU call.
set -x
mkdir -p gpurun_out/r02c5
O=gpurun_out/r02c5
timeout 200 python tools/quick.py c5 --size 960x540 --frames 2 > $O/plain.json 2> $O/plain.err && cat $O/plain.json | cut -c1-300 && \
timeout 330 ncu --set full --clock-control none --import-source on -k regex:render_pass -s 3 -c 1 -o $O/prof_c5 python tools/quick.py c5 --size 960x540 --frames 1 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls -la $O
