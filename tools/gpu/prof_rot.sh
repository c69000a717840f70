set -x; mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --configs none --sustained-seconds 0 --no-check --tex rotated"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_direct -s 3 -c 1 -o gpurun_out/r02_rot_direct $B --variant 1 > gpurun_out/r02_rot_direct.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_pipe -s 3 -c 1 -o gpurun_out/r02_rot_pipe $B --variant 2 > gpurun_out/r02_rot_pipe.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0 --tex rotated --variant 2 | tail -1 > gpurun_out/r02_rot_pipe_bench.json
tail -2 gpurun_out/r02_rot_direct.log gpurun_out/r02_rot_pipe.log
cut -c1-400 gpurun_out/r02_rot_pipe_bench.json
