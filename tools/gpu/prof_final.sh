mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --configs none --sustained-seconds 0 --no-check"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_pipe -s 3 -c 1 -o gpurun_out/r02_k1_pipe $B > gpurun_out/r02_k1_pipe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_pipe -s 3 -c 1 -o gpurun_out/r02_k1_rot $B --tex rotated > gpurun_out/r02_k1_rot.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_launches.log 2>&1
tail -2 gpurun_out/r02_launches.log | cut -c1-200; wc -l gpurun_out/r02_launches.csv
