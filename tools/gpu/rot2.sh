set -x; mkdir -p gpurun_out
python -m pytest tests/test_k1_gpu.py -x -q -m gpu 2>&1 | tail -5
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0"
for t in baseline aligned rotated rotated1080p color1080p; do
  $B --tex $t | tail -1 > gpurun_out/r02_rot3_$t.json; python -c "
import json;d=json.load(open('gpurun_out/r02_rot3_$t.json'));print('$t',round(d['value']),d['ms_per_step'],round(d['roofline']['frac'],3),d['stitched_check'])"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_pipe -s 3 -c 1 -o gpurun_out/r02_rot_pipe $B --steps 2 --warmup 3 --no-check --tex rotated > gpurun_out/r02_rot_pipe.log 2>&1
