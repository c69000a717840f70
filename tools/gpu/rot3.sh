set -x; mkdir -p gpurun_out
python -m pytest tests/test_k1_gpu.py -x -q -m gpu 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],round(d['value']),round(d['ms_per_step'],4),round(d['roofline']['frac'],3),d['stitched_check'])" $1 "$2"; }
for t in rotated rotated1080p; do
  $B --tex $t > gpurun_out/r02_rot4_$t.json; show gpurun_out/r02_rot4_$t.json $t
  PCS_PIPE_STAGES=2 $B --tex $t > gpurun_out/tmp.json; show gpurun_out/tmp.json "$t stages=2"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_pipe -s 3 -c 1 -o gpurun_out/r02_rot_pipe2 $B --steps 2 --warmup 3 --no-check --tex rotated > gpurun_out/r02_rot_pipe2.log 2>&1
