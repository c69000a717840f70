mkdir -p gpurun_out
B="timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],round(d['value']),round(d['ms_per_step'],4),round(d['roofline']['frac'],3),d['stitched_check'])" $1 "$2"; }
$B --tex rotated > gpurun_out/tmp.json; show gpurun_out/tmp.json "rotated tmap"
PCS_PIPE_TMAP=0 $B --tex rotated > gpurun_out/tmp.json; show gpurun_out/tmp.json "rotated rows"
$B --tex rotated1080p > gpurun_out/tmp.json; show gpurun_out/tmp.json "rotated1080p tmap"
timeout 300 python -m pytest tests/test_k1_gpu.py -x -q -m gpu -k "rot_small or translate_yz or guarded or transform_changed" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_k1_gpu.py -x -q -m gpu 2>&1 | tail -3
