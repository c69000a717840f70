set -x; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0"
for t in rotated rotated1080p baseline; do
  $B --tex $t > gpurun_out/r02_bench_tex_$t.json; python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],round(d['value']),round(d['ms_per_step'],4),round(d['roofline']['frac'],3),d['stitched_check'])" gpurun_out/r02_bench_tex_$t.json $t
done
