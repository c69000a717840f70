mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_k1_gpu.py -x -q -m gpu 2>&1 | tail -2
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],round(d['value']),round(d['ms_per_step'],4),round(d['roofline']['frac'],3),d['stitched_check'],d['clocks']['sm_mhz'])" $1 "$2"; }
for rep in 1 2; do for t in rotated rotated1080p; do
    $B --tex $t > gpurun_out/tmp.json; show gpurun_out/tmp.json "$t new"
    PCS_B200_LIB=$PWD/pointcloud_stitching_b200/libpcs_b200_base.so $B --tex $t > gpurun_out/tmp.json; show gpurun_out/tmp.json "$t base"
done; done
