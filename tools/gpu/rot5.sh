mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],round(d['value']),round(d['ms_per_step'],4),round(d['roofline']['frac'],3),d['stitched_check'])" $1 "$2"; }
for t in rotated rotated1080p; do
    $B --tex $t > gpurun_out/tmp.json; show gpurun_out/tmp.json "$t regs80"
    PCS_B200_LIB=$PWD/pointcloud_stitching_b200/libpcs_b200_r104.so $B --tex $t > gpurun_out/tmp.json; show gpurun_out/tmp.json "$t regs104"
    PCS_PIPE_RT=2 PCS_PIPE_STAGES=2 $B --tex $t > gpurun_out/tmp.json; show gpurun_out/tmp.json "$t rt2 s2 regs88"
done
