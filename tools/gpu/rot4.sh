set -x; mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],round(d['value']),round(d['ms_per_step'],4),round(d['roofline']['frac'],3),d['stitched_check'])" $1 "$2"; }
for t in rotated rotated1080p; do
  for rt in 0 1 2; do
    PCS_PIPE_RT=$rt $B --tex $t > gpurun_out/tmp.json; show gpurun_out/tmp.json "$t seg128 rt=$rt"
    PCS_PIPE_RT=$rt PCS_B200_LIB=$PWD/pointcloud_stitching_b200/libpcs_b200_seg256.so $B --tex $t > gpurun_out/tmp.json; show gpurun_out/tmp.json "$t seg256 rt=$rt"
  done
done
