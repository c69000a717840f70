mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0 --no-check"
for t in rotated rotated1080p; do
PCS_B200_LIB=$PWD/pointcloud_stitching_b200/libpcs_b200_probe.so $B --tex $t > gpurun_out/tmp.json; python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],round(d['value']),round(d['ms_per_step'],4),round(d['roofline']['frac'],3))" gpurun_out/tmp.json "$t noguard"
done
python -m pytest tests/test_k1_gpu.py -x -q -m gpu 2>&1 | tail -2
