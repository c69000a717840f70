mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --sustained-seconds 0 --no-check"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);c=d['configs'];print(sys.argv[2],'c3',round(c['c3']['value']),round(c['c3']['merge_ms_per_frame'],4),'c5',round(c['c5']['value']),round(c['c5']['merge_ms_per_frame'],4))" $1 "$2"; }
for b in 1 0; do for lanes in 4 8; do
PCS_SW_BALLOT=$b $B --merge-lanes $lanes > gpurun_out/tmp.json 2>gpurun_out/tmp.err; show gpurun_out/tmp.json "ballot=$b lanes=$lanes"
done; done
python tools/bench_stitch.py --skip-stitch --only-voxel-variant 0 | tail -1
PCS_SW_BALLOT=0 python tools/bench_stitch.py --skip-stitch --only-voxel-variant 0 | tail -1
