mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --sustained-seconds 0 --no-check"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);c=d['configs'];print(sys.argv[2],'c3',round(c['c3']['value']),round(c['c3']['merge_ms_per_frame'],4),'c5',round(c['c5']['value']),round(c['c5']['merge_ms_per_frame'],4), c['c3']['check'], c['c5']['check'])" $1 "$2"; }
for lanes in 4 6 8; do for ns in 512 2048; do
PCS_SW_POLL_NS=$ns $B --merge-lanes $lanes > gpurun_out/tmp.json 2>gpurun_out/tmp.err; show gpurun_out/tmp.json "lanes=$lanes poll=$ns"
done; done
