mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'sust',round(d['sustained']['frac'],3),'cpu',round(d['cpu_baseline']['value']), {k:(round(v['value']),v['check']) for k,v in d['configs'].items()}, d['clocks'], d['gpu_launches'])"
for t in rotated rotated1080p; do python bench.py --tex $t --no-cpu --no-e2e --configs none --sustained-seconds 0 > gpurun_out/r02_bench_tex_$t.json 2>/dev/null; python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],round(d['value']),round(d['roofline']['frac'],3),d['stitched_check'])" gpurun_out/r02_bench_tex_$t.json $t; done
