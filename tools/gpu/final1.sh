mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.err
python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 300 gpurun_out/r02_bench_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'sust',round(d['sustained']['frac'],3),'cpu',round(d['cpu_baseline']['value']), {k:(round(v['value']),v['check']) for k,v in d['configs'].items()}, d['clocks'])
r=json.loads(open('gpurun_out/r02_bench_ref.json').read().strip().splitlines()[-1])
print('ref',round(r['value']),r.get('product_library_loaded'),r['config']['workload'][:60])"
for t in aligned color1080p rotated rotated1080p; do python bench.py --tex $t --no-cpu --no-e2e --configs none --sustained-seconds 0 > gpurun_out/r02_bench_tex_$t.json 2>/dev/null; python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],round(d['value']),round(d['roofline']['frac'],3),d['stitched_check'])" gpurun_out/r02_bench_tex_$t.json $t; done
