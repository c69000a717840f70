mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -c 400 gpurun_out/r02_bench_n8.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1]);print(round(d['value']),d['ms_per_step'],d['stitched_check'],round(d['nvlink']['recv_GBps_per_gpu']),{k:(round(v['value']),v['check']) for k,v in d['configs'].items()}, round(d['e2e']['value']))"
