mkdir -p gpurun_out
python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -c 600 gpurun_out/r02_bench_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1]);print(round(d['value']),d['ms_per_step'],d['stitched_check'],d.get('nvlink'),{k:(round(v['value']),v['check']) for k,v in d['configs'].items()}, d['e2e']['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --tex rotated --configs none --no-e2e --no-cpu --sustained-seconds 0 > gpurun_out/r02_bench_n2_rotated.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_n2_rotated.json').read().strip().splitlines()[-1]);print('rotated N=2',round(d['value']),d['ms_per_step'],d['stitched_check'],d.get('nvlink'))"
