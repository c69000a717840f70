mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --tex rotated --configs none --no-e2e --no-cpu --sustained-seconds 0 > gpurun_out/r02_bench_n2_rotated.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_n2_rotated.json').read().strip().splitlines()[-1]);print('rotated N=2',round(d['value']),d['ms_per_step'],d['stitched_check'],round(d['nvlink']['recv_GBps_per_gpu']))"
timeout 300 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -2
