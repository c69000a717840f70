set -x; mkdir -p gpurun_out
python -m pytest tests/test_k1_gpu.py -x -q -m gpu 2>&1 | tail -15
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --configs none --sustained-seconds 0"
for t in rotated rotated1080p; do
  $B --tex $t | tail -1 > gpurun_out/r02_rot2_$t.json; cut -c1-200 gpurun_out/r02_rot2_$t.json
  $B --tex $t --variant 1 | tail -1 > gpurun_out/r02_rot2_${t}_direct.json; cut -c1-200 gpurun_out/r02_rot2_${t}_direct.json
done
$B --tex baseline | tail -1 | cut -c1-200
