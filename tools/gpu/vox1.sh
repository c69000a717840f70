mkdir -p gpurun_out
python -m pytest tests/test_stitch_gpu.py -x -q -m gpu 2>&1 | tail -3
for sc in noise smooth; do python tools/bench_stitch.py --skip-stitch --only-voxel-variant 0 --scene $sc | tail -1; done
python tools/bench_stitch.py --skip-stitch --only-voxel-variant 0 --cams 4 | tail -1
python tools/bench_stitch.py --skip-stitch --only-voxel-variant 0 --cams 1 | tail -1
