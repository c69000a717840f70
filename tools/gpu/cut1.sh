mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_k1_geom.py cutoff 2>&1 | grep case | cut -c1-330
