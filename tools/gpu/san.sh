mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_k1.py > gpurun_out/r02_k1_memcheck.log 2>&1; tail -2 gpurun_out/r02_k1_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_k1.py > gpurun_out/r02_k1_racecheck.log 2>&1; tail -2 gpurun_out/r02_k1_racecheck.log
timeout 300 python -m pytest tests/test_k1_gpu.py tests/test_stitch_gpu.py -x -q -m gpu -k "cutoff or compaction or counted or error or float" 2>&1 | tail -2
