mkdir -p gpurun_out
python tools/sanitize_k1.py
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_k1.py > gpurun_out/r02_k1_memcheck.log 2>&1; tail -4 gpurun_out/r02_k1_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_k1.py > gpurun_out/r02_k1_racecheck.log 2>&1; tail -4 gpurun_out/r02_k1_racecheck.log
