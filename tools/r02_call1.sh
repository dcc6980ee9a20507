mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/smi.txt
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
(timeout 600 python bench.py > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err); echo bench rc $?
timeout 900 bash tools/ncu_capture.sh > gpurun_out/ncu_capture.log 2>&1; echo ncu rc $?
tail -3 gpurun_out/pytest_gpu.log; head -c 600 gpurun_out/bench_c5_n1.json
