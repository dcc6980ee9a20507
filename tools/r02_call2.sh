mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_affine.py -x -q 2>&1 | tail -25) > gpurun_out/pytest_affine.log
tail -5 gpurun_out/pytest_affine.log
(timeout 600 python tools/ab_accum.py --size 64 --steps 4 > gpurun_out/ab_c5.jsonl 2> gpurun_out/ab_c5.err); echo ab rc $?
cat gpurun_out/ab_c5.jsonl | cut -c1-700; tail -3 gpurun_out/ab_c5.err
