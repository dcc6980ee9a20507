# 2-GPU check of the strong-scaling arm on the current build: ONE 2^22 proof by both GPUs (b2z_dist_prove), replicas,
# the c2 extra with its sharded == single-GPU byte check, and a CUDA-event timeline of one proof per rank
mkdir -p gpurun_out/timeline_n2
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_n2.txt
(B2Z_TIMELINE=gpurun_out/timeline_n2 timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
   --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
   > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err); echo bench n2 rc $?
head -c 900 gpurun_out/bench_c5_n2.json; echo; tail -5 gpurun_out/bench_c5_n2.err
ls gpurun_out/timeline_n2
