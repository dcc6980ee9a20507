# 4-GPU check of the strong-scaling arm: ONE 2^22 proof by four GPUs (b2z_dist_prove) + replicas, timeline per rank
mkdir -p gpurun_out/timeline_n4
(B2Z_TIMELINE=gpurun_out/timeline_n4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 \
   --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --no-extra \
   > gpurun_out/bench_c5_n4.json 2> gpurun_out/bench_c5_n4.err); echo bench n4 rc $?
head -c 600 gpurun_out/bench_c5_n4.json; echo; grep -v Warn gpurun_out/bench_c5_n4.err | tail -4
