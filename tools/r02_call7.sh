# configs[3] sweep on ONE GPU with the current build (every result checked inside tools/sweep.py)
mkdir -p gpurun_out
(timeout 170 python tools/sweep.py --ntt 16,18,20,22,24 --wm 16,17,18,20,22 --msm 16,18,20,22,24 --g2 20 \
   --mixes uniform,witness --adversarial 20 --reps 3 > gpurun_out/sweep_n1.jsonl 2> gpurun_out/sweep_n1.err); echo sweep rc $?
wc -l gpurun_out/sweep_n1.jsonl; tail -3 gpurun_out/sweep_n1.err; cut -c1-260 gpurun_out/sweep_n1.jsonl | tail -12
