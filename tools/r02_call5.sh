# Final single-GPU evidence of round 2 on the current build (one gpurun call):
#   1. pytest -m gpu   2. bench.py (default: 2^22 circuit + c2 extra + cpu_baseline)   3. ncu launch list + --set full of
#   the first timed proof's A, B1 (batched-affine G1) and B (XYZZ G2) accumulations   4. A/B of the opt-in G2 variant
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/smi.txt
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
(timeout 400 python bench.py > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err); echo bench rc $?
head -c 400 gpurun_out/bench_c5_n1.json; echo
CMD="python bench.py --steps 2 --warmup 3 --cpu-steps 0 --no-extra"
timeout 200 $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; }
SKIP=$(grep -o 'before the timed region: [0-9]*' gpurun_out/ncu_plain.err | grep -o '[0-9]*$')
echo "library kernels before the timed region: $SKIP"
KREGEX='regex:msm_|ntt_|wm_|fr_from_mont|canonicalize|scan_|bitrev|pack_flags|fb_|batch_norm|twiddle|pow_table|r1cs_|spmv_|z_scatter'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s $SKIP -c 90 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# accumulation launches per proof in issue order: A, B1 (G1), B (G2), L, H (G1); 2 check + 3 warm-up proofs = 25
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msm_accum_ -s 25 -c 3 \
    -o gpurun_out/prof_accum $CMD > gpurun_out/ncu_full_accum.log 2>&1
echo "full accum exit $?"
# one report, three launches: export the A query (batched-affine G1) and the B query (G2) separately
for spec in "prof_accum_affine_g1 0" "prof_accum_g2 2"; do
  set -- $spec
  ncu -i gpurun_out/prof_accum.ncu-rep -s $2 -c 1 --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_accum.ncu-rep -s $2 -c 1 --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/prof_accum.ncu-rep -s $2 -c 1 --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
done
ncu -i gpurun_out/prof_accum.ncu-rep --page raw --csv > gpurun_out/prof_accum_all3.raw.csv 2>/dev/null
(B2Z_AFFINE_G2=1 timeout 200 $CMD > gpurun_out/bench_c5_affine_g2.json 2> gpurun_out/bench_c5_affine_g2.err); echo g2 rc $?
head -c 300 gpurun_out/bench_c5_affine_g2.json; echo
for f in gpurun_out/*.ncu-rep gpurun_out/*.source.csv; do
  sz=$(stat -c %s "$f"); if [ "$sz" -gt 20000000 ]; then echo "dropping $f ($sz bytes)"; rm -f "$f"; fi
done
du -sh gpurun_out
