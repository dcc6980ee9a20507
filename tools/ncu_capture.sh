#!/bin/bash
# Profiling evidence for the default bench configuration (2^22 circuit), run under gpurun on ONE GPU.
#   1. plain run (must exit 0)   2. per-launch durations of one timed proof   3. --set full on the top kernels
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --cpu-steps 0 --no-extra"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
SKIP=$(grep -o 'before the timed region: [0-9]*' gpurun_out/ncu_plain.err | grep -o '[0-9]*$')
echo "library kernels before the timed region: $SKIP"
KREGEX='regex:msm_|ntt_|wm_|fr_from_mont|canonicalize|scan_|bitrev|pack_flags|fb_|batch_norm|twiddle|pow_table|r1cs_|spmv_|z_scatter'
# one timed proof is ~75 launches; list a little more than one
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s $SKIP -c 90 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# accumulation launches per proof, in issue order: A, B1 (G1), B (G2), L, H (G1); 2 check proofs + 3 warm-up proofs = 25 launches
# -> index 25 = A (G1), 27 = B (G2) of the first timed proof.  Transform passes: 5 per proof -> 25 = first pass (three
# vectors, high bits), 26 = the fused low pass.
ncu --set full --clock-control none --import-source on -k regex:msm_accum_kernel -s 25 -c 1 \
    -o gpurun_out/prof_accum_g1 $CMD > gpurun_out/ncu_full_accum_g1.log 2>&1
echo "full accum g1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:msm_accum_kernel -s 27 -c 1 \
    -o gpurun_out/prof_accum_g2 $CMD > gpurun_out/ncu_full_accum_g2.log 2>&1
echo "full accum g2 exit $?"
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 25 -c 2 \
    -o gpurun_out/prof_ntt $CMD > gpurun_out/ncu_full_ntt.log 2>&1
echo "full ntt exit $?"
for f in prof_accum_g1 prof_accum_g2 prof_ntt; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page details > gpurun_out/$f.details.txt 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/$f.source.csv 2>/dev/null
done
ls -la gpurun_out/ | head -40
# keep the transfer under the 64 MiB cap
for f in gpurun_out/*.ncu-rep gpurun_out/*.source.csv; do
  sz=$(stat -c %s "$f"); if [ "$sz" -gt 15000000 ]; then echo "dropping $f ($sz bytes)"; rm -f "$f"; fi
done
du -sh gpurun_out
