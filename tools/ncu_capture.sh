#!/bin/bash
# Profiling evidence for the default bench configuration (2^22 circuit), run under gpurun on ONE GPU.
#   1. plain run (must exit 0)   2. per-launch durations of one timed proof   3. --set full of the first timed proof's
#   A, B1 (batched-affine G1) and B (G2) accumulations in ONE report, exported per kernel
# Summaries for profiles/: python tools/ncu_summary.py gpurun_out profiles/<round>_ncu
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --cpu-steps 0 --no-extra"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
SKIP=$(grep -o 'before the timed region: [0-9]*' gpurun_out/ncu_plain.err | grep -o '[0-9]*$')
echo "library kernels before the timed region: $SKIP"
KREGEX='regex:msm_|ntt_|wm_|fr_from_mont|canonicalize|scan_|bitrev|pack_flags|fb_|batch_norm|twiddle|pow_table|r1cs_|spmv_|z_scatter'
# one timed proof is 90 launches
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s $SKIP -c 90 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# accumulation launches per proof in issue order: A, B1 (G1), B (G2), L, H (G1); 2 check + 3 warm-up proofs = 25 launches
# -> 25 = A, 26 = B1, 27 = B of the first timed proof (msm_accum_affine_kernel and msm_accum_kernel both match)
ncu --set full --clock-control none --import-source on -k regex:msm_accum_ -s 25 -c 3 \
    -o gpurun_out/prof_accum $CMD > gpurun_out/ncu_full_accum.log 2>&1
echo "full accum exit $?"
for spec in "prof_accum_affine_g1 0" "prof_accum_g2 2"; do
  set -- $spec
  ncu -i gpurun_out/prof_accum.ncu-rep -s $2 -c 1 --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_accum.ncu-rep -s $2 -c 1 --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/prof_accum.ncu-rep -s $2 -c 1 --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
done
# transform passes: 5 per proof -> 25 = the first pass of the first timed proof (three vectors), 26 = the fused low pass
if [ "${WITH_NTT:-0}" = "1" ]; then
  ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 25 -c 2 \
      -o gpurun_out/prof_ntt $CMD > gpurun_out/ncu_full_ntt.log 2>&1
  echo "full ntt exit $?"
  ncu -i gpurun_out/prof_ntt.ncu-rep --page raw --csv > gpurun_out/prof_ntt.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_ntt.ncu-rep --page details > gpurun_out/prof_ntt.details.txt 2>/dev/null
  ncu -i gpurun_out/prof_ntt.ncu-rep --page source --csv > gpurun_out/prof_ntt.source.csv 2>/dev/null
fi
# keep the transfer under the 64 MiB cap
for f in gpurun_out/*.ncu-rep gpurun_out/*.source.csv; do
  sz=$(stat -c %s "$f"); if [ "$sz" -gt 20000000 ]; then echo "dropping $f ($sz bytes)"; rm -f "$f"; fi
done
du -sh gpurun_out
