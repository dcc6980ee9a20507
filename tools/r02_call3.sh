mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:msm_accum_affine_kernel -s 4 -c 1 \
    -o gpurun_out/prof_aff_g1 python tools/ab_accum.py --size 64 --steps 1 --settings affine > gpurun_out/ncu_aff.log 2>&1
echo ncu rc $?
ncu -i gpurun_out/prof_aff_g1.ncu-rep --page raw --csv > gpurun_out/prof_aff_g1.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_aff_g1.ncu-rep --page details > gpurun_out/prof_aff_g1.details.txt 2>/dev/null
ncu -i gpurun_out/prof_aff_g1.ncu-rep --page source --csv > gpurun_out/prof_aff_g1.source.csv 2>/dev/null
rm -f gpurun_out/prof_aff_g1.ncu-rep
ls -la gpurun_out | tail -5
