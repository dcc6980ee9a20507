mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_affine.py -x -q 2>&1 | tail -25) > gpurun_out/pytest_affine.log
tail -3 gpurun_out/pytest_affine.log
(timeout 600 python tools/ab_accum.py --size 64 --steps ${AB_STEPS:-4} --settings ${AB_SETTINGS:-xyzz,affine} > gpurun_out/ab_c5.jsonl 2> gpurun_out/ab_c5.err); echo ab rc $?
cat gpurun_out/ab_c5.jsonl | cut -c1-620; tail -3 gpurun_out/ab_c5.err
if [ "${WITH_NCU:-1}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:msm_accum_affine_kernel -s 4 -c 1 \
    -o gpurun_out/prof_aff_g1 python tools/ab_accum.py --size 64 --steps 1 --settings affine > gpurun_out/ncu_aff.log 2>&1
echo ncu rc $?
ncu -i gpurun_out/prof_aff_g1.ncu-rep --page raw --csv > gpurun_out/prof_aff_g1.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_aff_g1.ncu-rep --page details > gpurun_out/prof_aff_g1.details.txt 2>/dev/null
ncu -i gpurun_out/prof_aff_g1.ncu-rep --page source --csv > gpurun_out/prof_aff_g1.source.csv 2>/dev/null
rm -f gpurun_out/prof_aff_g1.ncu-rep
fi
