mkdir -p gpurun_out
export TIP_SEG3_TUNE=1
python tools/profile_em.py 10 800000 4 32 > gpurun_out/r2e_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:seg3_ -s 4 -c 4 -o gpurun_out/r2e_seg3 python tools/profile_em.py 10 800000 4 32 > gpurun_out/r2e_ncu_full.log 2>&1
tail -3 gpurun_out/r2e_ncu_full.log; cat gpurun_out/r2e_plain.log
