mkdir -p gpurun_out
for tune in 0 2 4 6 8 10 14; do
echo "tune=$tune"
TIP_SEG3_TUNE=$tune ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_l.csv python tools/profile_em.py 10 800000 3 32 > /dev/null 2>&1
grep -i "seg3_pass" gpurun_out/r2g_l.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -2
done > gpurun_out/r2g_exp.txt 2>&1
cat gpurun_out/r2g_exp.txt
