mkdir -p gpurun_out
for skip in 0 1 2 4 8 16 31; do
  echo "skip=$skip"; TIP_SEG3_SKIP=$skip timeout 120 python tools/hub_probe.py --flags 32 --steps 10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['shape'],round(d['ms_per_iteration'],4))"
done > gpurun_out/r2h_skip.txt 2>&1
cat gpurun_out/r2h_skip.txt
for tune in 0 1; do echo "tune=$tune"; TIP_SEG3_TUNE=$tune ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_l.csv python tools/profile_em.py 10 800000 3 32 > /dev/null 2>&1
grep -i "seg3_pass" gpurun_out/r2h_l.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -2; done
