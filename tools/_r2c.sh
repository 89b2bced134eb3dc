mkdir -p gpurun_out
for tune in 0 1; do for ch in 1 2 4 8; do
  echo "tune=$tune chunk=$ch"; TIP_SEG3_TUNE=$tune TIP_SEG3_CHUNK=$ch timeout 120 python tools/hub_probe.py --flags 32 --steps 10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['shape'],round(d['ms_per_iteration'],4))"
done; done > gpurun_out/r2c_tune.txt 2>&1
cat gpurun_out/r2c_tune.txt
python tools/profile_em.py 10 800000 4 32 > gpurun_out/r2c_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches.csv python tools/profile_em.py 10 800000 4 32 > gpurun_out/r2c_ncu_list.log 2>&1
cat gpurun_out/r2c_plain.log
python tools/profile_em.py 10 800000 4 32 > gpurun_out/r2c_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:seg3_pass -s 2 -c 2 -o gpurun_out/r2c_seg3_pass python tools/profile_em.py 10 800000 4 32 > gpurun_out/r2c_ncu_full.log 2>&1
tail -3 gpurun_out/r2c_ncu_full.log
