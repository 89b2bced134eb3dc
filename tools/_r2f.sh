mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slot_segmented or order_rows or hub_shaped" > gpurun_out/r2f_pytest.log 2>&1; tail -5 gpurun_out/r2f_pytest.log
for st in 2 3; do for tune in 0 1; do for ch in 2 3; do
  echo "stages=$st tune=$tune chunk=$ch"; TIP_SEG3_STAGES=$st TIP_SEG3_TUNE=$tune TIP_SEG3_CHUNK=$ch timeout 120 python tools/hub_probe.py --flags 32 --steps 10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['shape'],round(d['ms_per_iteration'],4))"
done; done; done > gpurun_out/r2f_tune.txt 2>&1
cat gpurun_out/r2f_tune.txt
TIP_SEG3_TUNE=1 python tools/profile_em.py 10 800000 4 32 > gpurun_out/r2f_plain.log 2>&1 && TIP_SEG3_TUNE=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv python tools/profile_em.py 10 800000 4 32 > gpurun_out/r2f_ncu_list.log 2>&1
cat gpurun_out/r2f_plain.log; grep -i "seg3_" gpurun_out/r2f_launches.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -4
