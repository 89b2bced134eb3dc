mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slot_segmented or hub_shaped" > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
for tune in 0 6; do echo "tune=$tune"; TIP_SEG3_TUNE=$tune ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_l.csv python tools/profile_em.py 10 800000 3 32 > /dev/null 2>&1
grep -i "seg3_pass" gpurun_out/r2j_l.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -2; done
for ch in 4 8 16; do for tune in 0 1; do echo "chunk=$ch tune=$tune"; TIP_SEG3_CHUNK=$ch TIP_SEG3_TUNE=$tune timeout 120 python tools/hub_probe.py --flags 32 --steps 10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['shape'],round(d['ms_per_iteration'],4))"
done; done
