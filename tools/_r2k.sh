mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slot_segmented or hub_shaped" > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
for sa in 1 2; do for sb in 1 2; do echo "stagesA=$sa stagesBC=$sb"; TIP_SEG3_CHUNK=4 TIP_SEG3_STAGES_A=$sa TIP_SEG3_STAGES_BC=$sb ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_l.csv python tools/profile_em.py 10 800000 3 32 > /dev/null 2>&1
grep -i "seg3_pass" gpurun_out/r2k_l.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -2; done; done
for sa in 1 2; do for sb in 1 2; do for tune in 0 1; do echo "stagesA=$sa stagesBC=$sb tune=$tune"; TIP_SEG3_CHUNK=4 TIP_SEG3_STAGES_A=$sa TIP_SEG3_STAGES_BC=$sb TIP_SEG3_TUNE=$tune timeout 120 python tools/hub_probe.py --flags 32 --steps 10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['shape'],round(d['ms_per_iteration'],4))"
done; done; done
