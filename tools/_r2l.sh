mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; tail -8 gpurun_out/r2l_pytest.log
TIP_SEG3_CHUNK=4 timeout 200 python tools/hub_probe.py --flags 8,32,96 --steps 20 > gpurun_out/r2l_hub_probe.jsonl 2>gpurun_out/r2l_hub.err; cat gpurun_out/r2l_hub_probe.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['shape'],d['flags'],round(d['ms_per_iteration'],4),d['rel_err_vs_first_flags'])"
for skip in 2 16; do
  echo "skip=$skip"; TIP_SEG3_CHUNK=4 TIP_SEG3_SKIP=$skip timeout 120 python tools/hub_probe.py --flags 32 --steps 10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['shape'],round(d['ms_per_iteration'],4))"
done
