mkdir -p gpurun_out
rm -f gpurun_out/r2_call14.txt
for v in v16old v16 v16s v16h0; do VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call14.txt; done
for v in v16old v16 v16s v16h0; do MSP_LIB=$PWD/build/variants/$v.so python bench.py --scenes 12 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$v', 'pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'])" >> gpurun_out/r2_call14.txt; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v16.log 2>&1
echo pytest_v16 rc=$? >> gpurun_out/r2_call14.txt
cat gpurun_out/r2_call14.txt; tail -3 gpurun_out/r2_pytest_v16.log
