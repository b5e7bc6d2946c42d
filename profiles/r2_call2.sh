set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo smoke rc=$? > gpurun_out/r2_call2.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v9.log 2>&1
echo pytest_v9 rc=$? >> gpurun_out/r2_call2.txt
MSP_CULL_V9=0 python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call2.txt
for v in v9 v9coef v9f32 v9ns v9f96; do MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call2.txt; done
for v in v9 v9coef v9f32 v9ns v9f96; do MSP_LIB=$PWD/build/variants/$v.so python bench.py --scenes 12 --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$v', 'pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'])" >> gpurun_out/r2_call2.txt; done
cat gpurun_out/r2_call2.txt; tail -15 gpurun_out/r2_pytest_v9.log; tail -5 gpurun_out/r2_smoke.log
