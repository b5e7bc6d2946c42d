mkdir -p gpurun_out
rm -f gpurun_out/r2_call9.txt
for v in v13c v13cf32; do MSP_LIB=$PWD/build/variants/$v.so python bench.py --scenes 12 --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$v', 'pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'])" >> gpurun_out/r2_call9.txt; done
MSP_LIB=$PWD/build/variants/v13cf32.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v13cf32.log 2>&1
echo pytest_v13cf32 rc=$? >> gpurun_out/r2_call9.txt
cat gpurun_out/r2_call9.txt; tail -3 gpurun_out/r2_pytest_v13cf32.log
