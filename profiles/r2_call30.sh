mkdir -p gpurun_out
rm -f gpurun_out/r2_call30.txt
V="$(cd build/variants && ls *.so | sed 's/\.so$//' | tr '\n' ' ')"
for v in $V; do echo "$v $(VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>&1 | tail -1)" >> gpurun_out/r2_call30.txt; done
for v in $V; do MSP_LIB=$PWD/build/variants/$v.so python bench.py --scenes 12 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$v', 'pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'])" >> gpurun_out/r2_call30.txt; done
cat gpurun_out/r2_call30.txt
