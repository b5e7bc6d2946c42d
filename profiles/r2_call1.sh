set -x
mkdir -p gpurun_out
./build/ubench_f32x2 > gpurun_out/ubench_f32x2.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | head -30 > gpurun_out/lscpu.txt
numactl -H >> gpurun_out/lscpu.txt 2>&1
MSP_LIB=$PWD/build/variants/r1all.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_r1all.log 2>&1
echo pytest_r1all rc=$? >> gpurun_out/r2_call1.txt
for v in base defer r1ds r1all; do MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call1.txt; done
for v in base r1all; do MSP_LIB=$PWD/build/variants/$v.so python bench.py --scenes 12 --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$v', 'pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'])" >> gpurun_out/r2_call1.txt; done
cat gpurun_out/r2_call1.txt gpurun_out/ubench_f32x2.txt; tail -3 gpurun_out/r2_pytest_r1all.log
