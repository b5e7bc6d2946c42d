mkdir -p gpurun_out
rm -f gpurun_out/r2_call19.txt
for v in mb4 mb5; do VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call19.txt; done
for v in mb4 mb5; do MSP_LIB=$PWD/build/variants/$v.so python bench.py --scenes 12 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$v', 'pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'])" >> gpurun_out/r2_call19.txt; done
VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/mb5.so ncu --set full --clock-control none --import-source on -k regex:k_frame_cull9 -s 2 -c 1 -o gpurun_out/prof_mb5 python profiles/variant_check.py > /dev/null 2>&1
cat gpurun_out/r2_call19.txt
