mkdir -p gpurun_out
rm -f gpurun_out/r2_call22.txt
for v in bt1u1 bt1f64 bt1s0; do echo "$v $(VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>&1 | tail -1)" >> gpurun_out/r2_call22.txt; done
for v in bt1u1 bt1f64 bt1s0; do MSP_LIB=$PWD/build/variants/$v.so python bench.py --scenes 12 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$v', 'pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'])" >> gpurun_out/r2_call22.txt; done
VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/bt1u1.so ncu --set full --clock-control none --import-source on -k regex:k_frame_cull9 -s 2 -c 1 -f -o gpurun_out/prof_bt1u1 python profiles/variant_check.py > /dev/null 2>&1
cat gpurun_out/r2_call22.txt
