mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_call29.txt
VC_FRAMES=2000 python profiles/variant_check.py 2>&1 | tail -1 >> gpurun_out/r2_call29.txt
python bench.py --scenes 12 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r2_call29.err | tail -1 > gpurun_out/r2_call29_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2_call29_bench.json')); print('pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms']); 
for k in ('configs1_12_objects','configs1_150_objects','configs1_shuffled_as_given','configs1_shuffled_morton'): print(k, d['extra'][k]['k1_ms_per_launch'])" >> gpurun_out/r2_call29.txt
cat gpurun_out/r2_call29.txt
