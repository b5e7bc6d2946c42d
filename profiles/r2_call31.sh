mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_call31.txt
python bench.py --scenes 12 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r2_call31.err | tail -1 > gpurun_out/r2_call31_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2_call31_bench.json')); print('pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms']); 
for k,v in d['extra'].items(): print(k, {a:b for a,b in v.items() if a!='note'})" >> gpurun_out/r2_call31.txt
cat gpurun_out/r2_call31.txt
