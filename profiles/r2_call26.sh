mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_call26.txt
python bench.py --no-cpu 2>gpurun_out/r2_call26.err | tail -1 > gpurun_out/r2_call26_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2_call26_bench.json')); print('pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks']); print(json.dumps(d['extra'])[:1500])" >> gpurun_out/r2_call26.txt
cat gpurun_out/r2_call26.txt
