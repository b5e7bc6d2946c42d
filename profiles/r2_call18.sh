mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_call18.txt
python bench.py --scenes 12 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'], d['extra']['regen'], d['clocks'])" >> gpurun_out/r2_call18.txt
cat gpurun_out/r2_call18.txt
