mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_b.log 2>&1; echo pytest rc=$? > gpurun_out/r2_call10.txt
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo bench rc=$? >> gpurun_out/r2_call10.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo ref rc=$? >> gpurun_out/r2_call10.txt
cat gpurun_out/r2_call10.txt; tail -5 gpurun_out/r2_pytest_b.log; tail -5 gpurun_out/r2_bench_a.err
