mkdir -p gpurun_out
rm -f gpurun_out/r2_call16.txt
for v in z33 z28 z20 z12; do echo $v >> gpurun_out/r2_call16.txt; for r in "" morton; do MSP_LIB=$PWD/build/variants/$v.so REORDER=$r python profiles/stress_probe.py 2>/dev/null | tail -1 | cut -c1-90 >> gpurun_out/r2_call16.txt; done; MSP_LIB=$PWD/build/variants/$v.so python profiles/zbuffer_bench.py 2>/dev/null | tail -1 >> gpurun_out/r2_call16.txt; done
echo old >> gpurun_out/r2_call16.txt; MSP_CULL_V9=0 python profiles/zbuffer_bench.py 2>/dev/null | tail -1 >> gpurun_out/r2_call16.txt
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_pytest_k2b.log 2>&1
echo pytest rc=$? >> gpurun_out/r2_call16.txt
cat gpurun_out/r2_call16.txt
