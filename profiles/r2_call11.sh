mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_regen.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r2_pytest_c.log 2>&1; echo pytest rc=$? > gpurun_out/r2_call11.txt
for r in "" morton; do REORDER=$r python profiles/stress_probe.py 2>/dev/null | tail -1 >> gpurun_out/r2_call11.txt; done
REORDER= ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_stress.csv python profiles/stress_probe.py > /dev/null 2>&1
REORDER=morton ncu --set full --clock-control none --import-source on -k regex:k_frame_cull -s 6 -c 1 -o gpurun_out/prof_k2_stress python profiles/stress_probe.py > /dev/null 2>&1
cat gpurun_out/r2_call11.txt; tail -4 gpurun_out/r2_pytest_c.log
