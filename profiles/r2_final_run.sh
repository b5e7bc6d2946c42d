#!/bin/bash
# Round-2 GPU pass on the SHIPPED library: whole GPU suite, smoke, bench (both arms), one `ncu --set full` capture of K1 from
# the bench command, the launch list of the same command, compute-sanitizer on smoke + the K1 parity tests.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_final_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a gpurun_out/r2_final_pytest.log
python bench.py 2>gpurun_out/r2_final_bench.err | tail -1 > gpurun_out/r2_final_bench.json
cut -c1-300 gpurun_out/r2_final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/r2_final_bench_ref.err | tail -1 > gpurun_out/r2_final_bench_ref.json
cut -c1-200 gpurun_out/r2_final_bench_ref.json
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_frame_cull9" -s 3 -c 2 -f \
    -o gpurun_out/prof_k1_r2_final python bench.py --scenes 3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/prof_k1_r2_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final.csv \
    python bench.py --scenes 3 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/launches_r2_final.log 2>&1
compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r2_sanitizer_memcheck_smoke.log 2>&1
compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r2_sanitizer_racecheck_smoke.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r2_sanitizer_memcheck_pytest.log 2>&1
for f in gpurun_out/r2_sanitizer_*.log; do tail -n 3 "$f"; done
TB_STREAM_ONLY=1 python profiles/table_bench.py > gpurun_out/r2_tables_stream.json 2> gpurun_out/r2_tables_stream.err; cat gpurun_out/r2_tables_stream.json
