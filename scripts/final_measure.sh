#!/bin/bash
# Round-end evidence run on one B200: parity suite, the bench lines, the ncu launch list of the bench command and one
# --set full capture of the headline kernel.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3) > gpurun_out/r2_pytest_gpu.log 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_bench_under_ncu.json 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_group -s 120 -c 1 -o gpurun_out/r2_prof_group python bench.py --steps 20 --warmup 3 --no-cpu --no-configs > /dev/null 2>&1
timeout 300 python scripts/host_ceiling.py --pipeline > gpurun_out/r2_pipeline_1gpu.jsonl 2>/dev/null
timeout 300 python scripts/host_ceiling.py --stream > gpurun_out/r2_stream_1gpu.json 2>/dev/null
tail -n 3 gpurun_out/r2_pytest_gpu.log
head -c 700 gpurun_out/r2_bench_final.json
