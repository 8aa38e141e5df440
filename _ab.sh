#!/bin/bash
python bench.py > gpurun_out/bench_table_n1_c.json 2> gpurun_out/bench_c.err; tail -c 600 gpurun_out/bench_table_n1_c.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/launches_c.csv | head -12
timeout 600 compute-sanitizer --tool memcheck python profiles/sanitizer_target.py 2>&1 | tail -12
