#!/bin/bash
# Round-end evidence on one B200 (gpurun): smoke, the bench lines of every workload, the reference arm.  Outputs in gpurun_out/.
R=${1:-r02}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${R}.log 2>&1; tail -1 gpurun_out/smoke_${R}.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; tail -c 300 gpurun_out/bench_${R}.json; echo
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_${R}.json 2> gpurun_out/bench_ref_${R}.err; tail -c 400 gpurun_out/bench_ref_${R}.json; echo
timeout 600 python bench.py --workload distance --steps 5 --warmup 3 > gpurun_out/bench_distance_${R}.json 2> gpurun_out/bench_distance_${R}.err; tail -c 300 gpurun_out/bench_distance_${R}.json; echo
timeout 600 python bench.py --workload viral --steps 20 --warmup 3 > gpurun_out/bench_viral_${R}.json 2> gpurun_out/bench_viral_${R}.err; tail -c 300 gpurun_out/bench_viral_${R}.json; echo
timeout 900 python bench.py --workload cli > gpurun_out/bench_cli_${R}.json 2> gpurun_out/bench_cli_${R}.err; tail -c 300 gpurun_out/bench_cli_${R}.json; echo
timeout 900 python bench.py --workload cli --impl reference > gpurun_out/bench_cli_ref_${R}.json 2> gpurun_out/bench_cli_ref_${R}.err; tail -c 300 gpurun_out/bench_cli_ref_${R}.json; echo
