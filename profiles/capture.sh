#!/bin/bash
# Run on the GPU box (gpurun): ncu evidence for the round.  Writes into gpurun_out/; copy the CSVs to profiles/.
#   launches_rNN.csv     every launch of one short bench run with its device time (cold-cache, serialised: compare shares)
#   <kernel>_rNN_raw.csv ncu --set full raw metrics of the top kernels (one launch each, 128 x 5 Mbp)
R=${1:-r01}
KERNELS=${2:-"k1_extract bucket_emit_smem bucket_count part_scatter part_hist filter_flags_rows filter_scatter transpose_words"}
mkdir -p gpurun_out
if [ -z "$SKIP_LAUNCH_LIST" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${R}.log 2>&1
fi
if [ -n "$ONE_PROCESS" ]; then
    # round 2: every kernel of one build + align in ONE profiled process (the genomes are synthesised once)
    RX=$(echo $KERNELS | tr ' ' '|')
    PROF_GENOMES=128 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -c 16 -f \
        -o gpurun_out/kernels_${R} python profiles/profile_driver.py > gpurun_out/kernels_${R}.log 2>&1
    ncu -i gpurun_out/kernels_${R}.ncu-rep --page raw --csv > gpurun_out/kernels_${R}_raw.csv 2>/dev/null
else
for kn in $KERNELS; do
    PROF_GENOMES=128 timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kn -c 1 -f \
        -o gpurun_out/${kn}_${R} python profiles/profile_driver.py > gpurun_out/${kn}_${R}.log 2>&1
    ncu -i gpurun_out/${kn}_${R}.ncu-rep --page raw --csv > gpurun_out/${kn}_${R}_raw.csv 2>/dev/null
    ncu -i gpurun_out/${kn}_${R}.ncu-rep --page source --csv > gpurun_out/${kn}_${R}_source.csv 2>/dev/null
    rm -f gpurun_out/${kn}_${R}.ncu-rep
done
fi
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_${R}.csv
ls -la gpurun_out | tail -30
