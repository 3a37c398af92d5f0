#!/bin/bash
# `ncu --set full` over one launch of every library kernel (tools/ncu_kernels.py) + a CSV summary of the metrics
# profiles/ quotes.  Run on the GPU box: ./tools/ncu_all_kernels.sh <tag>
TAG=${1:-r01_all_kernels}
KERNELS=${2:-'msda_|point_sample|add_layernorm|colsum|qv_cast'}
mkdir -p gpurun_out
set -e
timeout 600 python tools/ncu_kernels.py > gpurun_out/${TAG}_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log
timeout 1200 ncu --set full --clock-control none -k regex:"$KERNELS" -f -o gpurun_out/$TAG \
  python tools/ncu_kernels.py > gpurun_out/${TAG}_ncu.log 2>&1 || { tail -5 gpurun_out/${TAG}_ncu.log; exit 1; }
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
M=$M,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__inst_executed.sum
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv --metrics $M > gpurun_out/${TAG}.csv 2> gpurun_out/${TAG}_csv.err || true
wc -l gpurun_out/${TAG}.csv; ls -la gpurun_out/$TAG.ncu-rep
