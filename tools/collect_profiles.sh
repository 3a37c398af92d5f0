#!/bin/bash
# Copy the evidence of the last full gpurun (tests, bench, reference arm, ncu launch list, ncu --set full of every kernel)
# from gpurun_out/ into profiles/ under the round's tag and regenerate the derived summaries. No GPU needed.
#   tools/collect_profiles.sh [tag]
set -e
cd "$(dirname "$0")/.."
T=${1:-r02}
cp gpurun_out/${T}_bench_n1.json profiles/${T}_bench_n1.json
cp gpurun_out/${T}_bench_reference_arm.json profiles/${T}_bench_reference_arm.json
cp gpurun_out/${T}_launches.csv profiles/${T}_launches.csv
cp gpurun_out/${T}_all_kernels.csv profiles/${T}_ncu_all_kernels.csv
cp gpurun_out/${T}_gputest.log profiles/${T}_gputest.log
python tools/ncu_segments.py gpurun_out/${T}_all_kernels.ncu-rep --kernel msda_bwd_mma > profiles/${T}_ncu_bwd_mma_segments.txt
python tools/ncu_segments.py gpurun_out/${T}_all_kernels.ncu-rep --kernel msda_fwd_pair > profiles/${T}_ncu_fwd_pair_segments.txt
for k in msda_fwd_pair msda_bwd_mma; do
  ncu -i gpurun_out/${T}_all_kernels.ncu-rep --page raw --csv -k regex:$k -c 1 2>/dev/null | python -c "
import csv,sys,re
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
pat=r'^(l1tex__data_pipe_lsu_wavefronts(_mem_shared|_mem_lgds)?(_op_(ld|st|atom))?\.(sum|avg)(\.pct_of_peak_sustained_elapsed)?|l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st|atom))?\.sum|l1tex__t_(requests|sectors|set_accesses)[a-z_]*\.sum|l1tex__lsu_writeback_active(_mem_lgds)?\.(avg|sum)\.pct_of_peak_sustained_elapsed|smsp__inst_executed_pipe_(lsu|alu|fma|fmaheavy|tensor[a-z_]*|xu|uniform)\.sum|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|launch__(registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_[a-z_]+)|smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio)$'
print('# kernel:', v[h.index('Kernel Name')][:110])
for a,b,c in zip(h,u,v):
    if re.match(pat,a): print(f'{a} = {c} {b}')
"
done > profiles/${T}_ncu_counters_fwd_bwd.txt
python tools/update_traffic.py gpurun_out/${T}_all_kernels.csv ${T}_all_kernels
tools/dump_sass.sh $T > /dev/null
python - <<PY
import json
d = json.loads(open("profiles/${T}_bench_n1.json").read().strip().splitlines()[-1])
print("step %.4f ms, value %.1f, roofline %s" % (d["ms_per_step"], d["value"], {k: d["roofline"][k] for k in ("kernel", "frac", "kernel_ms", "traffic")}))
print("kernels", d["roofline_step"]["kernels_ms"])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["copy_ceiling"]["ms_per_step"], "train", d["train"]["images_per_s"], d["train"]["stock_hf"]["images_per_s"])
print("gpu_reference", d["gpu_reference"]["fp32_ms_per_step"], d["gpu_reference"]["autocast_bf16_ms_per_step"], "cpu", d["cpu_baseline"]["value"])
print({k: round(v["ms_per_step"], 3) for k, v in d["other_workloads"].items()})
PY
