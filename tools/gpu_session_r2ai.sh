#!/bin/bash
# Round 2, session AI (1 GPU): full-set ncu capture of the three tcgen05 projection kernels of the final build (tensor-pipe
# utilisation and DRAM throughput against peak), summarised on the box.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CMD="python bench.py --no-graph --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 3 --warmup 3"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:(project_tc_kernel|^wgrad_tc_kernel|sgc::tc::wgrad_tc_kernel)" -s 27 -c 9 -o $O/r2ai_project -f $CMD > $O/r2ai_ncu_project.log 2>&1
python tools/summarize_ncu.py full $O/r2ai_project.ncu-rep $O/r2ai_project_full_summary.md $O/r2ai_project_traffic.json > /dev/null 2>&1
ls -la $O/r2ai_project.ncu-rep | awk '{print $5}'
while [ "$(du -sm $O | cut -f1)" -gt 56 ]; do f=$(ls -S $O/*.ncu-rep 2>/dev/null | head -1); [ -z "$f" ] && break; rm -f "$f"; done
