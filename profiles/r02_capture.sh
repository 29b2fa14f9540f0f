#!/bin/bash
# profiles/r02_capture.sh — the ncu passes behind profiles/r02_* (run on a B200 through gpurun from the repo root; every ncu run is
# preceded by the same command line without ncu, B200_PROFILING.md). Outputs land in gpurun_out/; profiles/summarize.py and
# profiles/r02_summarize.py turn them into the committed text / JSON summaries.
set -u
C3="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass --no-graph --no-parity --no-strong --no-dropin --no-secondary"
C2="python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass --no-graph --no-parity --no-strong --no-dropin --no-secondary"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct"
# one C3 step = k_set_batch, k_intersect_packet (bounce 0: generates the camera rays and walks them), 15 x k_intersect_closest, 16 x k_shade,
# 15 x k_intersect_shadow, k_accumulate, k_resolve = 50 launches; 3 warm-up steps precede it
timeout 300 $C3 > gpurun_out/r02_c3_plain.log 2>&1 || exit 1
# (1) every launch of one C3 step with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_set_batch|k_intersect_packet|k_intersect_closest|k_intersect_shadow|k_shade|k_accumulate|k_resolve" -s 150 -c 50 --csv --log-file gpurun_out/r02_c3_launches.csv $C3 > gpurun_out/r02_c3_ncu1.log 2>&1
# (2) DRAM bytes, issue / L1 utilisation, active lanes of EVERY traversal and shading launch of the same step (the roofline's population)
timeout 900 ncu --metrics $M --clock-control none -k regex:"k_intersect_packet|k_intersect_closest|k_intersect_shadow|k_shade|k_accumulate|k_resolve" -s 147 -c 49 --csv --log-file gpurun_out/r02_c3_metrics.csv $C3 > gpurun_out/r02_c3_ncu2.log 2>&1
# (3) full sections + source for a mid-path bounce (bounce 4: closest, shade, shadow) of the same step, and for the packet kernel (bounce 0)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect_closest|k_intersect_shadow|k_shade" -s 149 -c 3 -o gpurun_out/r02_c3_bounce4 -f $C3 > gpurun_out/r02_c3_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect_packet" -s 3 -c 1 -o gpurun_out/r02_c3_packet -f $C3 > gpurun_out/r02_c3_ncu4.log 2>&1
# (4) C2: one step = k_set_batch, 16 k_bounce_brute + 13 k_brute_finish launches (12 of them return at once), k_accumulate, k_resolve = 32 launches
timeout 300 $C2 > gpurun_out/r02_c2_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics $M --clock-control none -s 96 -c 32 --csv --log-file gpurun_out/r02_c2_metrics.csv $C2 > gpurun_out/r02_c2_ncu.log 2>&1
echo captured
