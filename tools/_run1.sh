M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -s 33 -c 22 --csv --log-file gpurun_out/r2_chain_launches_c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:peak_limiter -s 9 -c 6 --csv --log-file gpurun_out/r2_lcout_launches_c.csv python bench.py --workload aac_lc_stereo_output --steps 2 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:sideinfo -s 3 -c 2 --csv --log-file gpurun_out/r2_sideinfo_launches_c.csv python bench.py --workload sbr_sideinfo --steps 2 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 > /dev/null 2>&1
grep -c '^"' gpurun_out/r2_chain_launches_c.csv gpurun_out/r2_lcout_launches_c.csv gpurun_out/r2_sideinfo_launches_c.csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
