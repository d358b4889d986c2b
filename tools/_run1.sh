timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"peak_limiter|imdct_ola" -c 24 --csv --log-file gpurun_out/r2_lcout_launches_d.csv python bench.py --workload aac_lc_stereo_output --steps 2 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 > /dev/null 2>&1
grep -c peak_limiter gpurun_out/r2_lcout_launches_d.csv
