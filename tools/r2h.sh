mkdir -p gpurun_out/r2h
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"p2p_|scan_|rs_|memset" -c 120 --csv --log-file gpurun_out/r2h/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2h/ncu.log 2>&1
python profiles/summarize.py launches gpurun_out/r2h/launches.csv | head -30
