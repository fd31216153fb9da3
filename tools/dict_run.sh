mkdir -p gpurun_out
timeout 800 python bench.py > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r1_f.json"))
r=d["roofline"]
print(d["value"], d["e2e"]["value"], r["frac"], r["stream_frac"], r["avg_launch_ms"], r["format"], d["full_solve"]["iterations"], d["clocks"])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 6 -c 2 -o gpurun_out/r01_spmv512_dict_full python bench.py --steps 1 --iters 4 --no-cpu --no-c2 --no-full-solve > gpurun_out/ncu_dict.log 2>&1; echo ncu rc=$?
ncu -i gpurun_out/r01_spmv512_dict_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,sm__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/r01_spmv512_dict_raw.csv 2>&1
cat gpurun_out/r01_spmv512_dict_raw.csv | cut -c1-1500
