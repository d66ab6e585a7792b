# GPU-box command sequence behind the round-1 numbers (run through gpurun from the repo root)
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python profiles/exp_sell_gather.py 64 2>&1 | tail -6
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -1 gpurun_out/final_bench.json | cut -c1-1500
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:spmv_sell_kernel -c 2 -f -o gpurun_out/prof_spmv_sell_kron_v2 python profiles/prof_kernels.py 64 1 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:fnp:: -c 9000 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --n1 64 --steps 1 --warmup 0 --profile-only --no-clocks > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-300
