#!/bin/bash
# Command sequence behind the round-2 numbers (run on the GPU box through gpurun; outputs under gpurun_out/).
set -x
python -m pytest tests -m gpu -x -q                                   # 110 passed on one GPU (multi-rank tests need 2)
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json       # N=1, 128^3 cavity, 53.07 M dofs
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
$TR --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_strong.json
$TR --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3 --opt fnp_halo_p2p=0 --no-refresh --no-dist-parity > gpurun_out/bench_n8_strong_nccl.json
$TR --nproc-per-node 2 bench.py --gpus 2 --steps 5 --warmup 3 --scaling weak > gpurun_out/bench_n2_weak.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json
python profiles/run_cfg.py cfg2 471 > gpurun_out/cfg2.json
python profiles/run_cfg.py cfg3 7 50 > gpurun_out/cfg3.json
python profiles/run_cfg.py cfg5newton 24 > gpurun_out/cfg5newton.json
$TR --nproc-per-node 2 tests/dist_worker.py BRM2                       # 2-rank parity incl. the multi-rank value refresh
FNP_PCDR=1 $TR --nproc-per-node 2 tests/dist_worker.py BRM1            # PCDR with the distributed Rp
FNP_RUN_DROPIN_DIST=1 python -m pytest tests/test_dist_gpu.py -m gpu -q -k dropin   # open: re-run after the halo fix
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:spmv_sell_kernel -c 2 \
    -o gpurun_out/prof_a00_n128 python profiles/prof_kernels.py 128 1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:fnp:: -c 12000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --scaling weak --n1 64 --steps 1 --warmup 0 --profile-only --no-clocks
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_warp_sell_kernel and 4 or amg_vcycle and bfs_BRM1"
