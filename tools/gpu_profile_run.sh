#!/bin/bash
# One gpurun call that produces every raw artefact tools/make_profiles.py turns into profiles/ (1 GPU).
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_profile_run.sh r02'
# Each command first runs to completion WITHOUT ncu (numbers), then under ncu (launch lists / captures).
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
FULL="--set full --clock-control none --import-source on"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || echo "bench failed"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err || echo "reference bench failed"
timeout 300 python tools/hier_probe.py --iters 5 > $O/${TAG}_hier_probe.json 2> $O/${TAG}_hier_probe.err || echo "hier probe failed"
HG_BENCH_SKIP_TRAIN=1 HG_BENCH_SKIP_CPU=1 timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 > $O/${TAG}_ncu_l.log 2>&1 || echo "launch list failed"
timeout 300 python tools/loss_bench.py --no-cpu --iters 3 --warmup 2 > $O/${TAG}_lb_plain.log 2>&1 \
  && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_loss_launches.csv \
    python tools/loss_bench.py --no-cpu --iters 3 --warmup 2 > $O/${TAG}_lb_ncu.log 2>&1 || echo "loss list failed"
timeout 300 python tools/train_leg_probe.py --steps 3 > $O/${TAG}_tp_plain.log 2>&1 \
  && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_train_launches.csv \
    python tools/train_leg_probe.py --steps 2 > $O/${TAG}_tp_ncu.log 2>&1 || echo "train list failed"
# full-section captures: one launch of each kernel of interest from the second resident step
HG_BENCH_SKIP_TRAIN=1 HG_BENCH_SKIP_CPU=1 timeout 900 ncu $FULL -k regex:"blend_|tile_sort|scatter_inst|tile_scan|preprocess_" -s 8 -c 8 \
    -f -o $O/${TAG}_prof_raster python bench.py --steps 2 --warmup 3 > $O/${TAG}_ncu_f.log 2>&1 || echo "raster capture failed"
timeout 900 ncu $FULL -k regex:"fft_cols|fft_rows_jobs|ssim_fwd|ssim_bwd|pyramid_kernel" -s 10 -c 5 \
    -f -o $O/${TAG}_prof_loss python tools/loss_bench.py --no-cpu --iters 2 --warmup 1 > $O/${TAG}_ncu_fl.log 2>&1 || echo "loss capture failed"
timeout 900 ncu $FULL -k regex:"blend_|preprocess_|tile_sort|scatter" -s 60 -c 7 \
    -f -o $O/${TAG}_prof_uav python tools/train_leg_probe.py --steps 1 > $O/${TAG}_ncu_fu.log 2>&1 || echo "uav capture failed"
tail -c 300 $O/${TAG}_bench.json; echo; tail -c 300 $O/${TAG}_bench_ref.json; echo; ls -la $O/${TAG}_*
