#!/bin/bash
# One gpurun call that produces every raw artefact tools/make_profiles.py turns into profiles/ (1 GPU).
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_profile_run.sh'
# Each command first runs to completion WITHOUT ncu (numbers), then under ncu (launch lists / captures).
set -u
O=gpurun_out
mkdir -p $O
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_r1_final.json 2> $O/bench_r1_final.err || echo "bench failed"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_r1_final_ref.json 2> $O/bench_r1_final_ref.err || echo "reference bench failed"
HG_BENCH_SKIP_TRAIN=1 HG_BENCH_SKIP_CPU=1 timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file $O/launches_r1b.csv \
    python bench.py --steps 2 --warmup 3 > $O/ncu_l.log 2>&1 || echo "launch list failed"
timeout 300 python tools/loss_bench.py --no-cpu --iters 3 --warmup 2 > $O/lb_plain.log 2>&1 \
  && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/loss_launches2.csv \
    python tools/loss_bench.py --no-cpu --iters 3 --warmup 2 > $O/lb_ncu.log 2>&1 || echo "loss list failed"
timeout 300 python tools/train_probe.py --recipe c2 --steps 2 > $O/tp_plain.log 2>&1 \
  && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file $O/train_launches2.csv \
    python tools/train_probe.py --recipe c2 --steps 2 > $O/tp_ncu.log 2>&1 || echo "train list failed"
HG_BENCH_SKIP_TRAIN=1 HG_BENCH_SKIP_CPU=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:blend_ -c 2 -s 12 \
    -f -o $O/prof_blend_r1c python bench.py --steps 2 --warmup 3 > $O/ncu_f.log 2>&1 || echo "full capture failed"
tail -c 300 $O/bench_r1_final.json; echo; tail -c 300 $O/bench_r1_final_ref.json; echo; ls -la $O/*.csv $O/prof_blend_r1c.ncu-rep
