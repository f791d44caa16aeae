#!/bin/bash
# A/B sweep of the binning knobs (temporary tuning aid): prints the stage times of bench.py's resident leg per variant.
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" HG_BENCH_SKIP_TRAIN=1 HG_BENCH_SKIP_CPU=1 python bench.py --steps 10 --warmup 3 > $O/sweep_$name.json 2> $O/sweep_$name.err
  python - "$name" $O/sweep_$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    s = d["roofline"]["stage_ms"]
    print("%-28s value %.1f  pre %.4f scan %.4f bin %.4f blend_fwd %.4f" % (sys.argv[1], d["value"], s["preprocess_fwd"], s["scan"], s["binning"], s["blend_fwd"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run base HG_CTR_STRIDE=2
run per4 HG_CTR_STRIDE=2 HG_SCATTER_PER=4
run per1 HG_CTR_STRIDE=2 HG_SCATTER_PER=1
