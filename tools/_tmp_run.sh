HG_BENCH_SKIP_TRAIN=1 HG_BENCH_SKIP_CPU=1 python bench.py --steps 20 --warmup 5 > gpurun_out/s3_e.json 2> gpurun_out/s3_e.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/s3_e.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['stage_ms'])"
