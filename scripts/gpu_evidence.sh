#!/bin/bash
# round-2 evidence: bench lines (both arms), ncu launch list, ncu --set full of the top kernels, full GPU test run
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 600 --csv --log-file gpurun_out/launches_tf32x3.csv python bench.py --steps 2 --warmup 3 --skip-retrieval --skip-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel|conv_wgrad_pk_kernel|conv_wgrad_tc_kernel|conv1x1_wgrad_tc_kernel|pairwise_tc_kernel|pairwise_topk_finish|row_argsort_kernel" -c 24 -o gpurun_out/r2_kernels -f python scripts/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1; echo "ncu full exit $?"
timeout 400 python bench.py --workload config4 --steps 20 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/r2_bench_n1_config4.json 2> gpurun_out/r2_bench_n1_config4.err; echo "config4 exit $?"
timeout 400 python bench.py --workload config3 --steps 20 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/r2_bench_n1_config3.json 2> gpurun_out/r2_bench_n1_config3.err; echo "config3 exit $?"
timeout 300 python scripts/bench_conv.py 1x1 > gpurun_out/bench_conv_1x1.txt 2>&1
timeout 1500 python -m pytest tests/ -q -m gpu --timeout=900 2>&1 | tail -n 25 | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -c "
import json
for w in ('config3', 'config4'):
    d=json.load(open('gpurun_out/r2_bench_n1_%s.json' % w)); print(w, d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['retrieval']['value'], d['retrieval']['roofline']['frac'], d['cpu_baseline'], d['clocks'])
r=json.load(open('gpurun_out/r2_bench_ref.json')); print(r['value'], r['config'])"
ls -la gpurun_out/r2_kernels.ncu-rep
