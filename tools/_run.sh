python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py 2>gpurun_out/final.err | tail -1 > gpurun_out/bench_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['clocks'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['roofline']['share_of_step'], d['gpu_launches'], d['kernels'])"
