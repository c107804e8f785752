set -x
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "trunk" 2>&1 | tail -5
python -m pytest tests/test_gpu_networks.py tests/test_gpu_parity_at_size.py -x -q -m gpu 2>&1 | tail -5
python tools/profile_layers.py 2>&1 | sed -n 1,40p > gpurun_out/r2o_layers.txt
OU_TRUNK_OUT_TAIL=0 python tools/profile_layers.py 2>&1 | sed -n 1,40p > gpurun_out/r2o_layers_notail.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-other-configs --no-kernel-events > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
OU_TRUNK_OUT_TAIL=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-other-configs --no-kernel-events > gpurun_out/r2o_bench_notail.json 2> gpurun_out/r2o_bench_notail.err
python -c "
import json
for f in ('r2o_bench','r2o_bench_notail'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
"
