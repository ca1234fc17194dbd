#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/k_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_exchange_kernels.py tests/test_gpu_distributed.py "tests/test_gpu_parity.py::test_sort_payload_equals_sort_indices_then_take" -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
tail -12 gpurun_out/k_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 \
   > gpurun_out/k_bench_n2.json 2> gpurun_out/k_bench_n2.err; echo "bench rc=$?"
tail -c 800 gpurun_out/k_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/k_bench_n2.json') if l.startswith('{')][-1])
    print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}), json.dumps(d['e2e']), json.dumps(d['multi_gpu'], indent=1))
except Exception as e:
    print('no bench line', e)
PY
