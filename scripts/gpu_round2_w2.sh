#!/bin/bash
# round-2 session-2: 2-GPU bench line (weak-scaling pipeline + the strong-scaling group-by / sort exchange legs) and the NCCL tests
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/w_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/w_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 \
   > gpurun_out/w_bench_n2.json 2> gpurun_out/w_bench_n2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/w_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/w_bench_n2.json') if l.startswith('{')][-1])
    print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}), json.dumps(d['e2e']), json.dumps(d['multi_gpu'], indent=1))
except Exception as e:
    print('no bench line', e)
PY
