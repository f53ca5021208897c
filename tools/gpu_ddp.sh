#!/bin/bash
# Multi-GPU visit: the torchrun parity worker (N=2 only) and the bench line with both gradient-exchange transports.
#   gpurun --gpus N --timeout 600 -- 'bash tools/gpu_ddp.sh r03a N [test]'
tag=${1:-run}; n=${2:-2}
mkdir -p gpurun_out
if [ "$3" = "test" ]; then
  timeout 400 python -m pytest tests/test_multigpu_gpu.py -x -q -s > gpurun_out/${tag}_mgpu_pytest.log 2>&1; echo "pytest rc=$?"
  grep -E "world|ResNet|sharded|MGPU|passed|failed|Error|error" gpurun_out/${tag}_mgpu_pytest.log | tail -12
fi
for mode in p2p nccl; do
  B200_DDP=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 10 --warmup 3 --no-gallery --no-cpu-baseline > gpurun_out/${tag}_n${n}_${mode}.json 2> gpurun_out/${tag}_n${n}_${mode}.err
  echo "bench $mode rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${tag}_n${n}_${mode}.json').read().strip().splitlines()[-1])
    print('${mode}', 'img/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), d['config'].get('grad_exchange'), d['clocks'])
except Exception as e:
    print('${mode} failed', e)
PY
  tail -3 gpurun_out/${tag}_n${n}_${mode}.err
done
