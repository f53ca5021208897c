#!/bin/bash
# One GPU visit: parity tests, the bench line, an ncu launch list of one step, optional wait-hint sweep.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01f [sweep]'
tag=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
if [ "$2" = "sweep" ]; then
  for ns in 0 64 256 1000; do
    B200_WAIT_NS=$ns timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_wait${ns}.json 2> gpurun_out/${tag}_wait${ns}.err
    python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${tag}_wait${ns}.json').read().strip().splitlines()[-1])
    print('wait_ns=${ns}', 'ms/step', round(d['ms_per_step'], 3), 'gemm ms', round(d['roofline']['kernel_ms_per_step'], 3), 'gallery q/s', round(d['gallery']['value']))
except Exception as e:
    print('wait_ns=${ns} failed', e)
PY
  done
fi
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gallery > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
