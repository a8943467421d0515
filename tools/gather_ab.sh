#!/bin/bash
# A/B of the three ways rank 0 gets every rank's outputs (bench.py --gather push|peer|nccl), N = $1 GPUs
N=${1:-2}
for mode in ${MODES:-push peer nccl}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N \
    bench.py --gpus $N --steps 10 --warmup 3 --no-configs --no-cpu-baseline --gather $mode \
    > gpurun_out/r02_gather_${mode}_n$N.json 2> gpurun_out/r02_gather_${mode}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_gather_${mode}_n$N.json").read().strip().splitlines()[-1])
    print("${mode} N=$N: %.0f clips/s, %.2f ms/step, e2e %.0f (f32 %.0f)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["f32_input_value"]))
except Exception as e:
    print("${mode}: failed", e)
PY
done
