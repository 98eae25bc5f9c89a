#!/bin/bash
# config 3 (fused rate + sum by + tiled all-reduce) at N GPUs: "<reserve SMs>:<headstart us>:<NCCL max channels or ->" per run
cd /root/repo
N=${1:-2}; shift
for spec in "$@"; do
  IFS=: read r hs ch <<< "$spec"
  export B2P_COMM_RESERVE_SMS=$r B2P_COMM_HEADSTART_US=$hs
  if [ "$ch" != "-" ]; then export NCCL_MAX_NCHANNELS=$ch NCCL_MIN_NCHANNELS=$ch; else unset NCCL_MAX_NCHANNELS NCCL_MIN_NCHANNELS; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 6 --warmup 3 --workload sumby --no-cpu-baseline --e2e-series 0 --jitter-variant-ms 0 2>gpurun_out/multi_${N}_$spec.err | tail -1 > gpurun_out/multi_${N}_$spec.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/multi_${N}_$spec.json").read().strip().splitlines()[-1])
c=d["configs"]["3"]; k=c["collective"]
print("N=$N spec=$spec ms=%.2f  without_coll=%.2f exposed=%.2f last_ar=%.2f" % (c["ms_per_step"], k["ms_per_step_without_collective"], k["ms_exposed"], k["last_tile_allreduce_kernel_ms"]))
PY
done
