#!/bin/bash
# Strong-scaling series on ONE 8-GPU box (run under `gpurun --gpus 8`): the driver's own launch line for N = 1 (twice, for the
# run-to-run spread on this node), 2, 4, 8.  Lines land in gpurun_out/scale_<N>.json.
mkdir -p gpurun_out
for n in 1 1b 2 4 8; do
  N=${n%b}
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  tail -n 1 gpurun_out/scale_$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=%d device %.3f ms  e2e %.3f ms  frac %.3f  phases %s  host %s clocks %s' % (d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], {k: round(v,3) for k,v in d['phases_ms'].items()}, {k: round(v,3) for k,v in d['e2e'].get('host_ms_rank0',{}).items()}, d['clocks']))"
done
nvidia-smi --query-gpu=index,clocks.mem,clocks.max.mem,clocks.sm,clocks.max.sm,power.limit --format=csv
