#!/bin/bash
# round-2 GPU call W (N GPUs = $1): the default bench line with its extras (configs 3/5/4 at full size split over
# the ranks) on the final kernel-1 build (far-line grid, per-block scratch cached per plan)
N=${1:-2}
O=gpurun_out/r02w_n$N; mkdir -p $O
export MWRT_QUIET=1
timeout 400 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29611 \
  bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -n 5 $O/bench.err
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("bench N=$N value %.4e ms/step %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
for c in d.get("extra", {}).get("configs", []):
    print(c["config"][:28], {k: c[k] for k in ("n_gpus", "ms_per_pass", "ms_kernels_only", "gather_identical_to_nccl", "finite") if k in c})
PY
