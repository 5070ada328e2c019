#!/bin/bash
# round-2 GPU call D (2 GPUs): full GPU tests on one GPU, copy-engine gather vs NCCL, bench at N=2.
O=gpurun_out/r02d; mkdir -p $O
export MWRT_QUIET=1
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
python tools/prof_target2.py c2rt 10000 8 > $O/time_c2rt.log 2>&1; cat $O/time_c2rt.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tools/check_ce_gather.py > $O/check_ce_n2.log 2>&1; echo "check_ce rc=$?"; tail -3 $O/check_ce_n2.log
timeout 600 $TR --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2_ce.json 2> $O/bench_n2_ce.err; echo "bench ce rc=$?"
timeout 300 $TR --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 3 --gather nccl --no-extras > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err; echo "bench nccl rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"
nvidia-smi topo -m > $O/topo.txt 2>&1
tail -3 $O/*.err
ls -la $O
