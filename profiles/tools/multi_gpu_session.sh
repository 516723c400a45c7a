#!/bin/bash
# one multi-GPU gpurun call: feed ceiling, pool parity on every device count the box has, bench.py under torchrun at N = 2 / 4 / 8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_r02.txt 2>&1
timeout 300 ./profiles/tools/feed_probe 0.4 all > gpurun_out/feed_ceiling.json 2> gpurun_out/feed_probe.err; cat gpurun_out/feed_ceiling.json
timeout 600 python -m pytest tests/test_gpu_pool.py -x -q -m gpu 2>&1 | tail -3
for n in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 2 --warmup 3 --no-large --no-latency --no-stereo > gpurun_out/bench_r02_n$n.json 2> gpurun_out/bench_r02_n$n.err
  echo "N=$n rc=$?"; tail -2 gpurun_out/bench_r02_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r02_n$n.json')); print('value',round(d['value']),'e2e',d['e2e'] and {k:d['e2e'][k] for k in ['value','ms_per_step','frac_of_feed_ceiling'] if k in d['e2e']},'one_process',d['e2e_one_process_all_gpus'] and round(d['e2e_one_process_all_gpus']['value']),'parity',d['parity'])"
done
