set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
profiles/run_ncu.sh ${TAG:-r01t} 27
python bench.py > gpurun_out/bench_${TAG:-r01t}.json 2> gpurun_out/bench_${TAG:-r01t}.err; tail -c 600 gpurun_out/bench_${TAG:-r01t}.json
