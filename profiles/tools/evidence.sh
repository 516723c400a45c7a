set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
profiles/run_ncu.sh ${TAG:-r01u} 27
python bench.py > gpurun_out/bench_${TAG:-r01u}.json 2> gpurun_out/bench_${TAG:-r01u}.err; tail -c 600 gpurun_out/bench_${TAG:-r01u}.json
