timeout 1000 python -m pytest tests -m gpu -q > gpurun_out/gputests_r2m.log 2>&1; echo tests rc=$?; tail -n 2 gpurun_out/gputests_r2m.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_r2m.json 2> gpurun_out/bench_ref_r2m.err; echo ref rc=$?
timeout 600 python bench.py --gpus 1 > gpurun_out/bench_r2m_default.json 2> gpurun_out/bench_r2m_default.err; echo bench-default rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
