python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke_final.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
