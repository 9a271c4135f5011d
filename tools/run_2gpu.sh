python -m pytest tests -m gpu -x -q > gpurun_out/pytest_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_2gpu.log
tail -n 4 gpurun_out/pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err || echo "N=2 failed"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/scale_ref_n2.json 2> gpurun_out/scale_ref_n2.err || echo "ref N=2 failed"
python tools/sw_sweep.py > gpurun_out/sw_sweep.json 2> gpurun_out/sw_sweep.err; echo "sweep rc=$?"
tail -n 1 gpurun_out/scale_n2.json | cut -c1-200
