# bench.py at N = 1, 2, 4, 8 ranks of one box (torchrun like the driver), short steps
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err || echo "N=$N failed"
done
python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
python - <<PY
import json
for N in (1,2,4,8):
    try:
        d=json.loads(open("gpurun_out/scale_n%d.json"%N).read().strip().splitlines()[-1])
        print(N, round(d["value"]/1e6,2), round(d["e2e"]["value"]/1e6,2), round(d["ms_per_step"],1))
    except Exception as e:
        print(N, "ERR", e)
PY
