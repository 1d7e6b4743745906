# 1/2/4/8-GPU scaling of bench.py on one box
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py"; fi
  $CMD --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_$n.json
  python - <<PY
import json
d=json.load(open("gpurun_out/scale_$n.json"))
print("N=$n ray", round(d["value"]), "Mrays/s e2e", round(d["e2e"]["value"]), "| raster", round(d["secondary"]["value"]), "Mtris/s e2e", round(d["secondary"]["e2e"]["value"]), "|", d["config"]["partition"][:50])
PY
done
