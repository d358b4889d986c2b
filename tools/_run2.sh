timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu_final6.json 2> gpurun_out/r2_bench_2gpu_final6.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_2gpu_final6.json").read().strip().splitlines() if l.startswith("{")][-1])
print("n_gpus", d["n_gpus"], "value %.4g ms %.4g e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "scaling", d["scaling"])
s=d.get("sharded_io") or {}
print(s.get('ms_per_step_overlapped'), (s.get('peer_memory') or {}).get('ms_per_step'))
P
