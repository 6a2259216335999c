set -x
run() { # N config steps
  if [ "$1" = "1" ]; then python bench.py --gpus 1 --config $2 --steps $3 --warmup 5 --no-cpu-baseline; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500+$1)) bench.py --gpus $1 --config $2 --steps $3 --warmup 5 --no-cpu-baseline; fi
}
for n in 1 2 8; do run $n cfg2 20 > gpurun_out/r2_scale_cfg2_n$n.json 2> gpurun_out/r2_scale_cfg2_n$n.err; done
for n in 1 2 8; do run $n cfg5 50 > gpurun_out/r2_scale_cfg5_n$n.json 2> gpurun_out/r2_scale_cfg5_n$n.err; done
DRB_BENCH_GATE=0 run 8 cfg2 20 > gpurun_out/r2_scale_cfg2_n8_nogate.json 2> /dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_scale_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "N",d["n_gpus"], "value %.1fM"%(d["value"]/1e6), "ms %.4f"%d["ms_per_step"], "e2e %.1fM"%(d["e2e"]["value"]/1e6), d["config"].get("allreduce"))
    except Exception as e: print(f, "ERR", e)
PY
