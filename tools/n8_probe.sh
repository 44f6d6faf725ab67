# 8-GPU runs of bench.py (gpurun --gpus 8 -- bash tools/n8_probe.sh): a kernel timeline of the device-resident respond loop on rank 0
# (CHPIR_BENCH_PROFILE, quick flags), then the full bench line with the host-pipelined setup.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8"
CHPIR_BENCH_PROFILE=24 timeout 200 $T --skip-hint --no-e2e-setup --no-batch-tc --no-cpu-baseline > gpurun_out/n8_prof.json 2> gpurun_out/n8_prof.err; echo rc=$?
timeout 300 $T --a-expand host > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?
for f in n8_prof bench_n8; do python - <<PY
import json
d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
print("$f", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "us/q kernel", round(d["respond_us_per_query_kernel"],1), "ms/step", round(d["ms_per_step"],3), d["roofline"]["frac"], d.get("batched_respond_tc"))
PY
done
grep "respond kernels" gpurun_out/n8_prof.err
