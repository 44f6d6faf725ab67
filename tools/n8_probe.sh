# 8-GPU experiments around the device-resident respond loop of bench.py (run as: gpurun --gpus 8 -- bash tools/n8_probe.sh)
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --skip-hint --no-e2e-setup --no-batch-tc --no-cpu-baseline"
CHPIR_BENCH_PROFILE=24 timeout 200 $R > gpurun_out/n8_prof.json 2> gpurun_out/n8_prof.err; echo rc=$?
CHPIR_BENCH_NO_BCAST=1 timeout 200 $R > gpurun_out/n8_nobcast.json 2> gpurun_out/n8_nobcast.err; echo rc=$?
CHPIR_RING_Q_PER_CTA=1 timeout 200 $R > gpurun_out/n8_qpc1.json 2> gpurun_out/n8_qpc1.err; echo rc=$?
CHPIR_NCCL_MAX_CTAS=12 timeout 200 $R > gpurun_out/n8_ctas12.json 2> gpurun_out/n8_ctas12.err; echo rc=$?
for f in n8_prof n8_nobcast n8_qpc1 n8_ctas12; do python - <<PY
import json
d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
print("$f", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "us/q kernel", round(d["respond_us_per_query_kernel"],1), "ms/step", round(d["ms_per_step"],3))
PY
done
grep "respond kernels" gpurun_out/n8_prof.err
