"""ncu workload for `roofline.traffic`: a few launches of the respond GEMV at the benchmark's shape on ONE GPU -- the full 940-column
matrix (N = 1) and what one rank of an n-way sharded server streams: its row block (k_pitch x N, the default cut, keys ".../rows") or
its column slice (CHPIR_CLUSTER_SHARD=cols).  A rank's kernel only ever sees its own shard.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:respond_ring --csv \
        --log-file gpurun_out/traffic.csv python tools/traffic_probe.py
    python tools/traffic_probe.py --parse gpurun_out/traffic.csv        # -> profiles/respond_traffic.json (per query, per rank)
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = bytes(range(32))
Q = 16
# (log2 entries, arity, ranks, cut)
SHAPES = [(20, 3, 1, "cols"), (20, 3, 2, "rows"), (20, 3, 4, "rows"), (20, 3, 8, "rows"), (22, 3, 8, "rows"), (20, 3, 2, "cols"), (20, 3, 4, "cols"),
          (20, 3, 8, "cols"), (20, 4, 8, "cols"), (18, 3, 1, "cols")]


def workload():
    import torch

    import chalametpir_b200 as cp

    for log2n, arity, n, cut in SHAPES:
        b = cp.find_mat_elem_bit_len(1 << log2n)
        K, N = cp.db_matrix_shape(arity, 1 << log2n, 1024, b)
        pl = cp.cluster_plan(n, 0, K, N)
        if cut == "rows":
            K, nc = pl["k_pitch"], N
        else:
            nc = pl["col_count"]
        D = torch.randint(0, 1 << b, (K, nc), dtype=torch.int32, device="cuda")
        srv, _ = cp.Server.setup_from_device_matrix(SEED, D.data_ptr(), K, nc, b, skip_hint=True)
        del D
        q = torch.randint(-2**31, 2**31 - 1, (Q, K), dtype=torch.int32, device="cuda")
        r = torch.empty((Q, nc), dtype=torch.int32, device="cuda")
        for _ in range(3):  # the last launch of each shape is the one kept
            srv.respond_device(q.data_ptr(), Q, r.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        srv.close()


def parse(path):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    per_launch = {}
    for r in rows:
        per_launch.setdefault(r["ID"], {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    launches = [per_launch[k] for k in sorted(per_launch, key=int)]
    assert len(launches) == 3 * len(SHAPES), f"{len(launches)} respond launches in the log, expected {3 * len(SHAPES)}"
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for i, (log2n, arity, n, cut) in enumerate(SHAPES):
        m = launches[3 * i + 2]
        rd = m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
        wr = m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
        t = m["gpu__time_duration.sum"]
        out[f"2^{log2n}/{arity}/n{n}" + ("/rows" if cut == "rows" else "")] = {"dram_bytes_per_query": (rd + wr) / Q, "dram_read_bytes_per_query": rd / Q, "queries_per_launch": Q,
                                        "kernel_time_under_ncu": f"{t[0]} {t[1]}",
                                        "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one 16-query launch of respond_ring_kernel on one rank's shard "
                                                  "(tools/traffic_probe.py, round 2)"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "respond_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--parse":
        parse(sys.argv[2])
    else:
        workload()
