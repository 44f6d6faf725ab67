"""End-to-end respond through chpir_cluster_server_respond on N GPUs, swept over caller counts and ingest routes (one process).

    python tools/e2e_probe.py --gpus 2 [--log2n 20] [--threads 32,64,128,256] [--queries 256] [--calls 4096]

Prints, per setting, queries/s, the mean coalesced batch and the leaders' wall time per pipeline stage (chpir_cluster_server_info)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--arity", type=int, default=3)
    ap.add_argument("--threads", default="32,64,128,256")
    ap.add_argument("--queries", type=int, default=256)
    ap.add_argument("--calls", type=int, default=4096)
    ap.add_argument("--routes", default="batch,pull,dma")
    ap.add_argument("--cuts", default="rows")
    args = ap.parse_args()
    import torch

    import bench
    import chalametpir_b200 as cp

    cluster = cp.Cluster(n_gpus=args.gpus)
    out = []
    for cut in args.cuts.split(","):
        os.environ["CHPIR_CLUSTER_SHARD"] = cut
        srv, _, _, plans, (b, K, N), _ = bench.make_cluster_server(cp, torch, cluster, args.log2n, args.arity, True, batch_tc=1, respond_coalesce=True)
        os.environ.pop("CHPIR_CLUSTER_SHARD", None)
        Q = args.queries
        qlen, rlen = 8 + 4 * K, 8 + 4 * N
        q_pin, r_pin = cp.PinnedBuffer(Q * qlen), cp.PinnedBuffer(Q * rlen)
        qh = q_pin.array.reshape(Q, qlen)
        rng = np.random.default_rng(3)
        for i in range(Q):
            qh[i, :8] = np.array([1, K], dtype="<u4").view(np.uint8)
            qh[i, 8:] = rng.integers(0, 256, size=4 * K, dtype=np.uint8)
        ptrs = [q_pin.ptr + i * qlen for i in range(Q)]
        for route in args.routes.split(","):
            os.environ["CHPIR_CLUSTER_INGEST"] = route
            for th in [int(x) for x in args.threads.split(",")]:
                srv.respond_concurrent(ptrs, qlen, 2 * Q, r_pin.ptr, rlen, th)
                i0 = srv.get_info()
                sec = srv.respond_concurrent(ptrs, qlen, args.calls, r_pin.ptr, rlen, th)
                i1 = srv.get_info()
                nb = i1["batches"] - i0["batches"]
                row = {"gpus": args.gpus, "cut": cut, "route": route, "threads": th, "qps": round(args.calls / sec), "mean_batch": round(args.calls / max(1, nb), 1),
                       "tc_batches": i1["tc_batches"] - i0["tc_batches"], "batches": nb}
                for k in ("ingest_wait_s", "ingest_s", "exec_wait_s", "exec_s"):
                    row[k + "_per_batch_ms"] = round((i1[k] - i0[k]) / max(1, nb) * 1e3, 3)
                print(json.dumps(row), flush=True)
                out.append(row)
            os.environ.pop("CHPIR_CLUSTER_INGEST", None)
        lat = sorted(srv.respond_concurrent(ptrs[:1], qlen, 1, r_pin.ptr, rlen, 1) * 1e3 for _ in range(30))
        print(json.dumps({"cut": cut, "single_caller_ms_median": round(lat[15], 4), "min": round(lat[0], 4)}), flush=True)
        srv.close()
        q_pin.close()
        r_pin.close()


if __name__ == "__main__":
    main()
