"""Tensor-core limb GEMM (csrc/gemm_tc.cu) on ONE GPU: the CTA-pair kernel (cta_group::2) against the one-SM kernel swept over the
thread-block cluster size (A-tile multicast).

    python tools/gemm_sweep.py [--kernels pair,1sm] [--clusters 1,2,4,8] [--shapes full,row8,p18] [--batches 128,64,32]

Shapes: full = 2^20 entries 3-wise (K = 1 179 648, N = 940, b = 9: one rank's batched respond = one 128-row panel of the hint GEMM),
row8 = one rank's row block of an 8-GPU cluster (K = 147 456), p18 = 2^18 entries (K = 303 104, N = 846, b = 10: 7 -> 8 N tiles).
Per setting: ms per 128-query and 64-query batch (CUDA events, 20 launches after warm-up), issued int8 TOP/s, and whether the
first rows equal the streaming GEMV's."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = bytes(range(32))
SHAPES = {"full": (1179648, 940, 9), "row8": (147456, 940, 9), "p18": (303104, 846, 10), "slice2": (1179648, 470, 9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", default="1,2,4,8")
    ap.add_argument("--shapes", default="full,row8,p18")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--kernels", default="pair,1sm", help="pair = CTA pairs (cta_group::2, the default), 1sm = the one-SM kernel (swept over --clusters)")
    ap.add_argument("--batches", default="128,64,32")
    args = ap.parse_args()
    import torch

    import chalametpir_b200 as cp

    st = torch.cuda.current_stream()
    for name in args.shapes.split(","):
        K, N, b = SHAPES[name]
        D = torch.randint(0, 1 << b, (K, N), dtype=torch.int32, device="cuda")
        q = torch.randint(-2**31, 2**31 - 1, (128, K), dtype=torch.int32, device="cuda")
        ref = torch.zeros((4, N), dtype=torch.int32, device="cuda")
        settings = []
        for kern in args.kernels.split(","):
            settings += [("pair", 2)] if kern == "pair" else [("1sm", int(x)) for x in args.clusters.split(",")]
        for si, (kern, csz) in enumerate(settings):
            os.environ["CHPIR_GEMM_KERNEL"] = kern
            os.environ["CHPIR_GEMM_CLUSTER"] = str(csz)
            srv, _ = cp.Server.setup_from_device_matrix(SEED, D.data_ptr(), K, N, b, skip_hint=True, batch_tc=1)
            if si == 0:
                srv.respond_device(q.data_ptr(), 4, ref.data_ptr(), st.cuda_stream)
                torch.cuda.synchronize()
            row = {"shape": name, "K": K, "N": N, "b": b, "kernel": kern, "cluster": csz}
            for nq in [int(x) for x in args.batches.split(",")]:
                out = torch.zeros((nq, N), dtype=torch.int32, device="cuda")
                for _ in range(3):
                    srv.respond_device_tc(q.data_ptr(), nq, out.data_ptr(), st.cuda_stream)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(args.iters):
                    srv.respond_device_tc(q.data_ptr(), nq, out.data_ptr(), st.cuda_stream)
                e1.record(st)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                nlimb = 7 if b > 8 else 4
                row[f"ms_{nq}"] = round(ms, 4)
                row[f"qps_{nq}"] = round(nq / ms * 1e3)
                row[f"issued_tops_{nq}"] = round(nlimb * 2 * 128 * K * (-(-N // 128) * 128) / (ms * 1e-3) / 1e12, 1)
                row[f"equal_gemv_{nq}"] = bool(torch.equal(out[:4], ref))
            print(json.dumps(row), flush=True)
            srv.close()
        del D, q
        torch.cuda.empty_cache()
    os.environ.pop("CHPIR_GEMM_CLUSTER", None)
    os.environ.pop("CHPIR_GEMM_KERNEL", None)


if __name__ == "__main__":
    main()
