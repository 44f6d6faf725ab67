#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native ChalametPIR server hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration `metric` is quoted on): 2^20 entries x 32-byte keys x 1 kB values,
3-wise XOR binary fuse filter  ->  D is K x N = 1 179 648 x 940 with 9-bit entries, LWE dimension 1774.
D is synthetic (uniform 9-bit entries from a counter hash, generated directly in HBM; SURVEY.md section 8d "synthetic-D").

A "step" is one pass of Server::respond over a batch of `--queries-per-step` independent queries: every query streams the whole
resident D once (single-query GEMV, the reference's `server_respond`).  `value` is whole-job queries/s with the queries already
resident in HBM; `e2e` is the same metric through the C ABI call the reference's `Server::respond` would bind
(`chpir_cluster_server_respond`) with HOST buffers (pinned), H2D of each query and D2H of each response inside the timed region,
called by concurrent native threads as the reference's example server does with one task per connection.

N > 1: the server is ONE process driving N GPUs (chpir_cluster_*, csrc/cluster.cu) -- the reference's `Server::respond(&self, &[u8])`
has no room for a rank argument, so the sharding lives behind the handle.  Under torchrun every rank joins the process
group (NCCL for the communicator, gloo for the CPU-side barriers around the timed regions); rank 0 is the server process and owns all N
GPUs, ranks 1..N-1 hold no data and wait at the barriers.  Setup cuts D by COLUMNS (hint slices, gathered over NCCL); respond cuts it
by ROWS: GPU r keeps rows [k0_r, k0_r + K/N) at full width and needs only the words of a query it ingested over its own PCIe link,
so queries are resident K-sliced over the GPUs exactly as the ingest leaves them, no query word crosses NVLink, and GPU 0 adds the
N partial responses (exact mod 2^32) with one small kernel reading peer memory.  The round-1 column cut for respond is timed beside
it (`column_cut_comparison`).  The database is the same size at every N, so `scaling` is "strong".

`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP over all host cores) on the same workload at FULL
size; the Rust reference itself cannot be built in this image (no cargo/rustc), see DESIGN.md.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

SEED_MU = bytes((7 * i + 3) & 0xFF for i in range(32))
LWE = 1774
VALUE_BYTES = 1024
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
FALLBACK_BF16_TFLOPS = 1590.0
PUBLISHED = {"server_respond_2^20_3wise_ms": {"m8g.8xlarge": 10.06, "m7i.8xlarge": 14.06},
             "server_setup_2^20_3wise_s": {"m8g": 577, "m7i": 1282, "g6e(L40S offload)": 25.58}}

_JSON_OUT = None


def _jsonable(o):
    if isinstance(o, np.generic):
        return o.item()
    raise TypeError(f"not JSON serialisable: {type(o).__name__}")


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line, default=_jsonable) + "\n")
    out.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def shape_of(log2n: int, arity: int, use_oracle: bool = False):
    """(b, K, N) from the reference's formulas (server.rs:193-218, binary_fuse_filter.rs:52-67/:261-276, matrix.rs:699-700)."""
    n = 1 << log2n
    if use_oracle:  # the CPU arm must not touch the product library
        from oracle import oracle as O

        b = O.find_mat_elem_bit_len(n)
        K = O.filter_shape(arity, n)[2]
        N = -(-(256 + 8 * VALUE_BYTES + 8) // b)
    else:
        import chalametpir_b200 as cp

        b = cp.find_mat_elem_bit_len(n)
        K, N = cp.db_matrix_shape(arity, n, VALUE_BYTES, b)
    return b, K, N


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", FALLBACK_BF16_TFLOPS)),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", FALLBACK_BF16_TFLOPS))), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": FALLBACK_HBM_GBS, "bf16_tflops": FALLBACK_BF16_TFLOPS, "bf16_tflops_sustained": FALLBACK_BF16_TFLOPS, "source": "fallback"}


def gen_d_slice(torch, K: int, c0: int, nc: int, b: int, device, salt: int = 0x5EED):
    """D[k][n] = hash(k, n) mod 2^b for n in [c0, c0+nc): every rank's slice of the same matrix, generated where it will live."""
    D = torch.empty((K, nc), dtype=torch.int32, device=device)
    cols = torch.arange(c0, c0 + nc, device=device, dtype=torch.int64)
    M32 = 0xFFFFFFFF
    chunk = 1 << 16
    for r0 in range(0, K, chunk):
        r1 = min(K, r0 + chunk)
        rows = torch.arange(r0, r1, device=device, dtype=torch.int64)
        x = (rows[:, None] * 0x9E3779B1 + cols[None, :] * 0x85EBCA77 + salt) & M32
        x ^= x >> 15
        x = (x * 0x2C1B3C6D) & M32
        x ^= x >> 12
        x = (x * 0x297A2D39) & M32
        x ^= x >> 15
        D[r0:r1] = (x & ((1 << b) - 1)).to(torch.int32)
    return D


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]  # samples taken while the GPU was busy: the upper half of the clock readings
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w": round(statistics.median(pw), 1), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def use_all_host_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is the reference's rayon pool = one thread per hardware thread.
    Must run before the oracle library (libgomp) is loaded."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return n


def oracle_with_all_cores(want: int):
    """The oracle module with its OpenMP pool sized to every usable host cpu (also when libgomp was loaded earlier by someone else)."""
    from oracle import oracle as O

    O.set_num_threads(want)
    return O


def host_d(K: int, N: int, b: int, seed: int = 1234):
    rng = np.random.default_rng(seed)
    D = np.empty((K, N), dtype=np.uint32)
    step = 1 << 16
    for r0 in range(0, K, step):  # chunked: no 2-byte temporary of the whole matrix
        r1 = min(K, r0 + step)
        D[r0:r1] = rng.integers(0, 1 << b, size=(r1 - r0, N), dtype=np.uint16)
    return D


def cpu_server(K: int, N: int, b: int):
    """The CPU restatement of the reference's Server (transposed, row_wise_compress'ed D resident in host memory) at FULL size."""
    from oracle import oracle as O

    D = host_d(K, N, b)
    srv, _ = O.Server.setup_from_matrix(SEED_MU, D, b, want_hint=False)
    del D
    return srv


def cpu_queries(K: int, n: int, seed: int):
    from oracle import oracle as O

    rng = np.random.default_rng(seed)
    return [O.matrix_to_bytes(rng.integers(0, 2**32, size=K, dtype=np.uint64).astype(np.uint32).reshape(1, -1)) for _ in range(n)]


def cpu_respond_baseline(K: int, N: int, b: int, n_queries: int, want_threads: int):
    """Oracle (CPU restatement of Server::respond, matrix.rs:328-485) on the FULL database, a bounded number of queries."""
    O = oracle_with_all_cores(want_threads)
    srv = cpu_server(K, N, b)
    qs = cpu_queries(K, 4, 99)
    for q in qs[:2]:
        srv.respond(q)
    times = []
    for i in range(n_queries):
        t = time.perf_counter()
        srv.respond(qs[i % len(qs)])
        times.append(time.perf_counter() - t)
    threads = O.num_threads()
    return {"value": 1.0 / statistics.median(times), "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"full database (K={K}, N={N}, reference layout: transposed + row_wise_compress'ed in host memory), {n_queries} single-query responds, "
                      f"median {statistics.median(times) * 1e3:.2f} ms (min {min(times) * 1e3:.2f}); OpenMP threads={threads} of {want_threads} usable cpus",
            "threads_expected": want_threads}


def cpu_setup_extrapolated(K_full: int, N: int, b: int, frac: int, a_rows: int = 4):
    """CPU restatement of the reference's Server::setup cost at full size, EXTRAPOLATED from a bounded sample:
    the hint product M = A.D with the reference's own loop structure (matrix.rs:1040-1059: one fold per output element, D read
    column-wise; rayon -> OpenMP over outputs) on `a_rows` rows of A and the first K/frac rows of D, scaled by (1774 / a_rows) * frac;
    plus the TurboSHAKE128 squeeze of A (matrix.rs:541-558) timed on 8 MB of stream and scaled to 4 * 1774 * K bytes."""
    from oracle import oracle as O

    Ks = max(3, K_full // frac)
    D = host_d(Ks, N, b)
    A = O.generate_rows_from_seed(Ks, SEED_MU, 0, a_rows)
    O.matmul(A[:1], D, fast=False)  # warm the thread pool and the pages
    t = time.perf_counter()
    O.matmul(A, D, fast=False)
    gemm_sample_s = time.perf_counter() - t
    gemm_full_s = gemm_sample_s * (LWE / a_rows) * (K_full / Ks)
    nbytes = 8 << 20
    t = time.perf_counter()
    O.turboshake128(SEED_MU, nbytes)
    xof_sample_s = time.perf_counter() - t
    xof_full_s = xof_sample_s * (4.0 * LWE * K_full / nbytes)
    return {
        "extrapolated": True, "kind": "port", "cores": O.num_threads(),
        "hint_gemm_s": gemm_full_s, "expand_a_s": xof_full_s, "total_s": gemm_full_s + xof_full_s,
        "sample": f"hint GEMM: {a_rows} of {LWE} rows of A x rows [0,{Ks}) of K={K_full} (all {N} columns) in {gemm_sample_s:.3f} s, scaled by "
                  f"{LWE / a_rows * K_full / Ks:.0f}; XOF: {nbytes >> 20} MB of the stream in {xof_sample_s:.3f} s on one core, scaled by "
                  f"{4.0 * LWE * K_full / nbytes:.0f}; filter construction and row encoding not included",
    }


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's Server::respond on ALL host cores, full database (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    want = use_all_host_cores()
    b, K, N = shape_of(args.log2n, args.arity, use_oracle=True)
    O = oracle_with_all_cores(want)
    threads = O.num_threads()
    assert threads == want, f"the CPU arm must use every host core: OpenMP has {threads} threads, {want} cpus are usable"
    srv = cpu_server(K, N, b)
    sample_q = max(1, min(args.ref_queries_per_step, args.queries_per_step))
    qs = cpu_queries(K, sample_q, 5)
    for _ in range(max(1, args.warmup)):
        for qb in qs[: max(1, sample_q // 4)]:
            srv.respond(qb)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for qb in qs:
            srv.respond(qb)
    dt = time.perf_counter() - t0
    nq = args.steps * sample_q
    qps = nq / dt
    sample = (f"full database (K={K}, N={N}); each step answers a {sample_q}-query sample of the {args.queries_per_step}-query step, one query at a time "
              f"(the reference's server_respond); {nq} queries in {dt:.2f} s, measured, nothing extrapolated; OpenMP threads={threads} = all usable host cpus")
    line = {
        "impl": "reference", "metric": "server_respond_queries_per_s", "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "ms_per_step_covers_queries": sample_q, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, b, K, N),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement (oracle/chalamet_oracle.c) of the reference's Server::respond; the Rust reference cannot be built here (no cargo/rustc)",
    }
    emit(line)


def workload_config(args, b, K, N, by_rows=True):
    return {
        "workload": f"2^{args.log2n} entries x 32B keys x {VALUE_BYTES}B values, {args.arity}-wise XOR filter (BASELINE.json " +
                    {(16, 3): "configs[0]", (18, 3): "configs[1]", (20, 3): "configs[2]", (20, 4): "configs[3] shape", (22, 3): "configs[4]"}.get(
                        (args.log2n, args.arity), "shape outside configs") + ")",
        "K": K, "N": N, "mat_elem_bit_len": b, "lwe_dimension": LWE, "queries_per_step": args.queries_per_step,
        "sharding": (f"respond: rows of D (K)/{args.gpus}, partial responses summed on GPU 0; setup: columns/{args.gpus}; one process, chpir_cluster_* "
                     f"(csrc/cluster.cu)" if by_rows else f"columns/{args.gpus}, one process, chpir_cluster_* (csrc/cluster.cu)") if args.gpus > 1
        else "none (cluster of one GPU)",
        "l2": "inputs larger than L2 (every query streams the resident packed D: 1.28 GB / n_gpus per GPU vs 126 MB of L2)",
    }


# ------------------------------------------------------------------------------------------------ GPU arm helpers
class Ranks:
    """torchrun plumbing: NCCL group (communicator of `world` ranks), gloo group for CPU-side barriers that do not occupy any GPU."""

    def __init__(self, args):
        import torch

        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.cpu_pg = None
        if self.world > 1:
            import torch.distributed as dist

            self.dist = dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank), timeout=datetime.timedelta(hours=1))
            self.cpu_pg = dist.new_group(backend="gloo", timeout=datetime.timedelta(hours=1))
            t = torch.ones(1, device=f"cuda:{self.local_rank}")
            dist.all_reduce(t)  # every rank joins at once: the NCCL communicator exists, nobody spins on a GPU waiting for rank 0
            torch.cuda.synchronize()
            assert int(t.item()) == self.world

    def barrier(self):
        """CPU-side barrier (gloo): ranks 1..N-1 block in the kernel, not in a spinning NCCL kernel on their GPU."""
        if self.dist is not None:
            self.dist.barrier(group=self.cpu_pg)

    def max_over_ranks(self, values):
        if self.dist is None:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


N_SYNC_POINTS = 6  # barriers rank 0 passes in run_b200; ranks 1..N-1 pass the same number


def sync_devices(torch, devices):
    for d in devices:
        torch.cuda.synchronize(d)


def exact_columns(torch, q_row_dev, D_cols):
    """q . D[:, cols] mod 2^32, exact: 16-bit halves of q times entries < 2^14 over K <= 2^23 rows stay below 2^53 in float64."""
    q = q_row_dev.to(torch.int64) & 0xFFFFFFFF
    lo, hi = (q & 0xFFFF).double(), (q >> 16).double()
    Dd = D_cols.double()
    a = (lo @ Dd).to(torch.int64)
    c = (hi @ Dd).to(torch.int64)
    return ((a + ((c & 0xFFFF) << 16)) & 0xFFFFFFFF).cpu().numpy().astype(np.uint64)


def make_cluster_server(cp, torch, cluster, log2n, arity, skip_hint, keep_d=False, **opts):
    """Synthetic D of the named shape, generated slice by slice on the GPU that will own it, then Server::setup on the cluster."""
    b, K, N = shape_of(log2n, arity)
    n = cluster.n_gpus
    plans = [cp.cluster_plan(n, r, K, N) for r in range(n)]
    Ds = [gen_d_slice(torch, K, p["col_begin"], p["col_count"], b, f"cuda:{cluster.devices[r]}") for r, p in enumerate(plans)]
    sync_devices(torch, cluster.devices)
    t0 = time.perf_counter()
    srv, hint = cp.ClusterServer.setup_from_device_slices(cluster, SEED_MU, [d.data_ptr() for d in Ds], K, N, b, skip_hint=skip_hint, **opts)
    wall = time.perf_counter() - t0
    if not keep_d:
        Ds = None
        for d in cluster.devices:
            with torch.cuda.device(d):
                torch.cuda.empty_cache()
    return srv, hint, Ds, plans, (b, K, N), wall


def make_query_slices(torch, cluster, plans, K, ks, nq, seed, keep_full_rows=0):
    """nq uniform u32 queries, resident K-sliced over the GPUs as the PCIe ingest would leave them; returns (slices, first rows on GPU 0)."""
    dev0 = f"cuda:{cluster.devices[0]}"
    g = torch.Generator(device=dev0)
    g.manual_seed(seed)
    slices = [torch.zeros((nq, ks), dtype=torch.int32, device=f"cuda:{cluster.devices[r]}") for r in range(cluster.n_gpus)]
    head = torch.empty((min(keep_full_rows, nq), K), dtype=torch.int32, device=dev0) if keep_full_rows else None
    step = 32
    for r0 in range(0, nq, step):
        r1 = min(nq, r0 + step)
        q = torch.randint(-(2**31), 2**31, (r1 - r0, K), dtype=torch.int32, device=dev0, generator=g)
        for r, p in enumerate(plans):
            if p["k_count"]:
                slices[r][r0:r1, : p["k_count"]] = q[:, p["k_begin"]: p["k_begin"] + p["k_count"]].to(slices[r].device)
        if head is not None and r0 < head.shape[0]:
            head[r0: min(r1, head.shape[0])] = q[: min(r1, head.shape[0]) - r0]
    sync_devices(torch, cluster.devices)
    return slices, head


def check_against_exact(torch, cluster, plans, Ds, q_rows_dev0, out_rows_dev0, n_rows, cols_per_rank=3):
    """Rows of the gathered (nq x N) result on GPU 0 against exact dot products on columns drawn from EVERY rank's slice."""
    ok = True
    for r, p in enumerate(plans):
        nc = p["col_count"]
        local = sorted({0, nc // 2, nc - 1})[:cols_per_rank]
        Dc = Ds[r][:, local]
        for i in range(n_rows):
            want = exact_columns(torch, q_rows_dev0[i].to(Dc.device), Dc)
            got = out_rows_dev0[i].cpu().numpy().view(np.uint32)[[p["col_begin"] + c for c in local]].astype(np.uint64)
            ok = ok and bool(np.array_equal(got, want))
    return ok


def h2d_bound_live(cp, torch, cluster, plans, q_ptr0, qlen, Q, ks, reps=4):
    """The PCIe ceiling of the e2e leg, measured on this box with the e2e leg's own traffic: every GPU copies its K/n words of the Q
    page-locked queries (one strided copy-engine transfer per GPU and pass, all GPUs at once) -- no kernels, no batching logic.
    Returns (aggregate GB/s, queries/s that rate allows)."""
    from chalametpir_b200._lib import lib
    from chalametpir_b200.errors import check

    devs = cluster.devices
    dsts = [torch.empty((Q, ks), dtype=torch.int32, device=f"cuda:{d}") for d in devs]
    streams = [torch.cuda.Stream(device=d) for d in devs]

    def one_pass():
        for r, p in enumerate(plans):
            if p["k_count"]:
                with torch.cuda.device(devs[r]):
                    check(lib.chpir_upload_rows(dsts[r].data_ptr(), ks * 4, q_ptr0 + 8 + 4 * p["k_begin"], qlen, p["k_count"] * 4, Q, streams[r].cuda_stream))

    one_pass()
    sync_devices(torch, devs)
    t0 = time.perf_counter()
    for _ in range(reps):
        one_pass()
    sync_devices(torch, devs)
    dt = time.perf_counter() - t0
    nbytes = reps * Q * 4 * sum(p["k_count"] for p in plans)
    return nbytes / dt / 1e9, reps * Q / dt


def int8_peak_live(torch, dev):
    """MEASURED_PEAKS.json has no int8 figure: measure the library's dense int8 GEMM here (cuBLASLt through torch._int_mm, 8192^3, best of
    10); fall back to 2 x the measured bf16 figure."""
    pk = peaks()
    peak, src = 2.0 * pk["bf16_tflops"], f"2 x bf16_tflops ({pk['source']})"
    try:
        with torch.cuda.device(dev):
            ia = torch.randint(-100, 100, (8192, 8192), dtype=torch.int8, device=dev)
            ib = torch.randint(-100, 100, (8192, 8192), dtype=torch.int8, device=dev)
            for _ in range(3):
                torch._int_mm(ia, ib)
            best = 1e9
            for _ in range(10):
                i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                i0.record()
                torch._int_mm(ia, ib)
                i1.record()
                torch.cuda.synchronize()
                best = min(best, i0.elapsed_time(i1))
            peak, src = 2 * 8192**3 / best / 1e9, "measured live: torch._int_mm 8192^3 (cuBLASLt int8), best of 10"
    except Exception:
        pass
    return peak, src


def rank_timing(srv, n):
    infos = [srv.shard_info(r) for r in range(n)]
    tmax = {k: max(i["timing"][k] for i in infos) for k in infos[0]["timing"]}
    return infos, tmax


def batched_leg(cp, torch, cluster, srv, plans, K, N, b, BQ, iters, seed, Ds=None, label=""):
    """Batched respond on the tensor cores (north_star (2)): BQ device-resident queries, one pass over every rank's byte planes."""
    slices, head = make_query_slices(torch, cluster, plans, K, srv.k_pitch, BQ, seed, keep_full_rows=2)
    dev0 = f"cuda:{cluster.devices[0]}"
    out_tc = torch.zeros((BQ, N), dtype=torch.int32, device=dev0)
    out_gv = torch.zeros((BQ, N), dtype=torch.int32, device=dev0)
    ptrs = [s.data_ptr() for s in slices]
    srv.respond_device(ptrs, BQ, out_tc.data_ptr(), mode=cp.RESPOND_TC, repeats=2)  # warm-up
    nchk = min(BQ, 8)
    srv.respond_device(ptrs, nchk, out_gv.data_ptr(), mode=cp.RESPOND_GEMV, repeats=1)
    sync_devices(torch, cluster.devices)
    parity = {"tc_equals_gemv": bool(torch.equal(out_tc[:nchk], out_gv[:nchk]))}
    if Ds is not None:
        parity["tc_equals_exact_on_columns_of_every_rank"] = check_against_exact(torch, cluster, plans, Ds, head, out_tc, 2)
    ms = srv.respond_device(ptrs, BQ, out_tc.data_ptr(), mode=cp.RESPOND_TC, repeats=iters) / iters
    nlimb = 7 if b > 8 else 4
    tiles = -(-BQ // 128)
    pair = not os.environ.get("CHPIR_GEMM_KERNEL", "").startswith("1")
    # rows of the query operand the MMAs are issued for: the CTA-pair kernel rounds every pass to 32 rows, the one-SM kernel always pays for 128
    rows_issued = sum(-(-min(128, BQ - 128 * t) // 32) * 32 for t in range(tiles)) if pair else 128 * tiles

    def n_pad(nc):  # columns of D the MMAs are issued for: 128 per CTA, CTAs in pairs (one-SM kernel: tiles of <= 128 columns)
        if pair:
            return -(-nc // 256) * 256
        return -(-nc // 16) * 16 if nc <= 128 else -(-nc // 128) * 128

    by_rows = bool(srv.get_info()["respond_by_rows"])
    if by_rows:  # every rank: rows_issued x N_pad x k_pitch
        issued = nlimb * 2 * rows_issued * srv.k_pitch * n_pad(N) * cluster.n_gpus
        plane_bytes = (2 if b > 8 else 1) * srv.k_pitch * N * cluster.n_gpus * tiles
        inc = "limb split of each rank's own query words, int8-limb GEMM over its rows of D on every rank, sum of the partial responses on GPU 0 (peer reads)"
    else:
        issued = sum(nlimb * 2 * rows_issued * K * n_pad(p["col_count"]) for p in plans)
        plane_bytes = sum((2 if b > 8 else 1) * K * p["col_count"] for p in plans) * tiles
        inc = "NVLink gather of the K-sliced queries fused with the limb split, int8-limb GEMM on every rank, strided peer copy of the columns to GPU 0"
    return {"label": label, "queries_per_batch": BQ, "ms_per_batch": ms, "queries_per_s": BQ / (ms * 1e-3),
            "includes": inc,
            "issued_int8_tops_all_gpus": issued / (ms * 1e-3) / 1e12, "d_plane_bytes_streamed_per_batch_all_gpus": plane_bytes,
            "d_plane_gbs_per_gpu": plane_bytes / cluster.n_gpus / (ms * 1e-3) / 1e9, "parity": parity}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch

    R = Ranks(args)
    world, rank = R.world, R.rank
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    if world > 1:
        args.gpus = world
    if rank != 0:
        # the server is ONE process (rank 0) that owns every GPU; this rank holds no data and only keeps the barriers' count
        for _ in range(N_SYNC_POINTS):
            R.barrier()
        R.max_over_ranks([0.0, 0.0, 0.0])
        R.close()
        return
    n_gpus = args.gpus
    if torch.cuda.device_count() < n_gpus:
        raise SystemExit(f"--gpus {n_gpus}: only {torch.cuda.device_count()} CUDA devices are visible")
    want_cpus = use_all_host_cores()  # for the cpu_baseline leg (libgomp reads it when the oracle is loaded)

    import chalametpir_b200 as cp

    cluster = cp.Cluster(n_gpus=n_gpus)
    devices = cluster.devices
    dev0 = f"cuda:{devices[0]}"
    torch.cuda.set_device(devices[0])
    Q = args.queries_per_step
    parity = {}

    # ---------------- Server::setup on the cluster (one-off; reported, not part of the timed steps)
    a_expand = {"host": "host", "device": "device", "auto": "auto"}[args.a_expand]
    srv, hint, Ds, plans, (b, K, N), setup_wall = make_cluster_server(cp, torch, cluster, args.log2n, args.arity, args.skip_hint, keep_d=True, batch_tc=1,
                                                                     a_expand=a_expand, respond_coalesce=True)
    info = srv.get_info()
    shard_infos, tmax = rank_timing(srv, n_gpus)
    setup = {"api": "chpir_cluster_server_setup_device (D synthetic, generated in HBM slice by slice)", "a_expand": a_expand, "wall_s": setup_wall,
             **{k: round(v, 6) for k, v in tmax.items()}, "timing_is": "max over ranks of each phase", "gemm_kernel_ms": max(i["gemm_ms"] for i in shard_infos),
             "hint_gather_s": info["hint_gather_s"], "hint_gather_uses_nccl": bool(info["gather_uses_nccl"]), "nccl_version": info["nccl_version"],
             "skipped_hint": bool(args.skip_hint), "xof_impl": cp.host_xof_impl()}
    if not args.skip_hint:
        # tensor roofline of the hint GEMM on rank 0: issued int8 ops = limb pairs x 2 x padded M x K x padded N, summed over the 128-row panels
        nc0 = plans[0]["col_count"]
        nlimb = 7 if b > 8 else 4
        m_pad = -(-LWE // 128) * 128
        pair = not os.environ.get("CHPIR_GEMM_KERNEL", "").startswith("1")
        n_pad = -(-nc0 // 256) * 256 if pair else (-(-nc0 // 128) * 128 if nc0 > 128 else -(-nc0 // 16) * 16)
        m_pad = (LWE // 128) * 128 + -(-(LWE % 128) // 32) * 32 if pair else m_pad  # the pair kernel issues the last panel in 32-row steps
        issued, useful = nlimb * 2 * m_pad * K * n_pad, nlimb * 2 * LWE * K * nc0
        int8_peak, int8_src = int8_peak_live(torch, dev0)
        gs = shard_infos[0]["gemm_ms"] * 1e-3
        setup["gemm_roofline"] = {"bound": "tensor", "kernel": ("gemm_tc_pair_kernel (tcgen05 cta_group::2 kind::i8 limb GEMM on CTA pairs" if pair else "gemm_tc_kernel (tcgen05 kind::i8 limb GEMM") + ", 14 panel launches, rank 0's column slice)",
                                  "achieved": issued / gs / 1e12, "useful": useful / gs / 1e12, "unit": "TOP/s",
                                  # denominators: the measured dense bf16 rate of this pool (MEASURED_PEAKS.json, burst) x 2 for int8 -- the
                                  # prescribed one --, the library's own int8 GEMM timed in this run, and the nominal 4.5 POP/s
                                  "peak": 2 * peaks()["bf16_tflops"], "peak_source": "2 x bf16_tflops (burst; every panel GEMM is timed alone) of MEASURED_PEAKS.json: " + peaks()["source"],
                                  "frac": issued / gs / 1e12 / (2 * peaks()["bf16_tflops"]),
                                  "library_int8_tops": int8_peak, "library_int8_source": int8_src, "frac_of_library_int8": issued / gs / 1e12 / int8_peak,
                                  "frac_of_nominal_4500": issued / gs / 1e12 / 4500.0, "u32_mac_equivalent_tmacs": LWE * K * nc0 / gs / 1e12}
        busy = tmax["xof_host_busy_s"] if tmax["xof_host_busy_s"] > 0 else tmax["expand_a_s"]
        setup["xof_ns_per_permutation"] = busy / (LWE * K * 4 / 168.0) * 1e9
        # hint parity: rows 0..1 of the GATHERED hint against exact products with the head of the XOF stream, on columns of every rank
        from oracle import oracle as O

        a0 = torch.from_numpy(O.generate_rows_from_seed(K, SEED_MU, 0, 2).view(np.int32).copy()).to(dev0)
        H = np.frombuffer(hint, dtype="<u4")
        assert H[0] == LWE and H[1] == N and len(hint) == 8 + 4 * LWE * N, "hint header / length"
        Hd = torch.from_numpy(H[2: 2 + 2 * N].view(np.int32).reshape(2, N).copy()).to(dev0)
        parity["hint_rows_0_1_on_columns_of_every_rank"] = check_against_exact(torch, cluster, plans, Ds, a0, Hd, 2)
        assert parity["hint_rows_0_1_on_columns_of_every_rank"], "gathered hint differs from the exact product"
        del a0, Hd
    R.barrier()  # sync point 1: setup done

    # ---------------- Server::setup(seed, db) from raw keys and values (server.rs:103) + a complete PIR round
    if args.e2e_setup and not args.skip_hint:
        n_db = 1 << args.log2n
        rs = np.random.default_rng(2024)
        keys = rs.integers(0, 256, size=(n_db, 32), dtype=np.uint8)
        keys[:, :8] = np.arange(n_db, dtype="<u8").view(np.uint8).reshape(n_db, 8)  # distinct by construction
        vals = rs.integers(0, 256, size=(n_db, VALUE_BYTES), dtype=np.uint8)
        sync_devices(torch, devices)
        t0 = time.perf_counter()
        srv_e, hint_e, fb_e = cp.ClusterServer.setup_from_arrays(cluster, SEED_MU, keys, vals, args.arity, filter_seed_rng=7, batch_tc=2, a_cache=True)
        wall_e = time.perf_counter() - t0
        _, te = rank_timing(srv_e, n_gpus)
        ie = srv_e.get_info()
        # a database UPDATE: same seed and keys, new values.  A depends on (seed, K) only and was left resident on every GPU (a_cache),
        # so this Server::setup walks no XOF chain.  The PIR round below runs against THIS server and hint.
        srv_e.close()
        vals = vals[::-1].copy()
        sync_devices(torch, devices)
        t0 = time.perf_counter()
        srv_u, hint_u, fb_u = cp.ClusterServer.setup_from_arrays(cluster, SEED_MU, keys, vals, args.arity, filter_seed_rng=8, batch_tc=2, a_cache=True,
                                                                 db_encode="device")
        wall_u = time.perf_counter() - t0
        _, tu = rank_timing(srv_u, n_gpus)
        assert tu["a_cache_hit"] == 1.0
        cached = cluster.drop_a_cache()
        t0 = time.perf_counter()
        client = cp.Client.setup(SEED_MU, hint_u, fb_u, device=devices[0], a_expand="host")
        client_setup_s = time.perf_counter() - t0
        rounds, q_ms = 0, []
        for j, i in enumerate([0, 1, n_db // 3, n_db // 2, n_db - 1, 777_777 % n_db]):
            key = keys[i].tobytes()
            try:
                t0 = time.perf_counter()
                qb = client.query(key, rng_seed=j)
                q_ms.append((time.perf_counter() - t0) * 1e3)
            except cp.ChalametPIRError as ex:
                if ex.variant != "ArithmeticOverflowAddingQueryIndicator":
                    raise
                continue
            assert client.process_response(key, srv_u.respond(qb)) == vals[i].tobytes(), "PIR round failed to recover the value"
            rounds += 1
        ci = client.info()
        assert rounds >= 3
        parity["full_size_pir_round_values_recovered"] = rounds
        setup["e2e_from_db"] = {
            "api": "chpir_cluster_server_setup_from_db (keys + values in host memory -> resident sharded server + complete hint + filter params)",
            "wall_s": wall_e, **{k: round(v, 6) for k, v in te.items()}, "hint_gather_s": ie["hint_gather_s"], "db_entries": n_db, "key_bytes": 32,
            "value_bytes": VALUE_BYTES, "hint_bytes": len(hint_e), "filter_param_bytes": len(fb_e),
            "after_database_update_with_cached_a": {"wall_s": wall_u, **{k: round(v, 6) for k, v in tu.items()}, "cached_a_bytes_per_gpu": cached[0],
                                                    "db_encode": "device (every GPU fills its own columns of D)"},
            "pir_round_gpu_client": {"values_recovered": rounds, "client_setup_s": client_setup_s, "client_query_ms_wall": statistics.median(q_ms),
                                     "client_query_kernel_ms": ci["last_query_kernel_ms"]}}
        client.close()
        srv_u.close()
        del hint_e, hint_u, keys, vals, client, srv_e, srv_u
    del hint

    # ---------------- queries: resident in HBM, K-sliced over the GPUs
    ks = srv.k_pitch
    KEEP = min(Q, 64)
    q_slices, q_head = make_query_slices(torch, cluster, plans, K, ks, Q, 1000, keep_full_rows=KEEP)
    q_ptrs = [t.data_ptr() for t in q_slices]
    out0 = torch.zeros((Q, N), dtype=torch.int32, device=dev0)

    ms1 = srv.respond_device(q_ptrs, Q, out0.data_ptr(), mode=cp.RESPOND_GEMV, repeats=1)
    sync_devices(torch, devices)
    parity["respond_gathered_result_on_columns_of_every_rank"] = check_against_exact(torch, cluster, plans, Ds, q_head, out0, 2)
    assert parity["respond_gathered_result_on_columns_of_every_rank"], "gathered response differs from the exact dot product"
    Ds = None
    for d in devices:
        with torch.cuda.device(d):
            torch.cuda.empty_cache()

    # ---------------- timed region: K steps, device-resident (CUDA events on GPU 0 around everything every GPU does)
    srv.respond_device(q_ptrs, Q, out0.data_ptr(), mode=cp.RESPOND_GEMV, repeats=max(args.warmup, 3))
    sampler = ClockSampler(devices[0])
    sampler.start()
    sync_devices(torch, devices)
    R.barrier()  # sync point 2
    w0 = time.perf_counter()
    ms_total = srv.respond_device(q_ptrs, Q, out0.data_ptr(), mode=cp.RESPOND_GEMV, repeats=args.steps)
    sync_devices(torch, devices)
    wall_total_ms = (time.perf_counter() - w0) * 1e3
    R.barrier()  # sync point 3
    n_queries = args.steps * Q
    chunk = int(os.environ.get("CHPIR_CLUSTER_GEMV_CHUNK", "32"))

    # kernel-only timing of the dominant kernel (the streaming GEMV on rank 0's shard): same launches on rank 0's stream, nothing else
    # running -- this is what the roofline is computed from.  Row cut: the shard is k_pitch x N and reads rank 0's own query words.
    by_rows = bool(info["respond_by_rows"])
    launches = args.steps * (n_gpus + 1) if by_rows else args.steps * (-(-Q // chunk)) * n_gpus
    sh0 = srv.shard(0)
    sh0_cols = N if by_rows else plans[0]["col_count"]
    q_k0 = q_slices[0] if by_rows else q_head
    resp_k = torch.zeros((KEEP, sh0_cols), dtype=torch.int32, device=dev0)
    st = torch.cuda.current_stream(devices[0])
    k_iters = max(1, min(args.steps * Q, 4096) // KEEP)
    for _ in range(2):
        sh0.respond_device(q_k0.data_ptr(), KEEP, resp_k.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize(devices[0])
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(st)
    for _ in range(k_iters):
        sh0.respond_device(q_k0.data_ptr(), KEEP, resp_k.data_ptr(), st.cuda_stream)
    k1.record(st)
    torch.cuda.synchronize(devices[0])
    ms_kernel = k0.elapsed_time(k1) / (k_iters * KEEP)
    launches += k_iters
    if by_rows:
        # the shards' own partial sums (each rank's kernel alone, its own stream) must add up to what the cluster call produced
        acc = resp_k[:2].cpu().numpy().view(np.uint32).astype(np.uint64)
        for r in range(1, n_gpus):
            with torch.cuda.device(devices[r]):
                part = torch.zeros((2, N), dtype=torch.int32, device=f"cuda:{devices[r]}")
                srv.shard(r).respond_device(q_slices[r].data_ptr(), 2, part.data_ptr(), torch.cuda.current_stream(devices[r]).cuda_stream)
                torch.cuda.synchronize(devices[r])
                acc += part.cpu().numpy().view(np.uint32)
        parity["kernel_only_partials_sum_to_cluster_result"] = bool(np.array_equal(acc & 0xFFFFFFFF, out0[:2].cpu().numpy().view(np.uint32)))
        assert parity["kernel_only_partials_sum_to_cluster_result"]
    else:
        parity["kernel_only_equals_cluster_result"] = bool(torch.equal(resp_k[:2], out0[:2, : plans[0]["col_count"]]))
        assert parity["kernel_only_equals_cluster_result"]

    # ---------------- e2e: the C ABI with HOST buffers, concurrent native callers of chpir_cluster_server_respond
    qlen, rlen = 8 + 4 * K, 8 + 4 * N
    q_pin, r_pin = cp.PinnedBuffer(Q * qlen), cp.PinnedBuffer(Q * rlen)
    qh = q_pin.array.reshape(Q, qlen)
    hdr = np.array([1, K], dtype="<u4").view(np.uint8)
    for r0 in range(0, Q, 32):  # reassemble the whole queries on the host from the resident slices
        r1 = min(Q, r0 + 32)
        qh[r0:r1, :8] = hdr
        for r, p in enumerate(plans):
            if p["k_count"]:
                qh[r0:r1, 8 + 4 * p["k_begin"]: 8 + 4 * (p["k_begin"] + p["k_count"])] = q_slices[r][r0:r1, : p["k_count"]].cpu().numpy().view(np.uint8)
    threads = args.e2e_threads or min(256, 32 * n_gpus)
    q_host_ptrs = [q_pin.ptr + i * qlen for i in range(Q)]
    srv.respond_concurrent(q_host_ptrs, qlen, max(args.warmup, 3) * Q, r_pin.ptr, rlen, threads)
    sync_devices(torch, devices)
    R.barrier()  # sync point 4
    e2e_s = srv.respond_concurrent(q_host_ptrs, qlen, n_queries, r_pin.ptr, rlen, threads)
    sync_devices(torch, devices)
    R.barrier()  # sync point 5
    clocks = sampler.stop()
    e2e_qps = n_queries / e2e_s
    rh = r_pin.array.reshape(Q, rlen)
    got = rh[:, 8:].view(np.uint32)
    parity["e2e_bytes_equal_device_path"] = bool(np.array_equal(got, out0.cpu().numpy().view(np.uint32))) and bool(
        np.array_equal(rh[:, :8].view("<u4"), np.tile(np.array([1, N], dtype="<u4"), (Q, 1))))
    assert parity["e2e_bytes_equal_device_path"], "responses through the host-buffer C ABI differ from the device path"
    co = srv.get_info()
    # single-caller latency through the ABI: one 4.7 MB H2D (split over the GPUs' PCIe links), one GEMV per GPU, one D2H -- the reference's
    # `server_respond` bench (integrations/benches/online_phase.rs) measures exactly this call
    lat = []
    for _ in range(30):
        lat.append(srv.respond_concurrent(q_host_ptrs[:1], qlen, 1, r_pin.ptr, rlen, 1) * 1e3)
    single_ms = statistics.median(lat[5:])
    h2d_gbs, h2d_qps = h2d_bound_live(cp, torch, cluster, plans, q_pin.ptr, qlen, Q, ks)

    # ---------------- batched respond on the tensor cores, device-resident (128 fill one M tile)
    batched = None
    if not args.no_batch_tc:
        batched = batched_leg(cp, torch, cluster, srv, plans, K, N, b, args.batch_queries, max(3, args.steps // 4), 77, label="this workload")
        assert all(batched["parity"].values())
        parity["batched_tc_equals_gemv"] = True

    # ---------------- roofline of the dominant kernel (rank 0's shard)
    pk = peaks()
    pb0 = shard_infos[0]["packed_bytes"]
    k_rank0 = srv.k_pitch if by_rows else K
    streamed = pb0 + 4 * k_rank0 + 4 * sh0_cols  # bytes one query on this rank must move: resident packed shard + its query words + its result words
    cf = 3 if b in (9, 10) else (2 if b >= 11 else 4)
    ref_layout = 4 * sh0_cols * (-(-k_rank0 // cf)) + 4 * k_rank0 + 4 * sh0_cols
    achieved = streamed / (ms_kernel * 1e-3) / 1e9
    shard_desc = (f"GPU 0's row block ({k_rank0} of {K} rows x {N} columns)" if by_rows else "GPU 0's column slice") if n_gpus > 1 else "the whole matrix"
    roofline = {
        "bound": "hbm", "kernel": f"respond_ring_kernel<{b},4> (persistent streaming u32 GEMV over K-major bit-packed D, cp.async.bulk smem ring), {shard_desc}",
        "achieved": achieved, "peak": pk["hbm_gbs"], "peak_source": pk["source"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": None,
        "queries_per_launch": KEEP, "bytes_per_launch": KEEP * streamed, "bytes_per_query": streamed, "bytes_per_query_reference_layout": ref_layout,
        "achieved_reference_layout_gbs": ref_layout / (ms_kernel * 1e-3) / 1e9, "us_per_launch": ms_kernel * 1e3 * KEEP, "launches_timed": k_iters,
        "timed_region_ms": ms_kernel * KEEP * k_iters, "rows": k_rank0, "columns": sh0_cols, "row_pitch_bytes": shard_infos[0]["row_pitch_bytes"],
    }
    prof = os.path.join(ROOT, "profiles", "respond_traffic.json")
    if os.path.exists(prof):
        try:
            tr = json.load(open(prof)).get(f"2^{args.log2n}/{args.arity}/n{n_gpus}" + ("/rows" if by_rows and n_gpus > 1 else ""))
            if tr:
                roofline["traffic"] = tr["dram_bytes_per_query"] * KEEP if isinstance(tr, dict) else tr
                roofline["traffic_source"] = tr.get("source") if isinstance(tr, dict) else "profiles/respond_traffic.json"
        except Exception:
            pass

    # ---------------- the measured comparison SURVEY.md section 8e asks for: the same device-resident pass with respond cut by COLUMNS
    # (every rank needs the whole query: NVLink all-gather by the copy engines, 118-column packed slivers at N = 8)
    column_cut = None
    if n_gpus > 1 and by_rows and not args.no_column_cut:
        os.environ["CHPIR_CLUSTER_SHARD"] = "cols"
        try:
            srv_c, _, _, _, _, _ = make_cluster_server(cp, torch, cluster, args.log2n, args.arity, True)
        finally:
            os.environ.pop("CHPIR_CLUSTER_SHARD", None)
        out_c = torch.zeros((Q, N), dtype=torch.int32, device=dev0)
        c_steps = max(3, args.steps // 4)
        srv_c.respond_device(q_ptrs, Q, out_c.data_ptr(), mode=cp.RESPOND_GEMV, repeats=3)
        ms_c = srv_c.respond_device(q_ptrs, Q, out_c.data_ptr(), mode=cp.RESPOND_GEMV, repeats=c_steps)
        sync_devices(torch, devices)
        same = bool(torch.equal(out_c, out0))
        ci = srv_c.get_info()
        column_cut = {"value": c_steps * Q / (ms_c * 1e-3), "unit": "queries/s", "ms_per_step": ms_c / c_steps, "steps": c_steps,
                      "respond_by_rows": ci["respond_by_rows"], "packed_bytes_max_rank": ci["packed_bytes_max_rank"],
                      "row_cut_packed_bytes_max_rank": info["packed_bytes_max_rank"], "bytes_equal_row_cut": same}
        parity["column_cut_bytes_equal_row_cut"] = same
        assert same and ci["respond_by_rows"] == 0
        srv_c.close()
        del out_c, srv_c

    # ---------------- the other BASELINE.json configurations, each with its own parity flags (separate servers)
    del q_slices, q_head, out0, resp_k, sh0
    q_pin.close()
    r_pin.close()
    configs = {}
    if args.configs:
        srv.close()
        for d in devices:
            with torch.cuda.device(d):
                torch.cuda.empty_cache()
        configs = config_legs(cp, torch, cluster, args, n_gpus)

    # ---------------- CPU baseline on the host cores (N = 1 only): the full database
    cpu_baseline = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_respond_baseline(K, N, b, 20, want_cpus)
        try:  # the reference's setup cost on these host cores, extrapolated from a sample (a reported baseline, never a target)
            setup["cpu_baseline_extrapolated"] = cpu_setup_extrapolated(K, N, b, frac=2, a_rows=2)
        except Exception as ex:  # pragma: no cover -- diagnostics must never cost the bench line
            setup["cpu_baseline_extrapolated"] = {"error": repr(ex)}

    R.barrier()  # sync point 6
    ms_total, e2e_s_max, ms_kernel = R.max_over_ranks([ms_total, e2e_s, ms_kernel])
    qps = n_queries / (ms_total * 1e-3)
    line = {
        "metric": "server_respond_queries_per_s", "value": qps, "unit": "queries/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, b, K, N, by_rows),
        "timing": {"device_ms_total": ms_total, "wall_ms_total": wall_total_ms, "first_pass_ms": ms1,
                   "how": "CUDA events on GPU 0's stream: t0 before the first byte moves on any GPU, t1 after GPU 0 has waited for the last columns of every GPU "
                          "(chpir_cluster_server_respond_device); wall clock of the same call beside it"},
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": Q * 4 * K, "d2h_bytes_per_step": Q * 4 * N, "threads": threads,
                "api": "chpir_cluster_server_respond (C ABI, page-locked host buffers from chpir_host_alloc): concurrent callers are coalesced; per batch "
                       "every GPU fetches its K/n words of all member queries over its own PCIe link with one pull kernel, multiplies them with its rows "
                       "of D (tensor-core limb GEMM from 6 queries up, GEMV below) and GPU 0 sums the partial responses; batches are pipelined "
                       "(collect | PCIe | SMs)" if by_rows and n_gpus > 1 else
                       "chpir_cluster_server_respond (C ABI, pinned host buffers): concurrent callers are coalesced into shared launches (tensor-core limb "
                       "GEMM from 6 queries up, GEMV below)",
                "pulled_queries": co["pulled_queries"],
                "seconds": e2e_s, "coalesced_batches": co["batches"], "coalesced_tensor_core_batches": co["tc_batches"],
                "mean_batch": co["queries"] / max(1, co["batches"]),
                "single_caller_ms": single_ms, "single_caller_published_reference_ms": PUBLISHED["server_respond_2^20_3wise_ms"],
                "pcie_bound_queries_per_s": h2d_qps, "pcie_aggregate_h2d_gbs": h2d_gbs,
                "pcie_bound_note": "measured in this run: the same page-locked query buffers copied to the GPUs' HBM (each GPU its K/n words of every query, all "
                                   "GPUs at once, copy engines only) -- the rate at which this host can feed 4K bytes per query to the n GPUs",
                "stage_ms_per_batch": {k: round(co[k] / max(1, co["batches"]) * 1e3, 3) for k in ("ingest_wait_s", "ingest_s", "exec_wait_s", "exec_s")}},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
        "setup": setup,
        "respond_us_per_query_kernel": ms_kernel * 1e3,
        "batched_respond_tc": batched,
        "column_cut_comparison": column_cut,
        "reshard_s": info["reshard_s"],
        "configs": configs,
        "parity": parity,
        "published_reference": PUBLISHED,
    }
    emit(line)
    R.close()


def config_legs(cp, torch, cluster, args, n_gpus):
    """BASELINE.json configs other than the headline one, as driver-visible numbers with their own parity flags."""
    out = {}
    devices = cluster.devices

    def free():
        for d in devices:
            with torch.cuda.device(d):
                torch.cuda.empty_cache()

    def gemv_leg(log2n, arity, nq, with_hint, label):
        srv, hint, Ds, plans, (b, K, N), wall = make_cluster_server(cp, torch, cluster, log2n, arity, not with_hint, keep_d=True, batch_tc=2)
        infos, tmax = rank_timing(srv, n_gpus)
        slices, head = make_query_slices(torch, cluster, plans, K, srv.k_pitch, nq, 31 + log2n, keep_full_rows=2)
        dev0 = f"cuda:{devices[0]}"
        o = torch.zeros((nq, N), dtype=torch.int32, device=dev0)
        ptrs = [s.data_ptr() for s in slices]
        srv.respond_device(ptrs, nq, o.data_ptr(), mode=cp.RESPOND_GEMV, repeats=2)
        sync_devices(torch, devices)
        par = {"respond_equals_exact_on_columns_of_every_rank": check_against_exact(torch, cluster, plans, Ds, head, o, 2)}
        if with_hint:
            from oracle import oracle as O

            a0 = torch.from_numpy(O.generate_rows_from_seed(K, SEED_MU, 0, 1).view(np.int32).copy()).to(dev0)
            H = np.frombuffer(hint, dtype="<u4")
            Hd = torch.from_numpy(H[2: 2 + N].view(np.int32).reshape(1, N).copy()).to(dev0)
            par["hint_row_0_on_columns_of_every_rank"] = check_against_exact(torch, cluster, plans, Ds, a0, Hd, 1) and H[0] == LWE and H[1] == N
        iters = 5
        ms = srv.respond_device(ptrs, nq, o.data_ptr(), mode=cp.RESPOND_GEMV, repeats=iters) / iters
        gi = srv.get_info()
        res = {"label": label, "K": K, "N": N, "mat_elem_bit_len": b, "queries_per_pass": nq, "ms_per_pass": ms, "queries_per_s": nq / (ms * 1e-3),
               "us_per_query": ms * 1e3 / nq, "packed_bytes_max_rank": gi["packed_bytes_max_rank"],
               "hbm_gbs_per_gpu": (gi["packed_bytes_max_rank"] + 4 * (gi["k_pitch"] if gi["respond_by_rows"] else K)) * nq / (ms * 1e-3) / 1e9,
               "respond_by_rows": gi["respond_by_rows"], "parity": par}
        if with_hint:
            res["setup"] = {"wall_s": wall, **{k: round(v, 6) for k, v in tmax.items()}, "gemm_kernel_ms": max(i["gemm_ms"] for i in infos),
                            "hint_gather_s": gi["hint_gather_s"], "hint_bytes": len(hint)}
        srv.close()
        del Ds, slices, head, o
        free()
        assert all(bool(v) for v in par.values()), (label, par)
        return res

    try:
        if n_gpus == 1 and (args.log2n, args.arity) != (18, 3):
            out["configs[1]"] = gemv_leg(18, 3, 64, True, "2^18 entries, 3-wise: setup hint GEMM + single-query respond on 1 B200")
        if (args.log2n, args.arity) != (20, 4):
            srv, _, Ds, plans, (b, K, N), wall = make_cluster_server(cp, torch, cluster, 20, 4, True, keep_d=True, batch_tc=1)
            leg = batched_leg(cp, torch, cluster, srv, plans, K, N, b, 64, 10, 64, Ds=Ds, label=f"2^20 entries, 4-wise: batched respond, 64-query int8-limb GEMM on {n_gpus} B200")
            leg128 = batched_leg(cp, torch, cluster, srv, plans, K, N, b, 128, 10, 65, label="the same with 128 queries (one full M tile)")
            leg.update({"K": K, "N": N, "mat_elem_bit_len": b, "with_128_queries": {k: leg128[k] for k in ("ms_per_batch", "queries_per_s", "issued_int8_tops_all_gpus")}})
            assert all(leg["parity"].values()), leg["parity"]
            out["configs[3]"] = leg
            srv.close()
            del Ds
            free()
        if n_gpus == 8 and (args.log2n, args.arity) != (22, 3):
            out["configs[4]"] = gemv_leg(22, 3, 32, not args.skip_hint, "2^22 entries, 3-wise: setup + respond sharded across 8 B200 (HBM-capacity sizing)")
    except AssertionError:
        raise
    except Exception as ex:  # a secondary leg must never cost the headline line
        out["error"] = repr(ex)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--arity", type=int, default=3, choices=[3, 4])
    ap.add_argument("--queries-per-step", type=int, default=256,
                    help="queries per step; 256 keeps the driver's 20-step timed region near one second on one GPU (a sustained, not a burst, figure)")
    ap.add_argument("--e2e-threads", type=int, default=0, help="concurrent callers of chpir_cluster_server_respond in the e2e leg (0 = 32 per GPU, at most 256)")
    ap.add_argument("--skip-hint", action="store_true", help="make D resident only (no A expansion / hint GEMM) -- development shortcut")
    ap.add_argument("--a-expand", default="auto", choices=["auto", "device", "host"],
                    help="where Server::setup walks the TurboSHAKE128 chain of A: the library's default (host core, pipelined), GPU warp, or host core")
    ap.add_argument("--no-e2e-setup", dest="e2e_setup", action="store_false", help="skip the Server::setup(seed, db) measurement from raw keys/values")
    ap.add_argument("--no-configs", dest="configs", action="store_false", help="skip the legs for the other BASELINE.json configurations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-column-cut", action="store_true", help="skip the column-cut comparison pass (N > 1)")
    ap.add_argument("--no-batch-tc", action="store_true", help="skip the tensor-core batched respond measurement")
    ap.add_argument("--batch-queries", type=int, default=128)
    ap.add_argument("--ref-queries-per-step", type=int, default=8, help="--impl reference: queries of each step the CPU arm answers (a bounded sample of the step)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library prints to fd 1 (e.g. NCCL's version banner) goes to stderr instead
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
