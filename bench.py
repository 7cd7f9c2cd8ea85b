#!/usr/bin/env python
"""bench.py — ADMM iterations/s of the cuADMM hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one sGS-ADMM iteration (src/solver.cu:469-799 of the reference): two A A^T y-solves, three
SpMV with A, two with A^T, the PSD projection of every block and the residual / sigma update.  Workload
at N=1: BASELINE.json configs[1], "~2,000 PSD blocks of size 6-60" = 2,000 blocks, n_k ~ U{6..60}
(numpy default_rng(0)), m = 700,000 chain-structured sparse constraints (cuadmm_b200/synthetic.py),
FP64, synthetic data.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nblk", type=int, default=2000)
    ap.add_argument("--con", type=int, default=700000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2b", choices=["c2b", "c3", "c4"],
                    help="c2b (default, BASELINE.json configs[1]); c3 = max-cut n=4000 (configs[2]); c4 = 10k mixed "
                         "blocks {10,50,200,800}, m=1e6 (configs[3], the multi-GPU scaling configuration)")
    return ap.parse_args()


def workload(args):
    from cuadmm_b200.synthetic import c2b_blocks, chain_sdp, maxcut_sdp, c4_blocks, random_sdp
    if args.workload == "c3":
        P = maxcut_sdp(4000, p=0.01, seed=0)
        return P, {"workload": "C3 max-cut SDP: one dense PSD block n=4000, G(n, 0.01) seed 0, m=n diagonal constraints, "
                               "sGS-ADMM iteration", "nblk": 1, "vec_len": int(P["vec_len"]), "con_num": int(P["con_num"]),
                   "nnz_A": int(len(P["vals"])), "l2": "one 128 MB dense block per n x n operand: exceeds the 126 MB L2"}
    if args.workload == "c4":
        blk = c4_blocks(seed=0)
        m = args.con if args.con != 700000 else 1000000
        P = random_sdp(blk, m, seed=0)
        return P, {"workload": "C4 synthetic multi-block SDP: 10,000 blocks {10: 6000, 50: 3000, 200: 900, 800: 100} shuffled "
                               "(seed 0), m=%d constraints with 1+Poisson(4) non-zeros at random svec positions of <= 2 blocks, sGS-ADMM iteration" % m, "nblk": int(len(blk)),
                   "vec_len": int(P["vec_len"]), "con_num": int(P["con_num"]), "nnz_A": int(len(P["vals"])),
                   "l2": "svec vectors of 434 MB each: far beyond the 126 MB L2"}
    blk = c2b_blocks(args.nblk, 6, 60, 0)
    P = chain_sdp(blk, args.con, seed=0)
    cfg = {"workload": "C2b synthetic moment-relaxation SDP: %d PSD blocks n~U{6..60} (seed 0), m=%d chain-structured "
                       "constraints, sGS-ADMM iteration" % (args.nblk, args.con),
           "nblk": int(args.nblk), "vec_len": int(P["vec_len"]), "con_num": int(P["con_num"]), "nnz_A": int(len(P["vals"])),
           "l2": "per-iteration working set (8 svec vectors + A + At + L ~ 190 MB) exceeds the 126 MB L2; no flush, "
                 "iterations run back to back exactly as in the solver"}
    return P, cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of the projection stage from the committed ncu --set full
    capture (profiles/ncu_traffic_r01.json); only valid for the workload it was taken on, else None"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")
    try:
        d = json.load(open(p))
        if args.workload == "c2b" and d["workload"]["nblk"] == args.nblk and d["workload"]["con"] == args.con and args.gpus == 1:
            st = d["projection_stage"]
            return int(st["dram_bytes_read"] + st["dram_bytes_write"])
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference's CPU path (LAPACK dsyevd thread pool + host sparse
# solve).  The only places bench.py may execute oracle/ (cpu_baseline and --impl reference).
# ---------------------------------------------------------------------------------------------
def cpu_reference_iterations(P, n_iters, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_np as onp
    blk = P["blk"]
    proj = lambda v: onp.project_svec_threads(blk, v, threads)
    t0 = time.time()
    o = onp.ADMMOracle(P["vec_len"], P["con_num"], P["col_ptrs"], P["row_ids"], P["vals"], P["b_idx"], P["b_val"],
                       P["C_idx"], P["C_val"], blk, project=proj)
    t_init = time.time() - t0
    o.solve(1, -1.0, 500, 50, 100, 1 << 30, 1.05)          # one untimed iteration
    t0 = time.time()
    o.solve(n_iters, -1.0, 500, 50, 100, 1 << 30, 1.05)
    dt = time.time() - t0
    return n_iters / dt, dt, t_init


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P, cfg = workload(args)
    cores = os.cpu_count() or 1
    threads = min(30, cores)          # the reference's default cpu_eig_thread_num (src/main.cu:11)
    steps = max(1, min(args.steps, 6))
    val, dt, t_init = cpu_reference_iterations(P, steps, threads)
    line = {"impl": "reference", "metric": "ADMM iterations/s", "value": val, "unit": "iter/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": "iter/s", "cores": threads, "kind": "port",
                             "sample": "%d sGS-ADMM iterations of the same workload after 1 untimed; LAPACK dsyevd "
                                       "(OpenBLAS, 1 thread per call) on a %d-thread pool + SuperLU solves "
                                       "(CHOLMOD-substitute); init %.1f s not counted" % (steps, threads, t_init)},
            "e2e": {"value": val, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host_cores": cores}
    # Baseline A (informational): the reference's own cuSOLVER projection stage (src/solver.cu:531-647,
    # unmodified reference sources in oracle/_ref) on GPU 0 of this box, same blocks, same input
    try:
        import ctypes as C
        ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcuadmm_ref.so"))
        ref.ref_proj_create.restype = C.c_void_p
        ref.ref_proj_run.restype = C.c_double
        from cuadmm_b200.synthetic import random_svec
        blk = np.ascontiguousarray(P["blk"], np.int32)
        x = random_svec(blk, seed=0)
        h = ref.ref_proj_create(blk.ctypes.data_as(C.POINTER(C.c_int)), len(blk), 15)
        ms = ref.ref_proj_run(C.c_void_p(h), x.ctypes.data_as(C.POINTER(C.c_double)), None, 2)
        ref.ref_proj_destroy(C.c_void_p(h))
        line["baseline_A_cusolver_projection_ms"] = ms
    except Exception as ex:
        line["baseline_A_cusolver_projection_ms"] = None
        line["baseline_A_note"] = "oracle/_ref not usable here: %s" % ex
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import cuadmm_b200 as cu
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if cu.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device — the CUDA extension has no CPU fallback")
    torch.cuda.set_device(local)
    os.environ["CUADMM_DEVICE"] = str(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    P, cfg = workload(args)
    s = cu.Solver(verbose=False)
    if world > 1:
        # one process per GPU: blocks sharded by eig cost, partial A x all-reduced over NCCL every iteration
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(cu.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        s.set_distributed(rank, world, bytes(idt.cpu().numpy().tolist()))
        cfg["parallelism"] = "blocks sharded over %d GPUs (LPT on eig cost), y-solve replicated, 3 NCCL all-reduces of m doubles per iteration" % world
    t0 = time.time()
    s.init(15, 30, P["vec_len"], P["con_num"], P["col_ptrs"], P["row_ids"], P["vals"], P["b_idx"], P["b_val"],
           P["C_idx"], P["C_val"], P["blk"], None, None, None, 1.0)
    t_init = time.time() - t0
    s.solve(1, -1.0, 500, 50, 100, 1 << 30, 1.05)          # sets the run parameters (1 iteration)
    s.run_iterations(max(args.warmup, 3), sgs=True)          # untimed warm-up
    launches0 = s.launches

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    r = s.run_iterations(args.steps, sgs=True)              # EXACTLY K timed steps, CUDA events on the solver stream
    barrier()
    ms_total = r["total_ms"]
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    launches = s.launches - launches0
    # stage breakdown (separate profiled pass: events around every y-solve and projection)
    prof = s.run_iterations(min(args.steps, 50), sgs=True, profile=True)
    nprof = min(args.steps, 50)
    proj_ms = prof["projection_ms"] / nprof

    # e2e: the same iteration through the public C ABI with HOST buffers — every step uploads the iterate
    # (X, y, S) from pinned host memory, runs one iteration (solve(max_iter=1, if_first=False), the
    # reference's warm-restart path src/solver.cu:385-409) and downloads X, y, S.
    n, m = P["vec_len"], P["con_num"]
    hX = torch.zeros(n, dtype=torch.float64).pin_memory(); hy = torch.zeros(m, dtype=torch.float64).pin_memory()
    hS = torch.zeros(n, dtype=torch.float64).pin_memory()
    aX, ay, aS = hX.numpy(), hy.numpy(), hS.numpy()
    s.get_into(aX, ay, aS)
    e2e_steps = max(3, min(args.steps, 30))
    for _ in range(2):
        s.set_XyS(aX, ay, aS, 1.0); s.solve(1, -1.0, 500, 50, 100, 1 << 30, 1.05, if_first=False); s.get_into(aX, ay, aS)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.set_XyS(aX, ay, aS, 1.0)
        s.solve(1, -1.0, 500, 50, 100, 1 << 30, 1.05, if_first=False)
        s.get_into(aX, ay, aS)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    ys = s.ysolve_stats()
    if dist is not None:
        # leave the job together: the solver's NCCL communicator first, then torch's process group
        s.close()
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    hbm, hbm_src = measured_peaks()
    value = args.steps / (ms_total / 1e3)
    alg_bytes = 56 * n                      # fused projection stage: reads Xb, X, Rd1, C; writes Xproj, S, SmC
    f_alg = float(sum((20.0 / 3.0) * float(b) ** 3 for b in P["blk"]))
    achieved = alg_bytes / (proj_ms * 1e-3) / 1e9
    line = {"metric": "ADMM iterations/s", "value": value, "unit": "iter/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "clocks": clocks,
            "e2e": {"value": e2e_steps / e2e_dt, "unit": "iter/s", "h2d_bytes_per_step": 8 * (2 * n + m),
                    "d2h_bytes_per_step": 8 * (2 * n + m), "steps": e2e_steps,
                    "how": "per step: cuadmm_solver_set_XyS (pinned host -> device), cuadmm_solver_solve(max_iter=1, "
                           "if_first=0), cuadmm_solver_get_X/y/S (device -> pinned host); wall clock"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "proj_jacobi_kernel (fused svec->smat, one-sided Jacobi eig, clamp, rebuild, smat->svec, S/SmC "
                                   "epilogue; all size classes of one projection stage)",
                         "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": ncu_traffic(args), "traffic_source": "profiles/ncu_traffic_r01.json (ncu --set full, "
                         "dram__bytes_read.sum + dram__bytes_write.sum summed over the stage's three launches)",
                         "peak_source": hbm_src, "alg_bytes_per_launch": alg_bytes,
                         "launch_ms": proj_ms,
                         "note": "latency/issue-bound Jacobi kernel, neither HBM- nor tensor-bound (SURVEY 8d): "
                                 "F_alg=(20/3)sum n^3 = %.3g flop -> %.3f TFLOP/s vs 35.5 TFLOP/s measured cuBLAS DGEMM"
                                 % (f_alg, f_alg / (proj_ms * 1e-3) / 1e12),
                         "fp64_frac_of_dgemm_peak": f_alg / (proj_ms * 1e-3) / 1e12 / 35.5},
            "stage_ms_per_iter": {"projection": proj_ms, "ysolve_x2": prof["ysolve_ms"] / nprof, "spmv_and_rest": prof["other_ms"] / nprof},
            "ysolve": ys, "init_s": t_init}
    if args.workload != "c2b":
        # C3 / C4: the projection is dominated by the large blocks (n > 168), i.e. by sym_gemm_kernel, the FP64
        # tensor-core (DMMA) product of the sign iteration: tensor-bound, F_alg = (20/3) n^3 per block
        line["roofline"] = {"kernel": "projection stage: sym_gemm_kernel (FP64 DMMA sign iteration, blocks n > 168) + proj_jacobi_kernel",
                            "bound": "tensor", "achieved": f_alg / world / (proj_ms * 1e-3) / 1e12, "peak": 35.5, "unit": "TFLOP/s",
                            "frac": f_alg / world / (proj_ms * 1e-3) / 1e12 / 35.5, "traffic": None,
                            "peak_source": "cuBLAS DGEMM 8192^3 measured on this pool's B200 (gpurun_out/probe.json, round 1); "
                                           "MEASURED_PEAKS.json holds no FP64 figure",
                            "alg_flop_per_launch": f_alg / world, "launch_ms": proj_ms,
                            "note": "the sign iteration executes ~35 symmetric products of n^3 flop each (~5x F_alg)"}
    if not args.no_cpu_baseline and args.gpus == 1:
        cores = os.cpu_count() or 1
        threads = min(30, cores)
        try:
            val, dt, ti = cpu_reference_iterations(P, 3, threads)
            line["cpu_baseline"] = {"value": val, "unit": "iter/s", "cores": threads, "kind": "port",
                                    "sample": "3 sGS-ADMM iterations of the same workload (after 1 untimed) with the oracle port: "
                                              "LAPACK dsyevd on a %d-thread pool + SuperLU solves (CHOLMOD-substitute); %.1f s" % (threads, dt)}
        except Exception as ex:   # the baseline is informational; never lose the GPU line over it
            line["cpu_baseline"] = {"value": None, "unit": "iter/s", "cores": threads, "kind": "port", "sample": "failed: %s" % ex}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
