#!/usr/bin/env python
"""bench.py — MPPI update throughput / latency on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl reference]
  torchrun ... bench.py --gpus N ...        (one rank per GPU, NCCL)

A "step" is one complete MPPI update (ControllerBase::next): fresh Philox noise, rollout, cost,
softmin weights, weighted noise sum, sequence update and shift (plus the cross-rank exchange for
N > 1).  Default workload: BASELINE config 3 — point_mass3d, K = 1,048,576 samples, T = 100 — the
configuration the metric's target is quoted on; it fits one GPU, and at N GPUs the SAME total K is
sharded K/N per rank (strong scaling), so N = 8 is exactly config 3 as written.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: roofline (dominant kernel, effective
HBM GB/s on the algorithmic bytes), roofline_injected (the HBM-bound injected-noise kernel),
cpu_baseline (graph-faithful CPU port timed on this box), latency percentiles.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, K, T, s, a, n_controllers)
    "cfg1": ("point_mass1d K=1024 T=20 (BASELINE config 1)", 1024, 20, 2, 1, 1),
    "cfg2": ("point_mass2d K=65536 T=50 (BASELINE config 2)", 65536, 50, 4, 2, 1),
    "cfg3": ("point_mass3d K=1048576 T=100 (BASELINE config 3)", 1048576, 100, 6, 3, 1),
    "cfg5": ("4096 x point_mass2d K=1024 T=30 (BASELINE config 5)", 1024, 30, 4, 2, 4096),
    "cfg4": ("learned MLP 9->128->128->6 on point_mass3d state, K=262144 T=50, bf16 tcgen05 (BASELINE config 4)",
             262144, 50, 6, 3, 1),
    # not a BASELINE config: the AUV (Fossen) model of SURVEY.md section 8f row N4, Heun integrator, StaticCost, Python-twin
    # action cost; parameters = the "full" set of tests/golden/auv_fixtures.npz
    "auv": ("AUV Fossen model rk2 (s=13, a=6) K=262144 T=50 [next-row workload, not a BASELINE config]", 262144, 50, 13, 6, 1),
}
MLP_WORKLOADS = {"cfg4"}
AUV_WORKLOADS = {"auv"}


def auv_params():
    d = np.load(os.path.join(ROOT, "tests", "golden", "auv_fixtures.npz"))
    return json.loads(bytes(d["params_json"]).decode())["full"]


def cpu_auv_throughput(K_full, T, budget_s=15.0):
    """CPU arm of the AUV workload: the OpenMP C restatement (oracle/, kind "port") on a bounded sample of K."""
    from oracle import Oracle, pyoracle                     # bench.py's cpu_baseline leg may use oracle/
    orc = Oracle("f32")
    prm = auv_params()
    rng = np.random.default_rng(1)
    sigma = 80.0 * np.eye(6)
    x = np.zeros(13); x[6] = 1.0
    U = np.zeros((T, 6))
    k = 4096
    eps = (80.0 * rng.standard_normal((k, T, 6))).astype(np.float32)
    t0 = time.perf_counter()
    orc.mppi_update_auv(prm, 0.1, 2, 1.0, sigma, x, np.ones(13), x, U, eps)
    per = (time.perf_counter() - t0) / k
    k = int(max(4096, min(K_full, budget_s / 4 / max(per, 1e-9))))
    eps = (80.0 * rng.standard_normal((k, T, 6))).astype(np.float32)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        orc.mppi_update_auv(prm, 0.1, 2, 1.0, sigma, x, np.ones(13), x, U, eps)
        times.append(time.perf_counter() - t0)
    return dict(value=k * T * 3 / sum(times), threads=pyoracle.num_threads(),
                sample=f"K={k} of {K_full} samples per update (T={T}), 3 updates, OpenMP C restatement of the Python controller "
                       f"with AUVModel rk2 (noise generation not included)")
MLP_FLOPS_PER_SAMPLE_STEP = 2 * (9 * 128 + 128 * 128 + 128 * 6)      # 36 608
METRIC = "mppi_sample_steps_per_sec"
UNIT = "sample-steps/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst cuBLAS 8192^3; kernel timed alone)"
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


def glorot_mlp(s, a, H=128, seed=4):
    """default_rng(4), Glorot-uniform like Keras Dense, biases 0, unit normalisation."""
    rng = np.random.default_rng(seed)

    def g(i, o):
        lim = np.sqrt(6.0 / (i + o))
        return rng.uniform(-lim, lim, (i, o)).astype(np.float32)

    return dict(W1=g(s + a, H), b1=np.zeros(H, np.float32), W2=g(H, H), b2=np.zeros(H, np.float32),
                W3=g(H, s), b3=np.zeros(s, np.float32))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        if not sm:
            return None
        hi = [v for v in sm if v >= 0.5 * max(sm)]      # samples under load
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload][kernel]
    except (OSError, KeyError, ValueError):
        return None


def bytes_alg(K, T, a, n_ctrl=1):
    """eps read once (4*a B per sample-step) + per-sample cost written."""
    return n_ctrl * (4 * a * K * T + 4 * K)


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU implementation stands in as the graph-faithful torch-CPU port
# (oracle/graph_oracle.py; TensorFlow C++ r2.1 cannot be installed offline).
# ---------------------------------------------------------------------------------------------------
def cpu_port_throughput(K_full, T, s, a, steps, warmup, budget_s=20.0):
    import torch
    from oracle.graph_oracle import GraphOracle     # bench.py's cpu_baseline leg may use oracle/
    threads = torch.get_num_threads()
    gen = torch.Generator().manual_seed(1)
    goal = np.tile([1.0, 0.0], a)

    def make(k):
        return GraphOracle(k, T, 0.1, 1.0, s, a, 1.0, 0.25 * np.eye(a), goal, np.ones(s), dtype=torch.float32)

    x = torch.zeros(s)
    U = torch.zeros(T, a)
    probe_k = min(K_full, 2048)
    g = make(probe_k)
    g.next_generating(x, U, gen)
    t0 = time.perf_counter()
    g.next_generating(x, U, gen)
    per_sample = (time.perf_counter() - t0) / probe_k
    k = int(min(K_full, max(probe_k, budget_s / max(per_sample, 1e-9) / max(steps + warmup, 1))))
    k = max(32, (k // 32) * 32)
    g = make(k)
    for _ in range(warmup):
        g.next_generating(x, U, gen)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        g.next_generating(x, U, gen)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return dict(value=k * T * steps / total, k=k, threads=threads, ms_per_step=1e3 * total / steps,
                sample=f"K={k} of {K_full} samples per update (T={T}, a={a}), {steps} updates, "
                       f"torch-CPU fp32 op-for-op graph port, noise generated inside the step")


def run_reference(args, rank, world):
    if rank != 0:
        return
    desc, K, T, s, a, n_ctrl = WORKLOADS[args.workload]
    steps = max(1, min(args.steps, 5))
    warm = 3                                   # W >= 3 warm-up steps, each a bounded CPU sample
    if args.workload in AUV_WORKLOADS:
        r = cpu_auv_throughput(K, T, budget_s=60.0)
        r["ms_per_step"] = None
    else:
        r = cpu_port_throughput(K, T, s, a, steps, warm, budget_s=60.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "K": K, "T": T, "s_dim": s, "a_dim": a, "n_controllers": n_ctrl},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def main():
    # keep stdout clean for the ONE JSON line: libraries (NCCL version banner, ...) write to fd 1
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl", "torch"],
                    help="N>1: fused in-kernel exchange over peer memory (falls back to nccl if the mailboxes cannot be "
                         "mapped), in-library ncclAllGather, or torch.distributed all_gather on external buffers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-injected", action="store_true")
    ap.add_argument("--k-override", type=int, default=0, help="developer knob: replace the workload's K (not a bench line)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from mppi_tf_b200 import ControllerBase, comm_unique_id

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    desc, K, T, s, a, n_ctrl = WORKLOADS[args.workload]
    if args.k_override:
        K, desc = args.k_override, desc + f" [K overridden to {args.k_override}: developer run, not the BASELINE config]"
    if n_ctrl > 1:
        # independent controllers: partition the controllers across ranks, no exchange at all
        n_local = n_ctrl // world
        k_rank, k_world, k_rankid = K, 1, 0
    else:
        n_local = 1
        k_rank, k_world, k_rankid = K, world, rank
    is_auv = args.workload in AUV_WORKLOADS
    sigma = (80.0 if is_auv else 0.25) * np.eye(a, dtype=np.float32)
    rng = np.random.default_rng(5)
    goal = None
    if n_ctrl > 1:
        goal = rng.uniform(-1, 1, (n_ctrl, s)).astype(np.float32)[rank * n_local:(rank + 1) * n_local]
    # a non-default torch stream: the library launches on it, and torch.cuda.Event records on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    ctrl = ControllerBase(k_rank, T, 0.1, 1.0, s, a, lam=1.0, sigma=sigma, goal=goal, seed=1, device=local_rank,
                          rank=k_rankid, world=k_world, n_controllers=n_local,
                          goal_per_controller=(n_ctrl > 1), stream=stream, model=("auv" if is_auv else "point_mass"))
    if is_auv:
        ctrl.setAuvModel(auv_params(), rk=2)
        ctrl.setActionCost("python", gamma=1.0, upsilon=1.0)
    is_mlp = args.workload in MLP_WORKLOADS
    if is_mlp:
        ctrl.setMlp(glorot_mlp(s, a))
    exchange = k_world > 1
    if exchange and args.exchange == "peer":
        # fused exchange: all-gather the CUDA IPC handles of the mailboxes once, then no collective call at all
        ok = True
        try:
            handles = [None] * world
            dist.all_gather_object(handles, ctrl.peerHandle())
            ctrl.peerAttach(handles)
        except Exception as e:                      # noqa: BLE001  (IPC not permitted / no peer access)
            ok = False
            print(f"[bench] rank {rank}: peer exchange unavailable ({e}); using nccl", file=sys.stderr)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if ok:
                raise SystemExit("peer exchange attached on some ranks only")
            args.exchange = "nccl"
    if exchange:
        if args.exchange == "peer":
            pass
        elif args.exchange == "nccl":
            uid = [comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctrl.commInit(uid[0])
        else:
            stride = ctrl.exchangeStride()
            send = torch.zeros(stride, device=dev)
            recv = torch.zeros(world * stride, device=dev)
            ctrl.setExchangeBuffers(send.data_ptr(), recv.data_ptr())

    x = np.zeros((n_local, s), np.float32) if n_ctrl == 1 else \
        rng.uniform(-1, 1, (n_ctrl, s)).astype(np.float32)[rank * n_local:(rank + 1) * n_local]
    if is_auv:
        x[:, 6] = 1.0                               # identity attitude (a zero quaternion is not a state)
        x[:, 0] = 1.0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def one_update(eps_ptr=None):
        ctrl.enqueueUpdate(eps_ptr)
        if exchange and args.exchange != "peer":             # peer: the update kernel exchanges and finishes itself
            if args.exchange == "nccl":
                ctrl.enqueueExchange()                       # in-library ncclAllGather on the same stream
            else:
                dist.all_gather_into_tensor(recv, send)
            ctrl.enqueueFinish()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctrl.setState(x)

    # ---- device-timed loop: value ------------------------------------------------------------------
    if sampler:
        sampler.start()                    # runs through warm-up and both timed regions (short regions at N = 8)
        time.sleep(0.15)                   # nvidia-smi needs a moment before its first row
    for _ in range(args.warmup):
        one_update()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (untimed)
        ev[i][0].record()
        one_update()
        ev[i][1].record()
    barrier()
    per_step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = sum(per_step_ms)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    units_per_step = K * T * n_ctrl
    value = units_per_step * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API (host buffers, synchronous next()) -------------------------
    def api_next():
        if exchange and args.exchange == "torch":      # caller-side all-gather between the async halves
            ctrl.setState(x)
            one_update()
            return ctrl.fetchAction()
        return ctrl.next(x)                             # the public synchronous call (in-library exchange)

    act = None
    for _ in range(args.warmup):
        act = api_next()
    barrier()
    lat = []
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        act = api_next()
        lat.append(time.perf_counter() - t0)
    e2e_s = time.perf_counter() - t_all0
    barrier()
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = units_per_step * args.steps / e2e_s
    lat_ms = sorted(1e3 * v for v in lat)
    pct = lambda q: lat_ms[min(len(lat_ms) - 1, int(q * len(lat_ms)))]

    # ---- HBM-bound injected-noise kernel (N = 1 only; inputs resident in HBM) -------------------------
    inj = None
    if world == 1 and not args.no_injected and not is_mlp:
        n_eps = n_local * K * T * a
        g = torch.Generator(device=dev).manual_seed(1234)
        eps = torch.randn(n_eps, device=dev, generator=g) * (80.0 if is_auv else 0.25)
        isteps = max(10, min(args.steps, 50))
        for _ in range(3):
            one_update(eps.data_ptr())
        torch.cuda.synchronize()
        iev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(isteps)]
        for i in range(isteps):
            flush.zero_()
            iev[i][0].record()
            one_update(eps.data_ptr())
            iev[i][1].record()
        torch.cuda.synchronize()
        ims = [e0.elapsed_time(e1) for e0, e1 in iev]
        inj = statistics.mean(ims)
        del eps
    clocks = sampler.stop() if sampler else None
    # fraction of this rank's samples whose fp32 softmin weight is non-zero after the last update: the weighted
    # noise sum revisits only those (an exact optimisation: 0 * z adds nothing), so the number is part of the workload
    c_last = np.asarray(ctrl.getCosts(), np.float64).reshape(n_local, -1)
    nonzero_frac = float(np.mean((c_last - c_last.min(1, keepdims=True)) * 1.4426950408889634 / 1.0 < 126.0))

    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_step = total_ms / args.steps
        kernel_ms = statistics.mean(per_step_ms)
        b_alg = bytes_alg(K // k_world, T, a, n_local)          # per launch (this rank's shard)
        achieved = b_alg / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if n_ctrl == 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "K": K, "T": T, "s_dim": s, "a_dim": a, "n_controllers": n_ctrl,
                       "mode": "philox (fresh noise regenerated in registers every update)",
                       "sharding": f"K/{k_world} samples per rank" if n_ctrl == 1 else f"{n_local} controllers per rank",
                       "exchange": (args.exchange if exchange else "none"),
                       "l2": "flushed between timed iterations (256 MiB write, untimed)",
                       "nonzero_weight_frac": round(nonzero_frac, 4)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * s * n_local),
                    "d2h_bytes_per_step": int(4 * a * n_local),
                    "latency_ms": {"p10": pct(0.10), "p50": pct(0.50), "p90": pct(0.90)}},
            "gpu_launches": args.steps * (2 if (exchange and args.exchange != "peer") else 1),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "rollout_philox_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload, "rollout_philox_kernel") if world == 1 and not args.k_override else None,
                         "peak_source": peak_src,
                         "note": "effective GB/s on the algorithmic bytes 4*a*K*T + 4*K; the Philox kernel moves "
                                 "almost no HBM bytes and runs at the scheduler-dispatch bound of its instruction mix (DESIGN.md 3.1, profiles/)"},
        }
        if is_auv:
            line["config"]["mode"] = "philox noise + Fossen dynamics (Heun), fp32"
            line["roofline"]["kernel"] = "rollout_auv_kernel"
            line["roofline"]["traffic"] = None
            line["roofline"]["note"] = ("effective GB/s on the algorithmic bytes 4*a*K*T + 4*K; the kernel is bound by its fp32 FFMA "
                                        "chains (DESIGN.md 3.5: 70 % issue-active, 45 % of the FFMA peak)")
        if is_mlp:
            tpeak, tsrc = measured_tensor_peak()
            tf = MLP_FLOPS_PER_SAMPLE_STEP * (K // k_world) * T / (kernel_ms * 1e-3) / 1e12
            line["dtype"] = "bf16"
            line["config"]["mode"] = "philox noise + bf16 tcgen05 MLP rollout (fp32 state and accumulation)"
            line["roofline"] = {"bound": "tensor", "kernel": "rollout_mlp_kernel", "achieved": tf, "peak": tpeak,
                                "unit": "TFLOP/s", "frac": tf / tpeak,
                                "traffic": measured_traffic(args.workload, "rollout_mlp_kernel") if world == 1 and not args.k_override else None,
                                "peak_source": tsrc,
                                "note": "algorithmic flops 2*(9*128+128*128+128*6) = 36608 per sample-step (unpadded)"}
        if inj is not None:
            ia = bytes_alg(K, T, a, n_local) / (inj * 1e-3) / 1e9
            line["roofline_injected"] = {"bound": "hbm", "kernel": "rollout_auv_kernel<injected>" if is_auv else "rollout_injected_kernel", "achieved": ia,
                                         "peak": peak, "unit": "GB/s", "frac": ia / peak,
                                         "traffic": measured_traffic(args.workload, "rollout_injected_kernel") if not args.k_override else None,
                                         "ms_per_launch": inj,
                                         "inputs": "eps resident in HBM" + (" (larger than L2)" if 4 * n_eps > 126e6 else " (L2 flushed)")}
        if not args.no_cpu_baseline and world == 1 and is_auv:
            r = cpu_auv_throughput(K, T)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"]}
        elif not args.no_cpu_baseline and world == 1 and not is_mlp:
            r = cpu_port_throughput(K, T, s, a, steps=3, warmup=1, budget_s=15.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    ctrl.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
