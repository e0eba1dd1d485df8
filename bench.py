#!/usr/bin/env python
"""bench.py -- batched MPC step solves/s on B200 (BASELINE.json metric) and the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (nearest index -> reference sampling -> rollout -> linearise/condense ->
QP solve -> outputs) over one batch of synthetic instances: config 2 of BASELINE.json, 4096 ego instances x
20-step horizon per GPU (weak scaling: every rank owns its own seeded batch; the only communication is one
all-gather of the packed per-instance result record per step, as the north star describes).

Timed region: CUDA events on the launching stream around each step, L2 flushed (256 MiB write) and the
in-place warm-start buffers restored before every step outside the events; sum over K steps, max over ranks.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "av-simulation-at-intersections_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "batched MPC step solves/sec"
UNIT = "solves/s"


def flops_per_solve(T: int, iters: float) -> float:
    """Algorithmic flop count of the implemented algorithm."""
    n = 2 * T
    f_prep = 64.0 * T
    f_cond = 45.0 * n * (n + 1) / 2.0 + 70.0 * (T + 1) + 12.0 * T * (T + 1)   # O(1) entries from 23 suffix moments
    f_iter = n ** 3 / 3.0 + 7.0 * n * n + 240.0 * T
    return f_prep + f_cond + iters * f_iter


def bytes_per_solve(T: int) -> float:
    return 8.0 * (15 * T + 23)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference(w, sample: int, pool):
    """The CPU controller (oracle port of main/lib/mpc.py, numpy float64) on `sample` instances of the workload,
    one instance per call as the scenarios call it, spread over the worker processes of `pool` (started by the
    caller, outside the timed region)."""
    from helpers import oracle_batch
    t0 = time.perf_counter()
    res = oracle_batch(w, list(range(sample)), pool=pool)
    dt = time.perf_counter() - t0
    assert all(r.status == 0 for r in res)
    return sample / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  cvxpy+ECOS cannot be installed
    offline and the reference is pure Python (nothing to compile into oracle/_ref), so this is the oracle port,
    on all host cores, on the same config."""
    if rank != 0:
        return
    from junction_mpc import synth
    from helpers import make_pool
    w = synth.make_workload(args.config, B=args.ref_sample)
    cores = os.cpu_count() or 1
    times = []
    with make_pool(cores) as pool:
        for k in range(args.warmup + args.steps):
            rate, dt = cpu_reference(w, args.ref_sample, pool)
            if k >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = args.ref_sample / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config {args.config}: {w['name']}, {args.ref_sample}-instance sample per step of the "
                               f"4096 x T={w['T']} batch", "T": w["T"], "instances_per_step": args.ref_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.ref_sample} instances of config {args.config} per step, one solve per call, "
                                   f"multiprocessing over {cores} processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: the config's own size)")
    ap.add_argument("--ref-sample", type=int, default=256)
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    if rank == 0 or not os.path.exists(os.path.join(g.PKG, "junction_mpc", "libjmpc.so")):
        g.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()

    from junction_mpc import synth
    from junction_mpc.batched import BatchedMPC
    from junction_mpc import _cabi
    from junction_mpc.distributed import allgather_records

    # every rank owns its own batch (weak scaling); rank r draws from seed stream config*1000 + r
    w = synth.make_workload(args.config, B=args.batch or None, seed_offset=rank)
    B, T = w["B"], w["T"]
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=B, device=local)
    f64, i32 = torch.float64, torch.int32
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
    state, clen = t(w["state"], f64), t(w["course_len"], i32)
    tgt0, oa0, od0 = t(w["target_ind"], i32), t(w["oa"], f64), t(w["od"], f64)
    tgt, oa, od = tgt0.clone(), oa0.clone(), od0.clone()
    out = mpc.alloc_outputs(B)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    # multi-GPU gather of the result records: fused into the kernel epilogue over NVLink peer memory when symmetric
    # memory is available, otherwise one NCCL all-gather straight from the kernel's record buffer
    fused, gather_kind = None, "none (single GPU)"
    if world > 1:
        gather_kind = "nccl all_gather_into_tensor of the record buffer"
        if os.environ.get("JMPC_NO_FUSED_GATHER") is None:
            try:
                from junction_mpc.distributed import FusedRecordGather
                fused = FusedRecordGather(mpc, B)
                gather_kind = "fused: step-kernel epilogue stores into every peer's table (NVLink symmetric memory) + signal-pad barrier"
            except Exception as exc:            # noqa: BLE001
                if rank == 0:
                    print(f"[bench] fused gather unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
                fused = None

    def one_step(e0=None, e1=None, cold=True):
        flush.zero_()
        tgt.copy_(tgt0); oa.copy_(oa0); od.copy_(od0)
        if cold:
            # The bench re-solves the same batch every step.  The engine's default schedule orders the work queue
            # by the iteration counts of the previous step (meant for closed loops); on a repeated batch those
            # would be exact foreknowledge, so the headline forgets them before every step: the queue is ordered by
            # the a-priori key computed from this step's inputs only.
            mpc.reset_schedule_hints()
        if e0 is not None:
            e0.record()
        mpc.step(state, tgt, oa, od, out, course_len=clen)
        if fused is not None:
            fused.finish()                     # cross-GPU barrier; the records are already in every peer's table
        elif world > 1:
            allgather_records(out.record)      # [world * B, 8] on every rank, straight from the kernel's buffer
        if e1 is not None:
            e1.record()

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    if fused is not None:                      # the fused table must equal what the NCCL all-gather delivers
        ref = allgather_records(out.record)
        torch.cuda.synchronize()
        same = torch.equal(torch.nan_to_num(fused.table, nan=-1.0), torch.nan_to_num(ref, nan=-1.0))
        assert same, "fused record gather differs from the NCCL all-gather"
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = mpc.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for e0, e1 in evs:
        one_step(e0, e1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.finish()
    launches = mpc.launch_count - launches0
    per_step = np.array([e0.elapsed_time(e1) for e0, e1 in evs])            # ms
    total_ms = torch.tensor([per_step.sum()], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = world * B / (ms_per_step * 1e-3)
    status = out.status.cpu().numpy()
    iters = out.iters.cpu().numpy()
    assert (status == 0).all(), f"{(status != 0).sum()} instances not solved"

    # secondary figure: the same steps with the previous step's iteration counts as scheduling keys (what a closed
    # loop that re-solves slowly changing instances sees; on this repeated batch the keys are exact)
    hint_steps = max(3, min(args.steps, 20))
    one_step(cold=False)
    evs2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(hint_steps)]
    for e0, e1 in evs2:
        one_step(e0, e1, cold=False)
    torch.cuda.synchronize()
    hint_ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs2) / hint_steps], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(hint_ms, op=dist.ReduceOp.MAX)
    hint_ms = float(hint_ms.item())

    # end to end through the host API: pinned staging, H2D of the step's inputs and D2H of its results inside
    e2e_steps = max(3, min(args.steps, 20))
    h2d = w["state"].nbytes + w["oa"].nbytes + w["od"].nbytes + 4 * B * 2
    T1 = T + 1
    d2h = 8 * B * (2 * T + 4 * T1 + 4 * T1 + 1 + _cabi.RECORD_LEN) + 4 * B * 3
    hout = mpc.host_outputs(B)                      # page-locked result arrays, reused every step
    hin = {}
    for key in ("state", "target_ind", "oa", "od", "course_len"):      # the step's inputs, in page-locked host memory
        hin[key] = mpc.pinned_empty(w[key].shape, w[key].dtype)
        hin[key][...] = w[key]
    for _ in range(2):
        mpc.step_host(hin["state"], hin["target_ind"], hin["oa"], hin["od"], course_len=hin["course_len"], out=hout)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mpc.reset_schedule_hints()              # cold schedule, as in the device-timed region
        ho = mpc.step_host(hin["state"], hin["target_ind"], hin["oa"], hin["od"], course_len=hin["course_len"], out=hout)
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (float(e2e_ms.item()) * 1e-3)
    assert (ho.status == 0).all()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # latency of one controller call for a single ego (what a scenario runner sees per timestep)
    one = BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=1, device=local)
    lat = []
    for k in range(60):
        t0 = time.perf_counter()
        one.step_host(w["state"][:1], w["target_ind"][:1], w["oa"][:1], w["od"][:1], course_len=w["course_len"][:1])
        if k >= 10:
            lat.append((time.perf_counter() - t0) * 1e3)
    one.close()

    # roofline of the (single) kernel of the step: CUDA-core FP64 FMA bound
    fp64_peak, fp32_peak = mpc.measure_fma_peak()
    kernel_ms = float(np.mean(per_step))
    mean_iters = float(iters.mean())
    fl = flops_per_solve(T, mean_iters) * B
    achieved = fl / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1k_traffic.json")) as f:
            tr = json.load(f)
        if B == 4096 and T == 20:
            traffic = tr["dram_bytes_per_launch"]          # from the committed ncu --set full capture, per launch
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = bytes_per_solve(T) * B / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"config {args.config}: {w['name']}, {B} synthetic ego instances x T={T} per GPU "
                               f"(BASELINE.json configs[{args.config - 1}])", "instances_per_gpu": B, "T": T,
                   "l2": "flushed with a 256 MiB write before every timed step",
                   "solver": "condensed QP, Mehrotra predictor-corrector interior point, fp64",
                   "mean_solver_iters": mean_iters,
                   "schedule": "longest-first work queue by an a-priori key from the step's inputs (speed-cap proximity); "
                               "previous-step iteration counts deliberately forgotten before every step", "max_solver_iters": int(iters.max()), "gather": gather_kind},
        "p50_ms": float(np.percentile(per_step, 50)), "p99_ms": float(np.percentile(per_step, 99)),
        "single_instance_step_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                                    "api": "step_host, B=1, host in / host out"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": float(e2e_ms.item()), "steps": e2e_steps,
                "api": "BatchedMPC.step_host(out=host_outputs) -> jmpc_step_host_io: numpy inputs and results in page-locked host "
                       "memory; inputs copied host->device with cudaMemcpyAsync, results stored by the step kernel "
                       "straight into the host arrays over PCIe (device mapping of the page-locked memory, no "
                       "device->host copy pass); wall clock around the call, results readable on the host when it returns"},
        "gpu_launches": int(launches),
        "with_history_hints": {"value": world * B / (hint_ms * 1e-3), "unit": UNIT, "ms_per_step": hint_ms, "steps": hint_steps,
                               "note": "same steps with the work queue ordered by the previous step's iteration counts "
                                       "(the engine's closed-loop default); exact foreknowledge on this repeated batch, "
                                       "so it is reported beside the headline, not as it"},
        "roofline": {"bound": "fp64_fma", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
                     "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1k_traffic.json)",
                     "peak_source": "jmpc_measure_fma_peak: register-resident FP64 FMA loop on all SMs, measured in this run "
                                    "(MEASURED_PEAKS.json holds no FP64 figure)",
                     "flops_per_solve": flops_per_solve(T, mean_iters), "kernel": "jmpc::mpc_step_kernel",
                     "kernel_ms": kernel_ms, "fp32_peak_tflops": fp32_peak,
                     "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                             "bytes_per_solve": bytes_per_solve(T),
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback of B200_PROFILING.md"}},
    }
    if world == 1 and not args.no_cpu:
        from helpers import make_pool
        cores = os.cpu_count() or 1
        with make_pool(cores) as pool:
            cpu_reference(w, min(64, B), pool)                     # warm the workers (imports, BLAS)
            rate, dt = cpu_reference(w, min(args.cpu_sample, B), pool)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {min(args.cpu_sample, B)} instances of the same batch, numpy float64 "
                                          f"oracle (restated main/lib/mpc.py, certified QP solve), one solve per call, "
                                          f"{cores} processes, {dt:.1f} s wall"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
