#!/usr/bin/env python
"""bench.py — chunks/s built+interned ("perlin dunes" world of 32^3 chunks) on N B200s.

Contract (see the task brief): one JSON line on rank 0.
  step      = one pass of the hot path over the rank's world: interner reset + ONE apply_batches
              launch over 64x8x64 = 32 768 chunks of 32^3 voxels (u8), batches resident in HBM.
  value     = chunks/s, whole job (sum over ranks / max-over-ranks device time), weak scaling:
              every rank builds its own 64x8x64 world (X offset = rank * 64 chunks).
  e2e       = the same through the reference-facing C-ABI call on HOST batches: vx_apply_batches over
              n Batch + n VoxTree handles; bus traffic of the batches and D2H of roots/changed inside the
              timed region (the dense two-array slab entry is reported beside it).
  roofline  = frac: algorithmic bytes per launch (SURVEY §8d formula) / CUDA-event duration of the apply call,
              against MEASURED_PEAKS.json:hbm_gbs.  The headline input is 1 % dense and the kernels skip the values
              of untouched blocks, so the line also carries frac_compulsory (bytes this input really needs) and
              frac_dram (DRAM bytes of the call measured IN THIS RUN by an ncu replay of the same step,
              profiles/tools/ncu_traffic.py) — the honest figures.
  cpu_baseline / --impl reference = the CPU oracle (a C++ restatement of the reference's Rust
              apply_batch; the Rust reference cannot be built in this image) on the host cores.
Beside the headline: `headline_dense` (surface-and-below terrain, the dense variant), `others` (BASELINE config 2
and friends), `latency_single_chunk` (config 1), `strong` (ONE world split across the ranks, config 3), `d7`
(config 4's per-GPU share) and `global_dedup` (config 5) — the last three every time N > 1.
Inputs are synthetic (seeded integer value-noise height field); input size 1.34 GB per step is larger
than the 126 MB L2, so no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = (64, 8, 64)          # chunks along x, y, z (BASELINE.json config 3)
DEPTH = 5
BUDGET = 256 << 20          # VoxInterner memory budget per GPU (README quick-start value)
NODE_BYTES = 87             # 78 + sizeof(u8) payload + 8 B table slot per NEW node
CHUNK_BYTES = 2 * 4096 + 8 * 4096 + 8   # masks + values read once + root id written
WORKLOAD = "perlin_dunes_surface_only_64x8x64_d5_u8"
METRIC = "chunks/s built+interned (perlin 32^3)"
D7_CHUNKS_PER_GPU = 2048    # BASELINE config 4: 16 k chunks of 128^3 over 8 GPUs


def base_config(n):
    """Identical in both arms (the driver compares the two config objects)."""
    return {"workload": WORKLOAD, "chunks_per_step": int(n), "depth": DEPTH, "interner_budget_bytes": BUDGET}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        """Launch the sampler (before the warm-up, so it is already delivering when the timed region starts)."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def in_window(self):
        return sum(1 for t, r in self.rows if self.t0 <= t <= (self.t1 or time.monotonic()) and r and r[0].isdigit())

    def stop(self, window="timed region"):
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t))]
        sm = [int(r[0]) for r in rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


# stdout carries exactly ONE JSON line.  Native libraries (NCCL's version banner, driver notices) write to
# file descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON line
# goes to a private duplicate of the original stdout.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def log(msg: str):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def make_world(rank: int, variant="surface_only", grid=GRID, x_chunk_offset=None):
    from voxelis_b200 import workloads as wl
    return wl.terrain_world(grid, DEPTH, variant, wl.U8,
                            x_chunk_offset=rank * GRID[0] if x_chunk_offset is None else x_chunk_offset,
                            materials=3 if variant != "surface_only" else 1)


def cpu_baseline(masks, values, threads: int, target_s: float = 12.0):
    """Oracle (kind 'port') on the host cores: fresh-tree apply over the same batches, one private
    interner per thread, repeated until ~target_s of CPU work."""
    from oracle import oracle
    n = masks.shape[0]
    t, _ = oracle.time_apply_fresh(0, DEPTH, BUDGET, masks, values, threads)
    reps, total, chunks = 1, t, n
    while total < target_s and reps < 64:
        t, _ = oracle.time_apply_fresh(0, DEPTH, BUDGET, masks, values, threads)
        total += t
        chunks += n
        reps += 1
    return {"value": chunks / total, "unit": "chunks/s", "cores": threads, "kind": "port",
            "sample": f"{reps} x full {n}-chunk world ({total:.1f} s), fresh interner per repetition, "
                      f"{threads} thread(s) x private interner"}


def run_reference(args):
    """--impl reference: the reference's CPU apply_batch (oracle port) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    masks, values = make_world(0)
    n = masks.shape[0]
    from oracle import oracle
    for _ in range(max(args.warmup, 1)):
        oracle.time_apply_fresh(0, DEPTH, BUDGET, masks, values, threads)
    times = [oracle.time_apply_fresh(0, DEPTH, BUDGET, masks, values, threads)[0] for _ in range(args.steps)]
    total = sum(times)
    v = n * args.steps / total
    nonempty = int(np.count_nonzero(masks[:, :, 0].any(axis=1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "chunks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": base_config(n),
        "workload_detail": {"nonempty_chunks": nonempty},
        "value_nonempty": v * nonempty / n,
        "cpu_baseline": {"value": v, "unit": "chunks/s", "cores": threads, "kind": "port",
                         "sample": f"full {n}-chunk world per step, {threads} threads x private interner; "
                                   "C++ oracle port (Rust reference not buildable here)"},
        "e2e": {"value": v, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def measure_dram_traffic(workload: str, timeout_s: int = 420):
    """DRAM bytes of ONE apply call of `workload`, measured now on this GPU: profiles/tools/ncu_traffic.py under ncu
    (two metrics, one replay pass).  Returns the parsed dict or {"error": ...}."""
    tool = os.path.join(ROOT, "profiles", "tools", "ncu_traffic.py")
    out = tempfile.NamedTemporaryFile(prefix="vx_ncu_", suffix=".csv", delete=False).name
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--print-units", "base",
           "--clock-control", "none", "--csv", "--log-file", out, sys.executable, tool, workload]
    try:
        env = dict(os.environ)
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        if res.returncode != 0:
            return {"error": f"ncu rc {res.returncode}: {(res.stderr or res.stdout)[-300:]}"}
        sys.path.insert(0, os.path.dirname(tool))
        import ncu_traffic
        r = ncu_traffic.parse(out)
        if not r["launches"]:
            return {"error": "ncu produced no launch rows"}
        return r
    except Exception as e:  # ncu missing, counters not permitted, timeout: say so, never fake a number
        return {"error": f"{type(e).__name__}: {e}"[:300]}
    finally:
        try:
            os.unlink(out)
        except OSError:
            pass


def around_the_path(vx, dev, local_rank, peak):
    """Headline world generated IN device memory (vx_terrain_*_device), built, and unfolded into the greedy mesher's
    occupancy planes (vx_occupancy_masks, 2x2x2 chunks per 64^3 builder, device outputs).  CUDA events on the
    interner's stream; occupancy is a synchronous call (host placement of the roots included), so it is wall time."""
    import ctypes as C
    import torch
    from voxelis_b200 import workloads as wl
    gx, gy, gz = GRID
    N, B, n = 1 << DEPTH, 8 ** (DEPTH - 1), GRID[0] * GRID[1] * GRID[2]
    it = vx.VoxInterner.with_memory_budget(BUDGET, vx.U8, local_rank)
    st = torch.cuda.ExternalStream(it.stream, device=dev)
    h = torch.empty((gx * N, gz * N), dtype=torch.int32, device=dev)
    m = torch.empty((n, B, 2), dtype=torch.uint8, device=dev)
    v = torch.empty((n, B, 8), dtype=torch.uint8, device=dev)
    roots = torch.zeros(n, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def gen():
        it.terrain_heights_device(gx * N, gz * N, h.data_ptr(), wl.SEED_BASE, gy * N)
        it.terrain_batches_device(DEPTH, GRID, h.data_ptr(), m.data_ptr(), v.data_ptr(), True, 1)

    def build():
        it.reset_async()
        it.apply_batches_device(DEPTH, n, m.data_ptr(), v.data_ptr(), roots.data_ptr())

    def timed(f, reps=20):
        for _ in range(3):
            f()
        it.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            f()
        b.record(st)
        it.sync()
        return a.elapsed_time(b) / reps

    ms_gen = timed(gen)
    ms_both = timed(lambda: (gen(), build()))
    hroots = roots.cpu().numpy().astype(np.uint64)
    idx = np.arange(n)
    cx, cy, cz = idx // (gy * gz), (idx // gz) % gy, idx % gz
    bo = np.ascontiguousarray(((cx // 2) * (gy // 2) + cy // 2) * (gz // 2) + cz // 2, np.uint32)
    offs = np.ascontiguousarray(np.stack([(cx % 2) * N, (cy % 2) * N, (cz % 2) * N], 1), np.uint32)
    nb, M = n // 8, 2
    outs = [torch.empty(s, dtype=t, device=dev) for s, t in (((nb, 3 * 4096), torch.int64), ((nb, 6), torch.int64),
            ((nb,), torch.int32), ((nb, M), torch.int64), ((nb, M), torch.int64), ((nb, M, 3 * 4096), torch.int64))]
    torch.cuda.synchronize()
    p = lambda a: a.ctypes.data_as(C.c_void_p)

    def occ():
        rc = vx.lib().vx_occupancy_masks(it.h, DEPTH, 0, n, p(hroots), p(offs), p(bo), nb, M,
                                         *[C.c_void_p(t.data_ptr()) for t in outs])
        if rc != 0:
            raise RuntimeError(vx.lib().vx_last_error().decode())
    occ()
    t0 = time.perf_counter()
    for _ in range(5):
        occ()
    ms_occ = (time.perf_counter() - t0) / 5 * 1e3
    gen_bytes = m.numel() + v.numel()
    occ_bytes = int(nb * (3 * 4096 * 8 + 48) + int(outs[2].sum().item()) * 3 * 4096 * 8)
    return {"world": "perlin 64x8x64 d5 u8 surface_only, generated on the device",
            "generate_ms": ms_gen, "generate_gbs_written": gen_bytes / ms_gen / 1e6, "generate_frac_hbm": gen_bytes / ms_gen / 1e6 / peak,
            "generate_plus_build_ms": ms_both, "generate_plus_build_chunks_per_s": n / ms_both * 1e3,
            "occupancy_builders": nb, "occupancy_ms_wall": ms_occ, "occupancy_chunks_per_s": n / ms_occ * 1e3,
            "occupancy_gbs_written": occ_bytes / ms_occ / 1e6, "occupancy_frac_hbm": occ_bytes / ms_occ / 1e6 / peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-others", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the in-run ncu replay that measures DRAM bytes")
    ap.add_argument("--traffic-all", action="store_true", help="ncu replay for the dense secondary workloads too")
    ap.add_argument("--no-d7", action="store_true", help="skip the MaxDepth-7 leg (BASELINE config 4)")
    ap.add_argument("--no-dedup", action="store_true", help="skip the global-dedup leg when N > 1")
    ap.add_argument("--dedup", action="store_true", help="run the global-dedup leg at N = 1 too (it is on for N > 1)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="'strong' only runs the one-world legs faster at N = 1; both values are always in the line for N > 1")
    ap.add_argument("--workload", default="perlin", help="profiling only: time another workload in the main loop "
                    "(checkerboard | sum | sum_per_chunk | random255 | below); the headline is 'perlin'")
    ap.add_argument("--secondary-multi", action="store_true", help="run the rank-0-only secondary legs at N > 1 too "
                    "(they are in the N = 1 line; at N > 1 the other ranks would only wait for rank 0)")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import voxelis_b200 as vx
    from voxelis_b200 import sharding
    from voxelis_b200 import workloads as wl

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    solo_legs = rank == 0 and (world == 1 or args.secondary_multi)   # legs that only rank 0 runs
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: voxelis_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: stdout is ONE JSON line
        import datetime
        # a short collective timeout: a rank that fell out of step must fail the run in minutes, not hang the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier_all = barrier

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak, peak_src = peaks()
    stream = torch.cuda.Stream(dev)          # a real (non-default) stream: events on it bracket our launches

    def time_device(it, depth, n, dm, dv, dr, dc, steps, warmup=3, all_ranks=True):
        """`steps` x (reset + one apply call) on `stream`; -> (ms per step, mean ms of the apply call alone).
        all_ranks=False in the legs only rank 0 runs (no cross-rank barrier there: the other ranks are elsewhere)."""
        barrier = barrier_all if all_ranks else torch.cuda.synchronize
        for _ in range(warmup):
            it.reset_async(stream.cuda_stream)
            it.apply_batches_device(depth, n, dm.data_ptr(), dv.data_ptr(), dr.data_ptr(), dc.data_ptr() if dc is not None else 0,
                                    stream=stream.cuda_stream)
        torch.cuda.synchronize()
        it.sync()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        kev = []
        for _ in range(steps):
            it.reset_async(stream.cuda_stream)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            it.apply_batches_device(depth, n, dm.data_ptr(), dv.data_ptr(), dr.data_ptr(), dc.data_ptr() if dc is not None else 0,
                                    stream=stream.cuda_stream)
            k1.record(stream)
            kev.append((k0, k1))
        e1.record(stream)
        barrier()
        it.sync()
        return e0.elapsed_time(e1) / steps, float(np.mean([a.elapsed_time(b) for a, b in kev]))

    def stage_times(it, step, reps=30):
        it.profile_stages(True)
        acc = {}
        for _ in range(reps):
            step()
            for name, ms in it.stage_ms():
                acc[name] = acc.get(name, 0.0) + ms
        it.profile_stages(False)
        return {k: v / reps for k, v in acc.items()}

    # ---------------------------------------------------------------- inputs
    log("generating the perlin-dunes world")
    budget = BUDGET
    if args.workload == "perlin":
        masks, values = make_world(rank)
    elif args.workload == "below":
        masks, values = make_world(rank, "surface_and_below")
    else:
        masks, values = wl.named_workload(args.workload, 4096)
        budget = 2 << 30
    n = masks.shape[0]
    nonempty = int(np.count_nonzero(masks[:, :, 0].any(axis=1)))
    log(f"{n} chunks generated ({nonempty} non-empty); uploading")
    d_masks = torch.from_numpy(masks).to(dev)
    d_values = torch.from_numpy(values).to(dev)
    d_roots = torch.zeros(n, dtype=torch.int64, device=dev)
    d_changed = torch.zeros(n, dtype=torch.uint8, device=dev)
    it = vx.VoxInterner.with_memory_budget(budget, vx.U8, local_rank)
    torch.cuda.synchronize()

    def step_device():
        it.reset_async(stream.cuda_stream)
        it.apply_batches_device(DEPTH, n, d_masks.data_ptr(), d_values.data_ptr(), d_roots.data_ptr(),
                                d_changed.data_ptr(), stream=stream.cuda_stream)

    # ---------------------------------------------------------------- device-resident timing
    sampler = ClockSampler(local_rank)
    sampler.start()
    log("warm-up")
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    it.sync()
    new_nodes = it.stats()["total_cache_misses"]
    dbg = it.debug_counters()
    sampler.begin()
    step_ms, kern_ms = time_device(it, DEPTH, n, d_masks, d_values, d_roots, d_changed, args.steps, warmup=0)
    sampler.end()
    window = "timed region"
    if sampler.in_window() < 3:
        # the timed region was shorter than a few sampling periods: keep the same load running (untimed)
        # until the sampler has seen it, and say so
        t_probe = time.monotonic()
        while time.monotonic() - t_probe < 0.6:
            for _ in range(50):
                step_device()
            it.sync()
        sampler.end()
        window = "timed region + 0.6 s of identical untimed steps (region shorter than 3 sampling periods)"
    clocks = sampler.stop(window)
    step_ms_max = max_over_ranks(step_ms)
    total_chunks = sum_over_ranks(float(n))
    total_nonempty = sum_over_ranks(float(nonempty))
    value = total_chunks / (step_ms_max * 1e-3)

    log(f"device-resident: {step_ms_max:.3f} ms/step, apply {kern_ms:.3f} ms")
    # per-launch device times of the apply pipeline (CUDA events between its launches, outside the timed
    # region above: the extra event records would perturb it)
    stages = stage_times(it, step_device)
    launches_per_step = len(stages) + 2               # + the reset's two kernels (reset_used_kernel, init_scalars_kernel)
    dom = max(stages, key=stages.get) if stages else "apply_kernel"
    log("stages: " + ", ".join(f"{k} {v * 1e3:.1f} us" for k, v in stages.items()))
    # ---------------------------------------------------------------- end to end (host batches)
    e2e_value, e2e_ms_max, e2e_slab, touched_units, e2e_trace = None, None, None, 0, None
    if not args.no_e2e:
        # (1) the reference-facing call: n Batch handles (pinned host memory, filled by the caller before the
        # timed region, as a reference user fills Batch objects with set()) + n VoxTree handles ->
        # vx_apply_batches.  Inside the timed region: interner reset, the bus traffic of every touched unit,
        # the build, roots + changed flags back to the host, VoxTree handles updated.
        log("end-to-end: creating batch handles")
        trees = [vx.VoxTree(DEPTH, vx.U8) for _ in range(n)]
        batches = [t.create_batch() for t in trees]
        for b, m, v in zip(batches, masks, values):
            b.assign(m, v)
        touched_units = sum(b.touched_units for b in batches)
        cs = vx.ChunkSet(trees, batches)
        for _ in range(3):
            it.reset()
            cs.forget()
            cs.apply(it)
        e2e_steps = max(5, min(args.steps, 50))
        barrier()
        parts = [0.0, 0.0, 0.0]
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ta = time.perf_counter()
            it.reset_async()            # queued on the interner's stream, ahead of the apply
            tb = time.perf_counter()
            cs.forget()
            tc = time.perf_counter()
            cs.apply(it)                # synchronous: returns with every VoxTree handle updated
            td = time.perf_counter()
            parts[0] += tb - ta
            parts[1] += tc - tb
            parts[2] += td - tc
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        barrier()
        log("end-to-end host view per step: reset %.0f us, forget %.0f us, apply %.0f us"
            % tuple(1e6 * p / e2e_steps for p in parts))
        e2e_ms_max = max_over_ranks(e2e_ms)
        e2e_value = total_chunks / (e2e_ms_max * 1e-3)
        # (a Batch nobody wrote to reports "changed" with an EMPTY root, as in the reference: compare the roots)
        assert np.array_equal(cs.roots() != 0, d_changed.cpu().numpy() != 0), "handle path and device path disagree"
        log(f"end-to-end (batch handles, {touched_units} touched units): {e2e_ms_max:.3f} ms/step")
        it.profile_stages(True)
        it.reset()
        cs.forget()
        cs.apply(it)
        e2e_trace = it.host_trace()
        it.profile_stages(False)
        log(f"end-to-end phases: {e2e_trace}")
        # (2) the same world as two dense host arrays with no per-batch summary (vx_apply_batches_slab):
        # every mask byte has to cross the bus
        del cs, trees, batches                      # their pinned arena slots go back to the pool first
        h_masks = torch.from_numpy(masks).pin_memory()
        h_values = torch.from_numpy(values).pin_memory()
        roots_host = np.zeros(n, np.uint64)
        changed_host = np.zeros(n, np.uint8)
        for _ in range(2):
            it.reset()
            it.apply_batches_slab(DEPTH, h_masks.numpy(), h_values.numpy(), roots_out=roots_host,
                                  changed_out=changed_host)
        slab_steps = max(3, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(slab_steps):
            it.reset()
            it.apply_batches_slab(DEPTH, h_masks.numpy(), h_values.numpy(), roots_out=roots_host,
                                  changed_out=changed_host)
        torch.cuda.synchronize()
        slab_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / slab_steps)
        barrier()
        e2e_slab = {"value": total_chunks / (slab_ms * 1e-3), "unit": "chunks/s", "ms_per_step": slab_ms,
                    "h2d_bytes_per_step": int(masks.nbytes + 0),
                    "path": "vx_apply_batches_slab on two dense pinned arrays (masks H2D by copy engine in slabs, "
                            "values zero-copy for blocks with a set bit)"}
        log(f"end-to-end (dense slab): {slab_ms:.3f} ms/step")
        del h_masks, h_values

    # ---------------------------------------------------------------- roofline of the apply call
    # SURVEY §8(d) / BASELINE.md §3: masks + values read once + root written + 87 B per NEW node
    algo_bytes = n * CHUNK_BYTES + new_nodes * NODE_BYTES
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    # compulsory bytes for THIS input: values of blocks whose set_mask is 0 are never needed (phase 1
    # skips them, voxtree.rs:779-781) and the kernel does not load them
    touched_blocks = int(np.count_nonzero(masks[:, :, 0]))
    compulsory = n * (2 * 4096 + 8) + touched_blocks * 8 + new_nodes * NODE_BYTES
    stats_mem = it.memory()
    del it
    traffic, traffic_detail = None, None
    if solo_legs and not args.no_traffic:
        log("measuring DRAM traffic of one apply call (ncu replay of the same step)")
        t = measure_dram_traffic(args.workload if args.workload in ("perlin", "below") else f"{args.workload}:4096")
        if "error" in t:
            traffic_detail = {"source": "unavailable", "error": t["error"]}
            log("ncu replay failed: " + t["error"])
        else:
            traffic = t["dram_bytes"]
            traffic_detail = {"source": "ncu replay in this run (profiles/tools/ncu_traffic.py)", "read": t["read"], "write": t["write"],
                              "launches": t["launches"]}
            log(f"DRAM traffic per apply call: {traffic / 1e6:.1f} MB")

    # ---------------------------------------------------------------- the dense headline: surface-and-below terrain
    headline_dense = None
    if solo_legs and not args.no_others and args.workload == "perlin":
        try:
            log("dense headline: perlin surface-and-below (3 materials)")
            m2, v2 = make_world(0, "surface_and_below")
            ne2 = int(np.count_nonzero(m2[:, :, 0].any(axis=1)))
            tb2 = int(np.count_nonzero(m2[:, :, 0]))
            dm2, dv2 = torch.from_numpy(m2).to(dev), torch.from_numpy(v2).to(dev)
            itb = vx.VoxInterner.with_memory_budget(BUDGET, vx.U8, local_rank)
            sms, kms = time_device(itb, DEPTH, n, dm2, dv2, d_roots, None, max(20, min(args.steps, 200)), all_ranks=False)
            nn2 = itb.stats()["total_cache_misses"]

            def step2():
                itb.reset_async(stream.cuda_stream)
                itb.apply_batches_device(DEPTH, n, dm2.data_ptr(), dv2.data_ptr(), d_roots.data_ptr(), stream=stream.cuda_stream)
            st2 = stage_times(itb, step2, 10)
            ab2 = n * CHUNK_BYTES + nn2 * NODE_BYTES
            comp2 = n * (2 * 4096 + 8) + tb2 * 8 + nn2 * NODE_BYTES
            headline_dense = {"workload": "perlin_dunes_surface_and_below_3mat_64x8x64_d5_u8", "chunks": n, "nonempty_chunks": ne2,
                              "touched_blocks": tb2, "value": n / (sms * 1e-3), "value_nonempty": ne2 / (sms * 1e-3),
                              "unit": "chunks/s", "ms_per_step": sms, "kernel_ms": kms, "new_nodes": nn2,
                              "roofline": {"frac": ab2 / (kms * 1e-3) / 1e9 / peak, "achieved": ab2 / (kms * 1e-3) / 1e9,
                                           "algorithmic_bytes_per_launch": ab2, "compulsory_bytes_per_launch": comp2,
                                           "frac_compulsory": comp2 / (kms * 1e-3) / 1e9 / peak, "stages_ms": st2}}
            del dm2, dv2, itb
            if not args.no_traffic:
                t = measure_dram_traffic("below")
                if "error" not in t:
                    headline_dense["roofline"]["traffic"] = t["dram_bytes"]
                    headline_dense["roofline"]["frac_dram"] = t["dram_bytes"] / (kms * 1e-3) / 1e9 / peak
        except Exception as e:
            headline_dense = {"error": str(e)}

    # ---------------------------------------------------------------- config 1: one 32^3 set_uniform chunk, host Batch
    latency = None
    if solo_legs and not args.no_others:
        try:
            log("config 1: single-chunk latency")
            itl = vx.VoxInterner.with_memory_budget(BUDGET, vx.U8, local_rank)
            mu, vu = wl.named_workload("uniform", 1)
            tr = vx.VoxTree(DEPTH, vx.U8)
            b = tr.create_batch()
            b.assign(mu[0], vu[0])
            ts = []
            for i in range(260):
                itl.reset()
                vx.trees_forget([tr])
                t0 = time.perf_counter()
                tr.apply_batch(itl, b)            # synchronous: H2D of the touched units, launches, root back in the handle
                ts.append(time.perf_counter() - t0)
            ts = np.array(ts[10:]) * 1e6
            assert tr.is_leaf()
            from oracle import oracle
            to = []
            for i in range(200):
                t1, _ = oracle.time_apply_fresh(0, DEPTH, BUDGET, mu, vu, 1)
                to.append(t1 * 1e6)
            latency = {"workload": "single 32^3 chunk, set_uniform, u8, 256 MiB interner (BASELINE config 1)",
                       "gpu_us_median": float(np.median(ts)), "gpu_us_p10": float(np.percentile(ts, 10)),
                       "what_gpu": "vx_tree_apply_batch on a host Batch handle, wall clock around the synchronous call",
                       "oracle_us_median": float(np.median(to)), "oracle_note": "C++ port on one host core, apply only (interner creation excluded)",
                       "reference_published_us": 23.11, "reference_source": "docs/benches.md:221-222 (M3 Max, i32, includes old-tree teardown)"}
            del itl
        except Exception as e:
            latency = {"error": str(e)}

    # ---------------------------------------------------------------- secondary workloads (rank 0)
    others = {}
    if solo_legs and not args.no_others:
        # (name, generator, interner budget): random255 creates ~4 936 new nodes per chunk
        sets = [("checkerboard_x4096", lambda: wl.named_workload("checkerboard", 4096), BUDGET, "checkerboard:4096"),
                ("set_sum_x4096", lambda: wl.named_workload("sum", 4096), BUDGET, "sum:4096"),
                ("sum_per_chunk_x4096", lambda: wl.named_workload("sum_per_chunk", 4096), BUDGET, "sum_per_chunk:4096"),
                ("random255_x4096", lambda: wl.named_workload("random255", 4096), 2 << 30, "random255:4096"),
                ("set_sum_i32_x4096", lambda: wl.named_workload("sum", 4096, dtype=wl.I32), BUDGET, None),
                ("d6_64cube_cell4_random255_x256", lambda: wl.named_workload("cell4_random255", 256, depth=6), 1 << 30, None)]
        for name, gen, budget2, tname in sets:
            try:
                log(f"secondary workload {name}")
                m2, v2 = gen()
                n2 = m2.shape[0]
                depth2 = int(round(np.log2(m2.shape[1] * 8) / 3))
                chunk_bytes2 = 2 * m2.shape[1] + 8 * m2.shape[1] * (4 if v2.dtype == np.int32 else 1) + 8
                dt2 = vx.I32 if v2.dtype == np.int32 else vx.U8
                esz = 4 if dt2 == vx.I32 else 1
                it2 = vx.VoxInterner.with_memory_budget(budget2, dt2, local_rank)
                dm, dv = torch.from_numpy(m2).to(dev), torch.from_numpy(v2).to(dev)
                dr = torch.zeros(n2, dtype=torch.int64, device=dev)
                _, ms = time_device(it2, depth2, n2, dm, dv, dr, None, 20, all_ranks=False)
                nn = it2.stats()["total_cache_misses"]
                d2 = it2.debug_counters()

                def step2():
                    it2.reset_async(stream.cuda_stream)
                    it2.apply_batches_device(depth2, n2, dm.data_ptr(), dv.data_ptr(), dr.data_ptr(), stream=stream.cuda_stream)
                st2 = stage_times(it2, step2, 5)
                ab = n2 * chunk_bytes2 + nn * (NODE_BYTES + esz - 1)
                others[name] = {"depth": depth2, "chunks": n2, "chunks_per_s": n2 / (ms * 1e-3), "kernel_ms": ms,
                                "new_nodes": nn, "algorithmic_bytes": ab,
                                "achieved_gbs": ab / (ms * 1e-3) / 1e9, "frac": ab / (ms * 1e-3) / 1e9 / peak,
                                "branch_calls": d2["branch_calls"], "probe_steps": d2["probe_steps"],
                                "cache_hits_local": d2["cache_hits_local"], "stages_ms": st2}
                del dm, dv, dr, it2
                if args.traffic_all and tname and not args.no_traffic:
                    t = measure_dram_traffic(tname)
                    if "error" not in t:
                        others[name]["traffic"] = t["dram_bytes"]
                        others[name]["traffic_over_algorithmic"] = t["dram_bytes"] / ab
                        others[name]["frac_dram"] = t["dram_bytes"] / (ms * 1e-3) / 1e9 / peak
            except Exception as e:  # a secondary workload must never take the headline line down
                others[name] = {"error": str(e)}

        # the steps either side of the path, device-resident: batch generation before, mesher planes after
        try:
            log("secondary: device batch generation + occupancy planes")
            others["around_the_path"] = around_the_path(vx, dev, local_rank, peak)
        except Exception as e:
            others["around_the_path"] = {"error": str(e)}

    # ---------------------------------------------------------------- strong scaling: ONE world over the ranks (config 3)
    strong = None
    if world > 1 or args.scaling == "strong":
        log("strong scaling: one 64x8x64 world split into X slabs")
        lo, hi = sharding.slab_bounds(rank, world, GRID[0])
        per = GRID[1] * GRID[2]
        ms_, vs_ = make_world(0, grid=(hi - lo, GRID[1], GRID[2]), x_chunk_offset=lo)   # the slab of rank 0's world this rank owns
        ns = ms_.shape[0]
        assert ns == (hi - lo) * per
        its = vx.VoxInterner.with_memory_budget(BUDGET, vx.U8, local_rank)
        dms, dvs = torch.from_numpy(ms_).to(dev), torch.from_numpy(vs_).to(dev)
        drs = torch.zeros(ns, dtype=torch.int64, device=dev)
        s_ms, s_k = time_device(its, DEPTH, ns, dms, dvs, drs, None, max(50, min(args.steps, 500)))
        s_ms_max = max_over_ranks(s_ms)
        tot = sum_over_ranks(float(ns))
        one_gpu_ms = max_over_ranks(step_ms)       # every rank also timed a FULL world above (the weak leg)
        strong = {"world": WORLD_NAME, "chunks_total": int(tot), "chunks_per_gpu": ns, "slab_x": [lo, hi],
                  "ms_per_step": s_ms_max, "value": tot / (s_ms_max * 1e-3), "unit": "chunks/s",
                  "one_gpu_full_world_ms": one_gpu_ms, "speedup_vs_one_gpu": one_gpu_ms / s_ms_max,
                  "efficiency": one_gpu_ms / s_ms_max / world,
                  "note": "per-GPU interners, no data-path collective; a slab of 1/N of the world is a small call: the "
                          "level-synchronous pipeline has a floor of a few dependent round trips per launch"}
        del its, dms, dvs, drs

    # ---------------------------------------------------------------- config 4: MaxDepth 7, 2048 chunks per GPU
    d7 = None
    if not args.no_d7 and not args.no_others:
        try:
            log("config 4: 128^3 chunks, per-GPU share generated on the device")
            D7, n7 = 7, D7_CHUNKS_PER_GPU
            B7 = 8 ** (D7 - 1)
            budget7 = 56 << 30                     # 2048 x 299 593 new branches x 79 B
            it7 = vx.VoxInterner.with_memory_budget(budget7, vx.U8, local_rank)
            dm7 = torch.empty((n7, B7, 2), dtype=torch.uint8, device=dev)
            dv7 = torch.empty((n7, B7, 8), dtype=torch.uint8, device=dev)
            dr7 = torch.zeros(n7, dtype=torch.int64, device=dev)
            d7 = {"chunks_per_gpu": n7, "depth": D7, "bytes_of_batches_per_gpu": int(dm7.numel() + dv7.numel()),
                  "interner_budget_bytes": budget7, "per_gpu_interners": True, "distributions": {}}
            for name, k, cell in (("k255", 255, 1), ("k4", 4, 1), ("cell4_k255", 255, 4)):
                it7.random_batches_device(D7, n7, dm7.data_ptr(), dv7.data_ptr(), k, cell, chunk0=rank * n7, stream=stream.cuda_stream)
                torch.cuda.synchronize()
                s_ms, s_k = time_device(it7, D7, n7, dm7, dv7, dr7, None, 3, warmup=1)
                nn = it7.stats()["total_cache_misses"]
                s_ms_max = max_over_ranks(s_ms)
                ab = n7 * (10 * B7 + 8) + nn * NODE_BYTES
                d7["distributions"][name] = {"ms_per_step": s_ms_max, "chunks_per_s": world * n7 / (s_ms_max * 1e-3),
                                             "new_nodes_per_gpu": nn, "kernel_ms": s_k,
                                             "frac": ab / (s_k * 1e-3) / 1e9 / peak, "new_nodes_per_s_per_gpu": nn / (s_k * 1e-3)}
                log(f"  d7 {name}: {s_ms_max:.2f} ms/step, {nn} new nodes")
            d7["chunks_per_s"] = d7["distributions"]["k255"]["chunks_per_s"]
            d7["note"] = (f"BASELINE config 4 is 16 384 chunks over 8 GPUs = {n7} per GPU; every rank builds that share "
                          "(weak in N), batches generated in HBM by vx_random_batches_device, interner reset every step")
            del it7, dm7, dv7, dr7
        except Exception as e:
            d7 = {"error": str(e)}
            log(f"d7 leg failed: {e}")

    # ---------------------------------------------------------------- global dedup variant (config 5)
    dedup_info = None
    if (world > 1 and not args.no_dedup) or args.dedup:
        try:
            from voxelis_b200 import dedup as vd
            log("global dedup variant")
            itd = vx.VoxInterner.with_memory_budget(BUDGET, vx.U8, local_rank)
            rts, _ = itd.apply_batches_slab(DEPTH, d_masks.data_ptr(), d_values.data_ptr(), n=n)
            local_unique = itd.next_index - 1
            summ, dd_ms = None, None
            reps = []
            for rep in range(4):                   # first run: allocations + NCCL channel setup; then best of three
                shard = vx.VoxInterner.with_memory_budget(BUDGET, vx.U8, local_rank)   # this rank's (empty) global shard
                barrier()
                t0 = time.perf_counter()
                shard, groots, summ = vd.global_dedup(itd, rts, BUDGET, vx.U8, local_rank, shard=shard)
                torch.cuda.synchronize()
                barrier()
                reps.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
                del shard, groots
            dd_ms = min(reps[1:])
            dedup_info = {"ms": dd_ms, "ms_first_call": reps[0], "what": "one vx_world_global_dedup call per rank into an empty shard, wall clock, max over ranks",
                          "rounds": summ["rounds"], "global_unique_branches": summ["branches"],
                          "global_unique_leaves": summ["leaves"], "bytes_sent_all_ranks": summ["bytes_sent"],
                          "sum_of_per_gpu_unique_nodes": int(sum_over_ranks(float(local_unique))),
                          "exchange": summ.get("exchange", "single rank (no exchange)")}
            del itd
            if rank == 0:   # the reference's model: ONE interner for every chunk of every rank (oracle, CPU)
                from oracle import oracle
                ref = oracle.VoxInterner(4 * BUDGET)
                for r in range(world):
                    mr, vr = (masks, values) if r == 0 else make_world(r)
                    ref.apply_batches_fresh(DEPTH, mr, vr)
                st = ref.stats()
                dedup_info["oracle_single_interner"] = {"branches": st["branch_nodes"] - 1, "leaves": st["leaf_nodes"]}
                dedup_info["matches_oracle"] = bool(st["branch_nodes"] - 1 == summ["branches"] and st["leaf_nodes"] == summ["leaves"])
        except Exception as e:
            dedup_info = {"error": str(e)}
            log(f"dedup leg failed: {e}")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        log("cpu baseline")
        cpu = cpu_baseline(masks, values, 1)
        cpu["all_cores"] = cpu_baseline(masks, values, os.cpu_count() or 1, target_s=6.0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "chunks/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": base_config(n) if args.workload == "perlin" else dict(base_config(n), workload=args.workload),
            "workload_detail": {"chunks_per_step_per_gpu": n, "nonempty_chunks": nonempty, "touched_blocks": touched_blocks,
                                "new_nodes_per_step": new_nodes,
                                "l2": "inputs (1.34 GB/step) exceed the 126 MB L2; interner reset every step",
                                "step": "vx_interner_reset_async + one vx_apply_batches_device call",
                                "per_step_counters": dbg,
                                "device_memory": stats_mem},
            "value_nonempty": total_nonempty / (step_ms_max * 1e-3),
            "e2e": {"value": e2e_value, "unit": "chunks/s", "ms_per_step": e2e_ms_max,
                    # per step: descriptors by the copy engine; the batches' packed journals (block index + values of every
                    # block holding a voxel) by loads the staging kernel issues on pinned host memory (vx_stage.cuh)
                    "h2d_bytes_per_step": int(n * 12 + touched_units * 4 + touched_blocks * 10),
                    "h2d_bytes_note": "descriptors (8 + 4 B per chunk, 4 B per touched unit) + the batches' journals: 2 B of block "
                                      "index + 8 B of values per flagged block, contiguous per batch; ncu pcie__read_bytes for the "
                                      "staging kernels of one world: 19.8 MB",
                    "host_input_bytes_per_step": int(masks.nbytes + values.nbytes),
                    "d2h_bytes_per_step": int(n * 9),
                    "touched_units": int(touched_units),
                    "path": "vx_apply_batches on n Batch + n VoxTree handles (batches in pinned host memory; "
                            "stage_journal_kernel pulls the batches' journals over PCIe, builders run on the device slab), "
                            "roots + changed flags D2H into the VoxTree handles",
                    "phases_us": e2e_trace, "dense_slab_variant": e2e_slab},
            "gpu_launches": args.steps * launches_per_step,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_detail": traffic_detail, "peak_source": peak_src,
                         "frac_compulsory": compulsory / (kern_ms * 1e-3) / 1e9 / peak,
                         "frac_dram": (traffic / (kern_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         # the path is a pipeline of dependent launches (vx_bulk.cuh): the roofline is taken over
                         # the whole apply call, the per-launch times are listed beside it
                         "kernel": "apply call = " + " + ".join(stages) if stages else "apply_kernel<u8,false>",
                         "stages_ms": stages, "dominant_stage": dom,
                         "dominant_share": (stages[dom] / sum(stages.values())) if stages else 1.0,
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": algo_bytes,
                         "compulsory_bytes_per_launch": compulsory,
                         "achieved_compulsory": compulsory / (kern_ms * 1e-3) / 1e9,
                         "note": "frac follows the SURVEY 8(d) dense formula; the input is sparse "
                                 f"({touched_blocks} of {n * 4096} blocks have a set bit, {nonempty} of {n} chunks are "
                                 "non-empty) and the kernels skip the values of untouched blocks: frac_compulsory counts "
                                 "only the bytes this input needs, frac_dram the DRAM bytes ncu measured for the same step"},
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if headline_dense:
            line["headline_dense"] = headline_dense
        if latency:
            line["latency_single_chunk"] = latency
        if others:
            line["others"] = others
        if strong:
            line["strong"] = strong
        if d7:
            line["d7"] = d7
        if dedup_info:
            line["global_dedup"] = dedup_info
        emit(line)
    if world > 1:
        dist.destroy_process_group()


WORLD_NAME = "perlin_dunes_surface_only_64x8x64_d5_u8 (rank 0's world)"

if __name__ == "__main__":
    main()
