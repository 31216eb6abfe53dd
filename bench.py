#!/usr/bin/env python
"""bench.py — ICP scans/sec on BASELINE.json configs[1]:
64k-pt synthetic Velodyne HDL-64 sweeps registered against a 500k-pt accumulated local map,
30 ICP iterations max (point-to-point pipeline of the north_star), one B200 per rank.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

A "step" is one pass of the hot path over one batch of B sweeps: b2icp_align_batch (one fused sweep
launch per ICP iteration for the whole batch) against the resident map, then — when N > 1 — the NCCL
gather of the per-scan rigid transforms (the only exchange the path has).  Weak
scaling: every rank owns B sweeps; the map is replicated.

  value   scans/s with the sweeps already resident in HBM (device pointers through the C ABI)
  e2e     scans/s through the same C ABI with pinned HOST buffers: H2D of every sweep and D2H of the
          results inside the timed region
  roofline  fused sweep kernel: algorithmic bytes / CUDA-event launch time, vs the measured HBM peak
  cpu_baseline  the CPU oracle (kind "port": the reference's PCL path cannot be built here) timed on
          this box's host cores on a bounded sample of the same workload

`--impl reference` times the oracle alone (all host threads) on the same workload and prints the
same line with "impl": "reference".  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from icpslam_b200 import synth  # noqa: E402

CONFIG_ID = 2          # seeds = 1000 * config + index
N_MAP = 500_000
N_SWEEP = 65_536
MAX_ITERS = 30
METRIC = "icp_scans_per_sec_64k_sweeps_30_iters"
UNIT = "scans/s"
STATE_BYTES = 192      # sizeof(IcpState): what comes back per scan


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def load_workload(first: int, count: int, cache_dir: str = "/tmp/b2icp_bench_cache"):
    """(map[500000,4], [count sweeps already in the map frame up to a small odometry error])."""
    return prepare_workloads([first], count, cache_dir)[first]


def prepare_workloads(firsts, count: int, cache_dir: str = "/tmp/b2icp_bench_cache") -> dict:
    """{first: (map, sweeps)} for every start index in `firsts`; the map is generated once and all of
    it is cached under /tmp as .npy so that back-to-back runs on one box skip the ray casting."""
    os.makedirs(cache_dir, exist_ok=True)
    mpath = os.path.join(cache_dir, f"map_c{CONFIG_ID}_{N_MAP}.npy")
    qpaths = {f: os.path.join(cache_dir, f"q_c{CONFIG_ID}_{f}_{count}.npy") for f in firsts}
    out, m = {}, None
    for f in firsts:
        if os.path.exists(mpath) and os.path.exists(qpaths[f]):
            out[f] = (np.load(mpath), list(np.load(qpaths[f])))
            continue
        t0 = time.time()
        if m is None:
            m = synth.build_local_map(CONFIG_ID, n_points=N_MAP)
        qs, _ = synth.map_queries(m, CONFIG_ID, f, count)
        log(f"[bench] generated workload (first sweep {f}, {count} sweeps) in {time.time() - t0:.1f}s")
        try:
            tmp = os.path.join(cache_dir, f"tmp_{os.getpid()}.npy")
            if not os.path.exists(mpath):
                np.save(tmp, m["map"])
                os.replace(tmp, mpath)
            np.save(tmp, np.stack(qs))
            os.replace(tmp, qpaths[f])
        except OSError:
            pass
        out[f] = (m["map"], qs)
    return out


def touched_target_points(map_xyzw, queries, cell, origin, dims, radius):
    """N_t' of SURVEY.md §8d: map points lying in grid cells within `radius` of any query's cell."""
    dims = np.asarray(dims, dtype=np.int64)
    R = int(np.ceil(radius / cell))
    q = np.concatenate([x[:, :3] for x in queries]).astype(np.float64)
    qc = np.clip(np.floor((q - origin) / cell).astype(np.int64), 0, dims - 1)
    lin = np.unique((qc[:, 2] * dims[1] + qc[:, 1]) * dims[0] + qc[:, 0])
    cz, rem = np.divmod(lin, dims[0] * dims[1])
    cy, cx = np.divmod(rem, dims[0])
    touched = np.zeros(int(dims.prod()), dtype=bool)
    for dz in range(-R, R + 1):
        z = cz + dz
        okz = (z >= 0) & (z < dims[2])
        for dy in range(-R, R + 1):
            y = cy + dy
            ok = okz & (y >= 0) & (y < dims[1])
            base = (z[ok] * dims[1] + y[ok]) * dims[0]
            x0 = np.clip(cx[ok] - R, 0, dims[0] - 1)
            x1 = np.clip(cx[ok] + R, 0, dims[0] - 1)
            for dx in range(0, 2 * R + 1):
                x = np.minimum(x0 + dx, x1)
                touched[base + x] = True
    m = map_xyzw[:, :3].astype(np.float64)
    mc = np.clip(np.floor((m - origin) / cell).astype(np.int64), 0, dims - 1)
    ml = (mc[:, 2] * dims[1] + mc[:, 1]) * dims[0] + mc[:, 0]
    return int(touched[ml].sum())


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU oracle legs (the only places bench.py may execute oracle/)
# ------------------------------------------------------------------------------------------------
def oracle_scans_per_sec(map_xyzw, sweeps, threads: int):
    """P2P align of each sweep vs the map with the CPU oracle; the k-d tree build is excluded from the
    time like the resident grid is excluded on the GPU side.  Returns (scans/s, threads, iterations)."""
    from oracle import oracle as O
    O.build()
    threads = O.set_threads(threads)
    p = O.default_params("mapper")
    busy_ms, iters = 0.0, []
    for s in sweeps:
        r = O.align(p, s, map_xyzw)
        busy_ms += r["stages"]["total"] - r["stages"]["build"]
        iters.append(r["iterations"])
    return len(sweeps) / (busy_ms * 1e-3), threads, iters


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = O.max_threads()
    per_step = 2
    map_xyzw, sweeps = load_workload(0, max(args.batch, per_step))
    t_steps = []
    for k in range(args.warmup + args.steps):
        sample = [sweeps[(k * per_step + j) % len(sweeps)] for j in range(per_step)]
        t0 = time.perf_counter()
        sps, used, iters = oracle_scans_per_sec(map_xyzw, sample, threads)
        if k >= args.warmup:
            t_steps.append(per_step / sps)
    total = float(np.sum(t_steps))
    value = per_step * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, world, extra={"reference_sample": f"{per_step} sweeps per step"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"{per_step} sweeps/step x {args.steps} steps vs the 500k map, k-d tree build "
                                   f"excluded (map resident), OpenMP over queries"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(batch: int, world: int, extra=None) -> dict:
    c = {
        "workload": "BASELINE configs[1]: 64k-pt synthetic HDL-64 sweep vs 500k-pt accumulated local map, "
                    "30 ICP iterations max, point-to-point (north_star pipeline)",
        "sweep_points": N_SWEEP, "map_points": N_MAP, "max_iterations": MAX_ITERS, "transformation_epsilon": 1e-6,
        "max_correspondence_distance": 1.0, "batch_per_gpu": batch, "global_batch": batch * world,
        "sharding": "scans sharded across ranks, map replicated; NCCL all_gather of per-scan transforms per step",
        "l2": "a 256 MiB write is enqueued between step submissions (in the streamed legs it runs next to the batches in "
              "flight); independently of it the per-step working set (150 MB of per-query state per 32-sweep step, up to "
              "8 steps in flight) is larger than the 126 MB L2",
    }
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="sweeps per GPU per step")
    ap.add_argument("--impl", default="b2icp", choices=["b2icp", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=4, help="sweeps timed on the CPU oracle (rank 0, N=1)")
    ap.add_argument("--in-flight", type=int, default=8, help="streamed batches in flight (1..8)")
    ap.add_argument("--grid-cell", type=float, default=0.0, help="neighbour-grid cell edge in metres (0 = auto); tuning only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b2icp" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from icpslam_b200 import build as B
    from icpslam_b200 import registration as R

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libb2icp.so has no CPU fallback")
    B.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # rank 0 generates (or finds) the cached workload first so that the other ranks only load it
    if world > 1 and rank != 0:
        dist.barrier()
    if rank == 0:  # also pre-generates the other ranks' sweeps into the cache
        map_xyzw, sweeps = prepare_workloads([r * args.batch for r in range(world)], args.batch)[0]
        if world > 1:
            dist.barrier()
    else:
        map_xyzw, sweeps = load_workload(rank * args.batch, args.batch)

    stream = torch.cuda.current_stream()
    reg = R.Registration(preset=R.PRESET_MAPPER, device=local_rank, profile=1, grid_cell=args.grid_cell)
    reg.setStream(stream.cuda_stream)
    reg.setInputTarget(map_xyzw)
    grid = reg.gridInfo()

    Bn = args.batch
    d_sweeps = [torch.from_numpy(s).to(dev) for s in sweeps]
    d_ptrs = [t.data_ptr() for t in d_sweeps]
    n_src = [N_SWEEP] * Bn
    h_sweeps = []
    for s in sweeps:
        p = R.pinned_empty((N_SWEEP, 4))
        p[:] = s
        h_sweeps.append(p)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    gathered = torch.empty((world * Bn, 20), dtype=torch.float64, device=dev) if world > 1 else None

    def gather(results):
        if world == 1:
            return
        loc = np.array([list(r.T) + [r.converged, r.iterations, r.n_corr_last, r.mse_last] for r in results])
        dist.all_gather_into_tensor(gathered, torch.from_numpy(loc).to(dev, non_blocking=False))

    def step_resident():
        rc, res = reg.alignBatchDevice(d_ptrs, n_src)
        if rc:
            raise RuntimeError(f"b2icp_align_batch_device rc={rc}")
        gather(res)
        return res

    def step_e2e():
        rc, res = reg.alignBatch(h_sweeps)
        if rc:
            raise RuntimeError(f"b2icp_align_batch rc={rc}")
        gather(res)
        return res

    def timed(step_fn, steps, warmup, collect):
        for _ in range(warmup):
            step_fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_wall0 = time.perf_counter()
        last = None
        for k in range(steps):
            flush.zero_()  # L2 flush between timed steps, outside the per-step events
            ev[k][0].record(stream)
            last = step_fn()
            ev[k][1].record(stream)
            collect(last)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t_wall0
        dev_s = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
        t = torch.tensor([dev_s, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), last

    # ---- HBM-resident leg -----------------------------------------------------------------------
    sweep_ms, sweep_launches, iters_hist, searches = [], [], [], []

    def collect(res):
        tm = reg.timing()
        sweep_ms.append(tm.nn_sweep_ms)
        sweep_launches.append(tm.nn_sweep_launches)
        searches.append(tm.nn_searches)
        iters_hist.append([r.iterations for r in res])

    # 1. profiled synchronous pass (one call per step, CUDA events around every sweep launch inside the
    #    library): the per-launch durations the roofline needs.  Not the headline numbers.
    prof_dev_s, _, last = timed(step_resident, args.steps, args.warmup, collect)

    # 2. / 3. the two timed legs use the streaming form of the batch call (b2icp_align_batch_submit[_device] /
    #    _wait): up to --in-flight steps are submitted before the oldest is waited for, each on its own stream, so
    #    the host never leaves the device idle between steps, the sparse late iterations of one step share the
    #    device with the first iterations of the next and — in the end-to-end leg — the PCIe upload of the next
    #    steps overlaps the sweeps of the current one.
    #    Every step's copies (H2D of its 32 sweeps, D2H of its results) and the L2 flush between steps are inside
    #    the timed region: one event pair around all K steps.
    def run_streamed(steps, submit):
        submitted = done = 0
        while done < steps:
            while submitted < steps and submitted - done < args.in_flight:
                if submitted:
                    flush.zero_()
                if submit():
                    raise RuntimeError("b2icp_align_batch_submit failed")
                submitted += 1
            rc, res = reg.alignBatchWait()
            if rc:
                raise RuntimeError(f"b2icp_align_batch_wait rc={rc}")
            gather(res)
            done += 1

    def timed_streamed(submit, steps, warmup):
        run_streamed(max(warmup, 8), submit)  # every one of the library's 8 slot sets allocates its buffers once
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = reg.timing().kernel_launches
        e0.record(stream)
        run_streamed(steps, submit)
        e1.record(stream)
        timed_streamed.launches = reg.timing().kernel_launches - n0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    streamed = Bn <= 32
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = reg.timing().kernel_launches
    if streamed:
        dev_s = timed_streamed(lambda: reg.alignBatchSubmitDevice(d_ptrs, n_src), args.steps, args.warmup)
        wall_s = dev_s
    else:
        dev_s, wall_s, last = timed(step_resident, args.steps, args.warmup, lambda res: None)
    launches1 = reg.timing().kernel_launches
    if streamed:
        launches0, launches1 = 0, timed_streamed.launches  # the launches of the K timed steps only
    clocks = sampler.stop() if rank == 0 else None
    value = world * Bn * args.steps / dev_s

    if streamed:
        e2e_dev_s = timed_streamed(lambda: reg.alignBatchSubmit(h_sweeps), args.steps, 3)
        api = f"b2icp_align_batch_submit[_device] / b2icp_align_batch_wait ({args.in_flight} batches in flight)"
    else:
        e2e_dev_s, _, _ = timed(step_e2e, args.steps, 3, lambda res: None)
        api = "b2icp_align_batch[_device]"
    e2e_api = api
    e2e_value = world * Bn * args.steps / e2e_dev_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused sweep) -------------------------------------------
    origin = map_xyzw[:, :3].min(axis=0).astype(np.float64)
    nt_touched = touched_target_points(map_xyzw, sweeps, grid["cell"], origin, grid["dims"], 1.0)
    alg_bytes = 0.0
    for it_list in iters_hist:
        its = np.asarray(it_list)
        for launch in range(int(its.max())):
            active = int((its > launch).sum())
            alg_bytes += active * (16 + 8) * N_SWEEP + 16.0 * nt_touched * active / Bn
    kernel_s = float(np.sum(sweep_ms)) * 1e-3
    n_launch = int(np.sum(sweep_launches))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / kernel_s / 1e9 if kernel_s > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("icp_sweep_p2p_dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "icp_sweep_p2p (one launch = one ICP iteration of the whole batch)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(n_launch, 1),
                "avg_launch_us": 1e6 * kernel_s / max(n_launch, 1), "launches": n_launch,
                "nt_touched": nt_touched, "kernel_share_of_step": kernel_s / prof_dev_s,
                # the same algorithmic bytes over the step time of the timed (streamed, overlapping) leg
                "achieved_in_streamed_leg": alg_bytes / dev_s / 1e9,
                "measured_in": "a synchronous b2icp_align_batch_device pass of the same steps with CUDA events around "
                               "every launch (params.profile = 1)",
                # share of (query, iteration) pairs that needed a real search; the rest were settled by the
                # cached-neighbour certificate (icpslam_b200/csrc/nncache.cuh)
                "searched_fraction": float(np.sum(searches)) / max(1.0, float(sum(sum(x) for x in iters_hist)) * N_SWEEP)}

    # ---- the stand-alone NN search (b2icp_nn_search_device): the second half of BASELINE.json's metric ----
    # all 32 sweeps of the step as ONE query cloud (2.1 M queries) against the 500k map, exact unbounded 1-NN,
    # no seeds: algorithmic bytes = 16 n_q + 16 N_t' + 8 n_q, CUDA events inside the library.
    allq = torch.cat(d_sweeps)
    nq = allq.shape[0]
    d_idx = torch.empty(nq, dtype=torch.int32, device=dev)
    d_d2 = torch.empty(nq, dtype=torch.float32, device=dev)
    nn_ms = []
    for k in range(3 + 5):
        flush.zero_()
        reg.nearestKSearch1Device(allq.data_ptr(), nq, d_idx.data_ptr(), d_d2.data_ptr())
        if k >= 3:
            nn_ms.append(reg.timing().nn_sweep_ms)
    nn_bytes = 24.0 * nq + 16.0 * nt_touched
    nn_search = {"kernel": "nn_search_box_kernel (+ nn_brute_fallback)", "queries": int(nq), "ms": float(np.mean(nn_ms)),
                 "algorithmic_bytes": nn_bytes, "achieved": nn_bytes / (float(np.mean(nn_ms)) * 1e-3) / 1e9, "unit": "GB/s",
                 "frac": nn_bytes / (float(np.mean(nn_ms)) * 1e-3) / 1e9 / peak,
                 "queries_per_s": nq / (float(np.mean(nn_ms)) * 1e-3)}

    # ---- CPU baseline on this box's host cores (bounded sample) -----------------------------------
    cpu = None
    if world == 1 and args.cpu_sample > 0:
        from oracle import oracle as O
        O.build()
        sps, used, its = oracle_scans_per_sec(map_xyzw, sweeps[:args.cpu_sample], O.max_threads())
        cpu = {"value": sps, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"{args.cpu_sample} of the {Bn} sweeps vs the same 500k map, oracle P2P align with OpenMP over "
                         f"queries, k-d tree build excluded (map resident); iterations {its}"}

    its_all = np.concatenate([np.asarray(x) for x in iters_hist])
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(Bn, world, extra={
            "grid_cell_m": grid["cell"], "grid_dims": list(grid["dims"]), "grid_occupancy": grid["occupancy"],
            "mean_iterations": float(its_all.mean()), "max_iterations_seen": int(its_all.max()),
            "wall_ms_per_step": 1e3 * wall_s / args.steps, "api": api,
            "synchronous_call_scans_per_s": world * Bn * args.steps / prof_dev_s}),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": Bn * N_SWEEP * 16,
                "d2h_bytes_per_step": Bn * STATE_BYTES, "ms_per_step": 1e3 * e2e_dev_s / args.steps, "api": e2e_api},
        "gpu_launches": int(launches1 - launches0),
        "roofline": roofline,
        "nn_search": nn_search,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
