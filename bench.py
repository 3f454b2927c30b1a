#!/usr/bin/env python
"""Benchmark of the cpic hot path (BASELINE.json: particle-steps/s of the full step
push+deposit+gather+solve at 1/2/4/8 B200, and the fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload D|A|cyc|C|2s]

The default workload is BASELINE.json configs[4] ("D": weak scaling, 2048x2048 cells and 2.5e8
particles PER GPU, a warm electron beam crossing 0.77 cells per step over resting ions); with one
GPU it is also the largest single-GPU configuration. The same line carries, as extra objects:

    other_workloads   N = 1: configs[1] (conf/2d-2species.conf, 1024^2, 1e7 particles, the
                      reference's own initial conditions) and the two-stream plasma, device resident
    target_config     N > 1: configs[3] (strong scaling: 4096^2 grid, 1e9 particles over the N
                      GPUs, distributed FFT) -- the north-star target number
    multi_rank_check  N > 1: a small global problem on the N ranks against the same problem on
                      one rank (product only), fields and particles to 1e-12, count and charge
    f1                N = 1: the deposits the UNMODIFIED reference drops in configs[1]
                      and what that does to rho, next to the accumulate-correct variant

A "step" is one sim_step (reference src/sim.c:481-581). N > 1: torchrun, one rank per GPU.
`--impl reference` times the reference's own CPU implementation (oracle/_ref, unmodified
sources) on the host cores: one rank per core through the fork-based MPI shim.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec (push+deposit+gather+solve)"
UNIT = "particle-steps/s"
# Algorithmic bytes per particle-step (SURVEY 8d / DESIGN.md): the fused gather+push kernel reads
# x,y,ux,uy,uz and writes them back: 80 B; the deposit reads x,y: 16 B.
BYTES_GATHER_PUSH = 80.0
BYTES_DEPOSIT = 16.0
TINY = bool(os.environ.get("BENCH_TINY"))      # CPU dry run of this script (tests/test_simt_check.py)

# nx, ny: cells per GPU (weak) or in total (strong); nps: particles per species per GPU (weak) or in
# total (strong); init: "reference" = the reference's host initialiser on the conf (bit-identical
# initial conditions), "device" = device initialiser u = drift + U(-spread, spread)
WORKLOADS = {
    "A": dict(conf="2d-2species.conf", nx=1024, ny=1024, nps=5_000_000, init="reference", strong=False,
              drift=[(0.0, 0.0), (0.0, 0.0)], spread=[(5.0, 0.0), (3.0, 0.0)],
              what="BASELINE configs[1]: conf/2d-2species.conf, 1024x1024 grid, 1e7 particles"),
    "cyc": dict(conf="cyclotron-2048.conf", nx=2048, ny=2048, nps=100_000_000, init="device", strong=False,
                drift=[(0.0, 0.0)], spread=[(10.0, 10.0)],
                what="BASELINE configs[2]: cyclotron physics (uniform B), 2048x2048 grid, 1e8 particles"),
    "C": dict(conf="2d-2species.conf", nx=4096, ny=4096, nps=500_000_000, init="device", strong=True,
              drift=[(0.0, 0.0), (0.0, 0.0)], spread=[(5.0, 0.0), (3.0, 0.0)],
              what="BASELINE configs[3]: strong scaling, 4096x4096 grid, 1e9 particles, Y slabs, distributed FFT"),
    "D": dict(conf="hot-beam.conf", nx=2048, ny=2048, nps=125_000_000, init="device", strong=False,
              drift=[(0.45, 0.6), (0.0, 0.0)], spread=[(0.15, 0.15), (0.05, 0.05)],
              what="BASELINE configs[4]: weak scaling, 2048x2048 cells and 2.5e8 particles per GPU, warm electron "
                   "beam (0.58, 0.77) cells per step across the slab faces over resting ions"),
    "2s": dict(conf="two-streams-1024.conf", nx=1024, ny=1024, nps=5_000_000, init="reference", strong=False,
               drift=[(1.0, 0.0), (-1.0, 0.0)], spread=[(0.0, 0.0), (0.0, 0.0)],
               what="two-stream instability: conf/two-streams.conf scaled to 1024x1024 cells, 2 x 5e6 electrons, "
                    "position-delta initialisation, +-0.51 cells per step"),
}
if TINY:
    for w in WORKLOADS.values():
        w.update(nx=64, ny=64, nps=20_000 if not w["strong"] else 40_000)


def dbg(msg):
    if os.environ.get("BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def ncu_traffic(kernel="k_gather_push<2>", path=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed
    `ncu --set full` extract of this round's build (profiles/)."""
    import csv
    path = path or os.path.join(ROOT, "profiles", "r2_ncu_full_summary.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        vals = [float(r[ir]) + float(r[iw]) for r in rows[2:] if kernel in r[0]]
        unit = rows[1][ir]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit]
        return sum(vals) / len(vals) * scale if vals else None
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_label(name, world):
    """The same string in both arms (the driver compares config.workload)."""
    w = WORKLOADS[name]
    if w["strong"]:
        return f"{name}: {w['what']} (strong: divided over {world} GPU(s))"
    return f"{name}: {w['what']}" + (f" (weak: that much per GPU on {world} GPUs)" if world > 1 else "")


def scaled_conf(conf, w, nps, ny=None, cycles=None, stop_sem=None, tmpdir="/tmp"):
    """A copy of conf/<conf> with the workload's grid, `nps` particles per species and the physical
    lengths / permittivity scaled so that the cell size and the plasma frequency of the conf are kept."""
    text = open(conf).read()
    nx0, ny0 = [int(v) for v in re.search(r"points\s*=\s*\[\s*(\d+)\s*,\s*(\d+)\s*\]", text).groups()]
    n0 = int(re.search(r"particles\s*=\s*(\d+)", text).group(1))
    lx0, ly0 = [float(v) for v in re.search(r"space_length\s*=\s*\[\s*([^,\]]+),\s*([^\]]+)\]", text).groups()]
    e00 = float(re.search(r"vacuum_permittivity\s*=\s*([^\s;]+)", text).group(1))
    nx, ny = w["nx"], (ny or w["ny"])
    text = re.sub(r"particles\s*=\s*\d+", f"particles = {nps}", text)
    text = re.sub(r"points\s*=\s*\[[^\]]*\]", f"points = [{nx}, {ny}]", text)
    text = re.sub(r"space_length\s*=\s*\[[^\]]*\]", f"space_length = [{lx0 * nx / nx0!r}, {ly0 * ny / ny0!r}]", text)
    e0 = e00 * (nps / n0) / ((nx / nx0) * (ny / ny0))
    text = re.sub(r"vacuum_permittivity\s*=\s*[^\s;]+", f"vacuum_permittivity = {e0!r}", text)
    if cycles is not None:
        text = re.sub(r"cycles\s*=\s*\d+", f"cycles = {cycles}", text)
    if stop_sem is not None:
        text = re.sub(r"stop_SEM\s*=\s*[^\s;]+", f"stop_SEM = {stop_sem!r}", text)
    path = os.path.join(tmpdir, f"cpic_b200_bench_{os.getpid()}_{os.path.basename(conf)}")
    with open(path, "w") as f:
        f.write(text)
    return path


# ----------------------------------------------------------------------------- reference arm

def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference(name, steps, warmup, budget_s=150.0, ranks=None):
    """Times the reference's own CPU implementation of the path on a bounded sample of workload
    `name`: oracle/_ref/cpic_ref_mp = the UNMODIFIED reference sources (gcc -O3 AVX2; OmpSs-2
    pragmas inert) running their own rank decomposition -- Y slabs, src/sim.c:116-130 -- with one
    forked rank per host core (oracle/shim/shim_mpi_mp.c: sockets and a shared mapping stand in
    for MPI; the FFTW-MPI transform is a shared-memory transform over the same ranks). The time is
    the reference's own TIMER_ITERATION on rank 0, read from its `stats` lines (src/sim.c:440-479).
    Falls back to the single-rank in-process library, then to the oracle port."""
    w = WORKLOADS[name]
    conf = os.path.join(ROOT, "conf", w["conf"])
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    cores = host_cores()
    P = 1
    while P * 2 <= min(cores, ranks or cores) and w["ny"] % (P * 2) == 0 and w["ny"] // (P * 2) >= 16:
        P *= 2
    # bound the sample: ~5e6 particle-steps/s per core is typical of the reference
    nsp = len(w["drift"])
    max_n = int(budget_s * 5e6 * P / max(1, steps + warmup + 2))
    max_n = min(max_n, 40_000_000)     # the reference's host lists take 112 B per particle (src/def.h:88-133)
    nps = min(w["nps"], max(1000, max_n // nsp))
    sample = scaled_conf(conf, w, nps, cycles=warmup + steps, stop_sem=1e-30)
    n = nps * nsp
    full = w["nps"] * nsp
    exe = os.path.join(ref_dir, "cpic_ref_mp")
    info = None
    if os.path.exists(exe):
        env = dict(os.environ, CPIC_SHIM_NPROCS=str(P), CPIC_SHIM_ARENA_MB="4096")
        t0 = time.perf_counter()
        r = subprocess.run([exe, "-q", sample], env=env, capture_output=True, text=True, timeout=3600)
        wall = time.perf_counter() - t0
        last = [float(m.group(1)) for m in re.finditer(r"^stats iter=\d+ last=([0-9.eE+-]+)", r.stdout, re.M)]
        if r.returncode == 0 and len(last) >= warmup + steps:
            t = last[warmup:warmup + steps]
            mean = sum(t) / len(t)
            sem = (sum((v - mean) ** 2 for v in t) / max(1, len(t) - 1)) ** 0.5 / len(t) ** 0.5
            info = {"value": n / mean, "unit": UNIT, "cores": P, "kind": "reference",
                    "ms_per_step": mean * 1e3, "sem_ms": sem * 1e3, "wall_s": wall}
        else:
            dbg("cpic_ref_mp failed: " + (r.stderr or r.stdout)[-500:])
    if info is None:
        # single rank, in process (tests/_refbind.py), else the oracle port
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _refbind import RefSim, ref_available
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        if ref_available("ref"):
            kind, sim = "reference", RefSim(sample, "ref")
        else:
            from cpic_b200 import load_conf, init_particles
            from _parity import oracle_from
            kind, sim = "port", oracle_from(load_conf(sample)[0], init_particles(sample))
            sim.pre_step()
        for _ in range(warmup):
            sim.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            sim.step()
        mean = (time.perf_counter() - t0) / steps
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
        P = 1
        info = {"value": n / mean, "unit": UNIT, "cores": 1, "kind": kind, "ms_per_step": mean * 1e3}
    info["sample"] = (f"{w['conf']} at the workload's grid ({w['nx']}x{w['ny']}), {n} particles "
                      f"({'the full workload' if n == full else f'{n}/{full} of one GPU share of the workload; the reference is linear in the particle number, perf/particles/csv/regression.csv'}), "
                      f"{steps} steps after {warmup} warm-up, {P} rank(s) x 1 core of {cores} host cores "
                      f"(the reference's Y-slab ranks; OmpSs-2 tasks inert), shim FFT; time = the reference's TIMER_ITERATION")
    return info["value"], info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 30)
    warmup = min(args.warmup, 5)
    value, info = cpu_reference(args.workload, steps, warmup, budget_s=120.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if WORKLOADS[args.workload]["strong"] else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_label(args.workload, args.gpus),
                       "host": f"CPU, {info['cores']} core(s): the job is NOT scaled with --gpus (one host)"},
            "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample") if k in info},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if "sem_ms" in info:
        line["timer_iteration"] = {"mean_ms": info["ms_per_step"], "sem_ms": info["sem_ms"]}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------- our arm

class Env:
    """Process group, device and the few collectives the bench itself needs."""

    def __init__(self, gpus):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world == 1 and gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
        self.cuda = not TINY
        self.dist = None
        if self.cuda:
            torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            if self.cuda:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            else:
                dist.init_process_group("gloo")
        self.dev = "cuda" if self.cuda else "cpu"

    def barrier(self):
        if self.cuda:
            self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        if self.cuda:
            self.torch.cuda.synchronize()

    def reduce(self, v, op="max"):
        if not self.dist:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM,
                                    "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    def bootstrap(self, sim):
        if self.dist:
            from cpic_b200.dist import bootstrap
            bootstrap(sim, self.dist, device=self.dev)

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def build_sim(env, name, reference_ics=True):
    """sim_init of workload `name` on this rank: params, particles, communicator, pre-step."""
    from cpic_b200 import Sim, load_conf, init_particles
    w = WORKLOADS[name]
    world, rank = env.world, env.rank
    ny, nps = w["ny"], w["nps"]
    if w["strong"]:
        ny, nps = ny // world, nps // world          # per GPU: a slab of the fixed global problem
    conf = os.path.join(ROOT, "conf", w["conf"])
    sample = scaled_conf(conf, w, nps * world, ny=ny * world)
    params, run = load_conf(sample, rank=rank, nranks=world, device=env.local if env.cuda else 0)
    nspecies = len(params.q)
    assert nspecies == len(w["drift"])
    # exchange regions: what the fastest particles can send across one block side per step, with
    # slack; the capacities grow on demand (check_capacity, agreed over the ranks)
    dx = params.Lx / params.nx
    vmax = max(max(abs(d[0]) + s[0], abs(d[1]) + s[1]) for d, s in zip(w["drift"], w["spread"]))
    params.outbox_fraction = min(0.5, max(0.08, 1.6 * vmax * params.dt / dx / 8.0))
    sim = Sim(params)
    if w["init"] == "reference" and world == 1 and reference_ics:
        # the reference's own initial conditions (glibc rand() stream / position delta, src/particle.c:17-168)
        parts = init_particles(sample)
        for i, p in enumerate(parts):
            sim.set_particles(i, p["id"], p["x"], p["y"], p["ux"], p["uy"])
        data = "synthetic (the reference's host initialiser, bit-identical initial conditions, seed 138)"
    else:
        for i in range(nspecies):
            sim.init_beam(i, nps, id0=rank * nps, drift=w["drift"][i], spread=w["spread"][i], seed=138 + i)
        data = "synthetic (device initialiser: uniform positions per particle block, u = drift + U(-v, v))"
    env.bootstrap(sim)
    sim.pre_step()
    sim.sync()
    return sim, params, nps, data


def time_workload(env, name, steps, warmup, clocks=None, staged=True):
    """W warm-up steps, K timed steps (CUDA events on the simulation's stream, max over ranks), then
    a second pass with events around every stage, and the reference's separate stages."""
    w = WORKLOADS[name]
    world = env.world
    sim, params, nps, data = build_sim(env, name)
    nspecies = len(params.q)
    n_rank = nps * nspecies
    n_total = n_rank * world
    t_clk = time.perf_counter()
    sim.run(warmup)
    sim.timing(False)
    env.barrier()
    t0 = time.perf_counter()
    ms = sim.run_timed(steps)
    env.barrier()
    wall = time.perf_counter() - t0
    _, launches0 = sim.get_timing()
    ms = env.reduce(ms, "max")
    # per-stage device time over another K steps (events around every stage; a separate pass, because
    # the per-stage synchronisation serialises what the timed pass overlaps: the sum of the stages can
    # exceed ms_per_step)
    sim.timing(True)
    sim.run(steps)
    stage_ms, _ = sim.get_timing()
    sim.timing(False)
    st = None
    if staged and world == 1:
        sim.step_staged()            # allocates the per-particle E arrays: not timed
        sim.sync()
        sim.timing(True)
        ns = max(3, min(steps, 10))
        for _ in range(ns):
            sim.step_staged()
        sim.sync()
        st_ms, _ = sim.get_timing()
        sim.timing(False)
        st = {"steps": ns, "gather": st_ms["gather"] / (ns * nspecies), "push": st_ms["gather_push"] / (ns * nspecies),
              "deposit": st_ms["field_rho"] / ns}
    if clocks is not None:
        # nvidia-smi samples every 100 ms; K steps can be shorter: keep the same steps running until
        # ~1.5 s are covered (the same count on every rank: the steps contain collectives)
        reps = int(min(60, max(0, (1500.0 - (time.perf_counter() - t_clk) * 1e3) / max(ms, 1e-3))))
        reps = int(env.reduce(float(reps), "min"))
        for _ in range(reps):
            sim.run(steps)
        env.barrier()
    peak, peak_src = measured_peak()
    t_push = stage_ms["gather_push"] / (steps * nspecies)          # ms per k_gather_push launch
    t_dep = stage_ms["field_rho"] / steps
    gb = lambda nbytes, ms_: nbytes / (ms_ * 1e-3) / 1e9 if ms_ > 0 else 0.0
    ach = gb(BYTES_GATHER_PUSH * nps, t_push)
    per_gpu = n_rank * steps / (ms * 1e-3)
    roof = {"bound": "hbm", "kernel": "k_gather_push<2> (fused field gather + Boris push + exchange)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
            "frac_of_nominal_8TBs": ach / 8000.0,
            "algorithmic_bytes_per_particle": BYTES_GATHER_PUSH, "particles_per_launch": nps, "avg_launch_ms": t_push,
            "whole_step_frac_80B": per_gpu * BYTES_GATHER_PUSH / 1e9 / peak,
            "whole_step_frac_96B": per_gpu * (BYTES_GATHER_PUSH + BYTES_DEPOSIT) / 1e9 / peak,
            "stage_ms_per_step": {k: v / steps for k, v in stage_ms.items()},
            "stage_note": "stages timed in a separate pass with a synchronisation after each (their sum can exceed "
                          "ms_per_step, which overlaps the species' pushes on two streams)",
            "other_kernels": {
                "k_deposit (every species in one launch) + k_rho_assemble (16 B/particle)":
                    {"avg_launch_ms": t_dep, "achieved": gb(BYTES_DEPOSIT * n_rank, t_dep),
                     "frac": gb(BYTES_DEPOSIT * n_rank, t_dep) / peak}}}
    if st:
        roof["other_kernels"]["k_gather_push<0> (stage_plasma_E: gather, 32 B/particle)"] = {
            "avg_launch_ms": st["gather"], "achieved": gb(32.0 * nps, st["gather"]), "frac": gb(32.0 * nps, st["gather"]) / peak}
        roof["other_kernels"]["k_gather_push<1> (stage_plasma_r: push + exchange, 96 B/particle)"] = {
            "avg_launch_ms": st["push"], "achieved": gb(96.0 * nps, st["push"]), "frac": gb(96.0 * nps, st["push"]) / peak}
    res = {"value": n_total * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
           "n_rank": n_rank, "n_total": n_total, "nps": nps, "nspecies": nspecies, "params": params,
           "launches": int(launches0), "roofline": roof, "data": data, "wall": wall,
           "workload": workload_label(name, world),
           "grid": f"{params.nx}x{params.ny} cells in total, {params.nx}x{params.ny // world} per GPU",
           "capacity": {f"species{i}": sim.occupancy(i) for i in range(nspecies)}}
    return sim, res


def summary(res):
    """What a secondary workload contributes to the line."""
    r = res["roofline"]
    return {"workload": res["workload"], "grid": res["grid"], "particles": res["n_total"], "value": res["value"],
            "unit": UNIT, "ms_per_step": res["ms_per_step"], "steps": res["steps"], "warmup": res["warmup"],
            "data": res["data"], "gpu_launches": res["launches"],
            "roofline": {k: r[k] for k in ("kernel", "achieved", "peak", "unit", "frac", "avg_launch_ms",
                                           "whole_step_frac_80B", "whole_step_frac_96B", "stage_ms_per_step", "other_kernels")}}


def multi_rank_check(env, steps=12):
    """Product-only cross-check made before any timing with several ranks: the same small global
    problem (warm beam crossing the slab faces, two species, B along z) on the N ranks and, on every
    rank for itself, on one rank; this rank's slab of rho, phi, E and its particles (by id) must agree
    to 1e-12, the particle count and the total charge must be conserved."""
    import numpy as np
    from cpic_b200 import Sim, Params
    from cpic_b200.dist import partition, set_particles_collective, slab_rank
    world, rank = env.world, env.rank
    nx, nyl = 128, 32
    ny = nyl * world
    n = 12000 * world
    rng = np.random.default_rng(2024)
    L = (8.0, 8.0 * ny / nx)
    dt = 0.05
    dx = L[0] / nx
    parts = []
    for s, (drift, spread) in enumerate([((0.3, 0.9), 0.25), ((0.0, -0.2), 0.1)]):
        parts.append({"id": np.arange(n, dtype=np.int64), "x": rng.uniform(0, L[0], n), "y": rng.uniform(0, L[1], n),
                      "ux": (drift[0] + rng.uniform(-spread, spread, n)) * dx / dt,
                      "uy": (drift[1] + rng.uniform(-spread, spread, n)) * dx / dt, "uz": np.zeros(n)})
    common = dict(nx=nx, ny=ny, Lx=L[0], Ly=L[1], dt=dt, e0=2.0e3, B=(0.0, 0.0, 0.3), q=(-1.0, 1.0), m=(1.0, 4.0))
    one = Sim(Params(rank=0, nranks=1, device=env.local if env.cuda else 0, **common))
    for i, p in enumerate(parts):
        one.set_particles(i, p["id"], p["x"], p["y"], p["ux"], p["uy"])
    pm = Params(rank=rank, nranks=world, device=env.local if env.cuda else 0, **common)
    many = Sim(pm)
    set_particles_collective(many, partition(parts, pm, rank), env.dist, device=env.dev)
    env.bootstrap(many)
    one.pre_step()
    many.pre_step()
    worst, count_ok, crossed = {}, True, 0
    r0 = rank * nyl

    def compare(tag):
        nonlocal count_ok, crossed
        many.sync()
        one.sync()
        for k, rows in (("rho", slice(r0, r0 + nyl)), ("Ex", slice(r0, r0 + nyl)), ("Ey", slice(r0, r0 + nyl))):
            a, b = many.field(k)[:nyl], one.field(k)
            worst[k] = max(worst.get(k, 0.0), float(np.abs(a - b[rows]).max() / max(np.abs(b).max(), 1e-300)))
        a, b = many.field("phi_ghost"), one.field("phi")
        idx = [(r0 - 1 + j) % ny for j in range(nyl + 3)]
        worst["phi"] = max(worst.get("phi", 0.0), float(np.abs(a - b[idx]).max() / max(np.abs(b).max(), 1e-300)))
        total = 0
        for i in range(2):
            pa, pb = many.particles(i), one.particles(i)
            sel = slab_rank(pm, pb["y"]) == rank
            total += len(pa["id"])
            if len(pa["id"]) != int(sel.sum()) or not (pa["id"] == pb["id"][sel]).all():
                count_ok = False
                continue
            crossed += int((slab_rank(pm, parts[i]["y"][pb["id"][sel]]) != rank).sum())
            umax = max(np.abs(pb["ux"]).max(), np.abs(pb["uy"]).max())
            for k, scale in (("x", L[0]), ("y", L[1]), ("ux", umax), ("uy", umax)):
                worst[k] = max(worst.get(k, 0.0), float(np.abs(pa[k] - pb[k][sel]).max(initial=0.0) / scale))
        if int(env.reduce(float(total), "sum")) != 2 * n:
            count_ok = False
        # total charge: sum of rho over the slabs (ghost row excluded) against the single-rank sum
        qa = env.reduce(float(many.field("rho")[:nyl].sum()), "sum")
        qb = float(one.field("rho").sum())
        # relative to the charge of one sign (the plasma is neutral: the total itself is ~0)
        each = sum(abs(q) for q in common["q"]) * n / common["e0"]
        worst["charge"] = max(worst.get("charge", 0.0), abs(qa - qb) / each)

    compare("after sim_init")
    for it in range(steps):
        many.step()
        one.step()
        if it in (0, steps // 2, steps - 1):
            compare(f"iteration {it}")
    w = {k: env.reduce(v, "max") for k, v in sorted(worst.items())}
    ok = env.reduce(0.0 if (count_ok and all(v <= 1e-12 for v in w.values())) else 1.0, "max") == 0.0
    crossed = int(env.reduce(float(crossed), "sum"))
    many.close()
    one.close()
    return {"ok": bool(ok), "worst": max(w.values()), "tolerance": 1e-12, "by_quantity": w, "ranks": world,
            "steps": steps, "grid": f"{nx}x{ny}", "particles": 2 * n, "count_conserved": bool(count_ok),
            "particles_on_another_rank_than_at_start": crossed,
            "what": "N-rank run against the single-rank run of the same problem (product only, no oracle): slab rows "
                    "of rho, phi (with ghost rows), E_x, E_y relative to the global maximum, particles by id, total "
                    "particle count and total charge"}


def e2e_c_abi(env, sim, res, e_steps, banded=None):
    """The same steps with the particle state living in pinned HOST memory: every step uploads it,
    runs one sim_step through the C ABI and reads back the particles and the four grids."""
    import ctypes as C
    import numpy as np
    L = sim.L
    params, nspecies, n_rank, n_total = res["params"], res["nspecies"], res["n_rank"], res["n_total"]
    from cpic_b200._lib import check
    banded = env.world == 1 if banded is None else banded
    bands = int(min(16, max(4, n_rank * 48 // (256 << 20))))      # ~256 MB of particles per band
    nbytes = L.cpic_b200_banded_image_bytes(sim.h, bands) if banded else L.cpic_b200_image_bytes(sim.h)
    assert nbytes > 0, L.cpic_b200_last_error()
    host = L.cpic_b200_host_alloc(nbytes)
    assert host, "pinned allocation failed"
    fields = {k: np.empty(sim.field_shape(k)) for k in ("rho", "phi", "Ex", "Ey")}
    fbytes = sum(a.nbytes for a in fields.values())
    if banded:
        check(L.cpic_b200_banded_image_download(sim.h, host, nbytes, bands))
    else:
        check(L.cpic_b200_image_download(sim.h, host, nbytes))
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        if banded:
            # the image in bands of block rows: upload, push and download of the bands overlap (PCIe both ways at once)
            check(L.cpic_b200_step_host_banded(sim.h, host, nbytes))
        else:
            check(L.cpic_b200_image_upload(sim.h, host, nbytes))
            sim.step()
            check(L.cpic_b200_image_download(sim.h, host, nbytes))
        for k, a in fields.items():
            check(L.cpic_b200_get_field(sim.h, {"rho": 0, "phi": 1, "Ex": 2, "Ey": 3}[k], a.ctypes.data_as(C.c_void_p)))
    sim.sync()
    env.barrier()
    te = env.reduce(time.perf_counter() - t0, "max")
    L.cpic_b200_host_free(host)
    nb = (params.nx // 8) * (params.ny // env.world // 8)
    moved = n_rank * 48 + 8 * nspecies + 4 * nspecies * nb
    return {"value": n_total * e_steps / te, "unit": UNIT, "h2d_bytes_per_step": int(moved),
            "d2h_bytes_per_step": int(moved + fbytes), "steps": e_steps, "ms_per_step": te / e_steps * 1e3,
            "path": (f"C ABI (cpic_b200_step_host_banded: the image in {bands} bands of block rows per species, each uploaded, "
                     "pushed and downloaded as soon as its neighbours allow, on three streams; cpic_b200_get_field x 4)" if banded else
                     "C ABI (cpic_b200_image_upload, cpic_b200_step, cpic_b200_image_download, cpic_b200_get_field x 4)"),
            "note": "the whole particle state (x,y,ux,uy,uz,id of every particle) is uploaded from pinned host memory "
                    "before and downloaded after every sim_step, plus the four grids: host-owned particle lists, the "
                    "worst case of the drop-in"}


def e2e_plugin(steps=12, warmup=3):
    """The reference's own driver (src/cpic.c -> sim_run -> sim_step) with the four stage functions
    served by dropin/cpic_b200_stages.c over libcpic_b200.so, on configs[1] (the reference host lists
    need 112 B per particle, so this is the size a host can hold): the reference's TIMER_ITERATION per
    step, in both coherence modes of the binding."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_cpic")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/dropin_cpic is not built (needs the reference sources at build time)"}
    w = WORKLOADS["A"]
    conf = scaled_conf(os.path.join(ROOT, "conf", w["conf"]), w, w["nps"], cycles=warmup + steps, stop_sem=1e-30)
    n = w["nps"] * len(w["drift"])
    out = {"workload": workload_label("A", 1), "steps": steps, "warmup": warmup,
           "path": "reference main -> sim_step -> stage_field_E/plasma_E/plasma_r/field_rho(sim_t *) "
                   "(dropin/cpic_b200_stages.c) -> C ABI; time = the reference's TIMER_ITERATION (src/sim.c:440-479)"}
    for mode in ("eager", "lazy"):
        env = dict(os.environ, CPIC_B200_SYNC=mode)
        try:
            r = subprocess.run([exe, "-q", conf], env=env, capture_output=True, text=True, timeout=900)
            last = [float(m.group(1)) for m in re.finditer(r"^stats iter=\d+ last=([0-9.eE+-]+)", r.stdout, re.M)]
            if r.returncode != 0 or len(last) < warmup + steps:
                out[mode] = {"error": (r.stderr or r.stdout)[-300:]}
                continue
            t = last[warmup:warmup + steps]
            mean = sum(t) / len(t)
            out[mode] = {"value": n / mean, "unit": UNIT, "ms_per_step": mean * 1e3}
        except Exception as exc:
            out[mode] = {"error": repr(exc)}
    out["modes"] = {"eager": "host particle lists and grids refreshed after every stage that changes them (what the "
                             "reference's tests, which read list.b->p[] after sim_step, need)",
                    "lazy": "grids refreshed every step, particle lists only on request (cpic_b200_dropin_sync)"}
    return out


def f1_report():
    """the unmodified reference drops deposits when two lanes of a 4-particle pack share a
    cell (src/simd_avx2.h:226-249). For configs[1] after sim_init: the census of such packs and the
    rho of the unmodified build against the accumulate-correct one (which is what the GPU matches)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        from _refbind import RefSim, ref_available
        if not (ref_available("ref") and ref_available("ref_acc")):
            return {"unavailable": "oracle/_ref is not built"}
        w = WORKLOADS["A"]
        conf = scaled_conf(os.path.join(ROOT, "conf", w["conf"]), w, w["nps"])
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            a, b = RefSim(conf, "ref"), RefSim(conf, "ref_acc")
            packs, lost = a.collision_census()
            ra, rb = a.field("rho"), b.field("rho")
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        d = np.abs(ra - rb)
        return {"workload": workload_label("A", 1), "packs_with_a_collision": int(packs), "lost": int(lost),
                "particles": w["nps"] * len(w["drift"]), "nodes_that_differ": int((d > 1e-12 * np.abs(rb).max()).sum()),
                "rho_rel_diff": float(d.max() / np.abs(rb).max()),
                "note": "after sim_init; rho of the unmodified reference against its accumulate-correct variant "
                        "(oracle/_ref/libcpic_ref_acc.so: vmat_add_xy made lane-serial), relative to max|rho|. The GPU "
                        "deposits every particle: it matches the accumulate-correct variant to 1e-12 "
                        "(tests/test_gpu_reference.py) and differs from the unmodified one by this much"}
    except Exception as exc:
        return {"error": repr(exc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cpic_b200")
    ap.add_argument("--workload", default="D", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the main workload: no other_workloads, "
                    "target_config, multi_rank_check, f1, plugin e2e")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    # stdout carries the one JSON line and nothing else: whatever the libraries print on the way
    # (NCCL's version banner with NCCL_DEBUG=VERSION, for one) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    from cpic_b200._lib import lib as _cpic_lib
    if "sm_100a" not in _cpic_lib().cpic_b200_version().decode() and not TINY:
        raise SystemExit("bench.py measures the CUDA library only")
    env = Env(args.gpus)
    world, rank = env.world, env.rank
    extras = not args.no_extras

    check = None
    if world > 1 and extras:
        check = multi_rank_check(env)
        dbg(f"multi_rank_check {check['ok']} {check['worst']:.2e}")

    # ---- the main workload
    clocks = ClockSampler(env.local)
    if rank == 0 and env.cuda:
        clocks.start()
    sim, res = time_workload(env, args.workload, args.steps, args.warmup, clocks=clocks)
    clk = clocks.stop() if (rank == 0 and env.cuda) else None
    if clk is not None:
        clk["window"] = "warm-up + timed region + repeats of the same steps of the main workload"
    roofline = res["roofline"]
    if world == 1 and args.workload in ("A", "D"):
        csvp = os.path.join(ROOT, "profiles", f"r2_ncu_full_summary_{args.workload}.csv")
        roofline["traffic"] = ncu_traffic(path=csvp)
        roofline["traffic_source"] = (f"profiles/r2_ncu_full_summary_{args.workload}.csv (ncu --set full of this workload, "
                                      "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the species)")
        if roofline["traffic"]:
            # what the kernel really moves against the same peak (the algorithmic figure is `frac`)
            roofline["traffic_over_algorithmic"] = roofline["traffic"] / (BYTES_GATHER_PUSH * res["nps"])
            roofline["dram_frac"] = roofline["traffic"] / (roofline["avg_launch_ms"] * 1e-3) / 1e9 / roofline["peak"]
        td = ncu_traffic(kernel="k_deposit", path=csvp)
        for k, v in roofline["other_kernels"].items():
            if k.startswith("k_deposit") and td:
                v["traffic"] = td
                v["traffic_over_algorithmic"] = td / (BYTES_DEPOSIT * res["n_rank"])
                v["dram_frac"] = td / (v["avg_launch_ms"] * 1e-3) / 1e9 / roofline["peak"]
    else:
        roofline["traffic"] = None

    e2e = None
    if not args.no_e2e:
        # the particle image of the default workload is 12 GB per GPU: three steps keep the run short
        e_steps = 3 if res["n_rank"] > 50_000_000 else max(3, min(args.steps, 10))
        try:
            e2e = e2e_c_abi(env, sim, res, e_steps)
        except Exception as exc:
            if world > 1:
                raise            # the ranks must stay in step
            # (a band of the image outgrown, no pinned memory for the bands' slack ...): the serial image path
            dbg(f"banded e2e failed: {exc!r}")
            e2e = e2e_c_abi(env, sim, res, e_steps, banded=False)
            e2e["note"] += f"; the banded path failed ({exc!r})"
    sim.close()

    # ---- the other configurations of BASELINE.json, in the same driver record
    others, target = None, None
    if extras and world == 1:
        others = {}
        for name in ("A", "2s"):
            if name == args.workload:
                continue
            try:
                s2, r2 = time_workload(env, name, args.steps, args.warmup, staged=(name == "A"))
                others[name] = summary(r2)
                s2.close()
            except Exception as exc:      # a secondary workload must never cost the line
                others[name] = {"error": repr(exc)}
    # (two GPUs would need 5e8 particles and their exchange regions each: beyond 180 GB with this workload's
    # very mobile particles -- up to 6.4 cells per step)
    if extras and world >= 4 and args.workload != "C":
        try:
            s2, r2 = time_workload(env, "C", max(5, min(args.steps, 10)), 3, staged=False)
            target = summary(r2)
            target["north_star"] = ">= 1e10 particle-steps/s on 8 B200 for 4096x4096 cells and 1e9 particles"
            s2.close()
        except Exception as exc:
            target = {"error": repr(exc)}

    cpu, f1, plugin = None, None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, cpu = cpu_reference(args.workload, 5, 2, budget_s=25.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "ms_per_step") if k in cpu}
        if extras and not TINY:
            f1 = f1_report()
    if rank == 0 and world == 1 and extras and not args.no_e2e and not TINY:
        plugin = e2e_plugin()
    if e2e is not None and plugin is not None:
        e2e["plugin"] = plugin

    if rank == 0:
        p = res["params"]
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if WORKLOADS[args.workload]["strong"] else "weak", "vs_baseline": None,
                "dtype": "f64", "data": res["data"],
                "config": {"workload": res["workload"], "grid": res["grid"], "particles": res["n_total"],
                           "species": res["nspecies"], "B": list(p.B), "dt": p.dt,
                           "parallelism": f"{world} Y-slab(s), one per GPU",
                           "l2": "particle state per GPU (%.0f MB) exceeds the 126 MB L2" % (res["n_rank"] * 48 / 1e6),
                           "block_capacity": res["capacity"]},
                "clocks": clk, "e2e": e2e, "gpu_launches": res["launches"],
                "roofline": roofline, "cpu_baseline": cpu, "wall_s_timed_region": res["wall"]}
        if check is not None:
            line["multi_rank_check"] = check
        if others is not None:
            line["other_workloads"] = others
        if target is not None:
            line["target_config"] = target
        if f1 is not None:
            line["f1"] = f1
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    env.close()


if __name__ == "__main__":
    main()
