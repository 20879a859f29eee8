#!/usr/bin/env python
"""Benchmark of the collision step (BASELINE.json metric: cloth collision step ms & CCD pairs/sec at
1 M tris, 1/2/4/8 B200 vs host CPU).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path on host cores

A step = one resolveCollision over the 1 M-triangle layered-cloth scene (config 4), always started from the same
state.  metric = CCD pairs/s = candidate element pairs handed to the narrow phase by the timed steps (all CCD passes) /
step time; the reference-equivalent callback count and rate are reported beside it (config.ccd_pairs_reference_equivalent,
value_reference_equivalent); ms_per_step is the other half of the headline.  `value` times the step with x_old / x_new
already in HBM (CUDA events on the library's stream); `e2e` times the drop-in call (clsn_step_host) with pinned host
buffers, copies inside the timed region; `e2e_api_default` the same call with the impact-zone fail-safe and strain
limiting on (what the reference's resolveCollision does on this input); `roofline` = the dominant phase against the
measured FP64 issue peak (bound "fp64") or the HBM copy peak, `roofline_memory` = the memory-bound passes against the
HBM copy peak; `cpu_baseline` = the compiled reference on a bounded 50 K-triangle sample.
N > 1 (torchrun): the in-library multi-GPU step (csrc/dist.cuh); --exchange owner|gather selects the python-driven protocols.
--impl reference: the compiled reference on config 4 ITSELF, one step (~10 min on one core), see run_reference().

The one JSON line on stdout is the contract; everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from collision_b200 import scenes  # noqa: E402

METRIC = "ccd_pairs_per_sec"
UNIT = "pairs/s"


# The reference prints its progress on std::cout.  Keep the real stdout for the one JSON line and
# point fd 1 at stderr for everything else (C++ streams included).
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload(name: str):
    if name == "config4":
        return scenes.layered_cloth(8, 251), "layered_cloth 8x251^2: 1 000 000 tris, 504 008 verts, dt=1e-3 (config 4)"
    if name == "config3":
        return scenes.drape(256, 5), "drape 256^2 sheet on static icosphere: 150 530 tris (config 3)"
    if name == "sample":
        return scenes.layered_cloth(8, 57), "layered_cloth 8x57^2: 50 176 tris (bounded sample of config 4)"
    if name == "sample_ref":
        return scenes.layered_cloth(8, 33), "layered_cloth 8x33^2: 16 384 tris (bounded sample of config 4)"
    if name == "config5":
        return scenes.cloth_spheres(8, 501, 4, 3), "cloth_spheres 8x501^2 + 64 icospheres: 4 081 920 tris (config 5)"
    if name == "config5s":
        return scenes.cloth_spheres(8, 126, 4, 3), "cloth_spheres 8x126^2 + 64 icospheres: 331 920 tris (1/16-area member of the config-5 family)"
    if name == "tiny":
        return scenes.layered_cloth(4, 33), "layered_cloth 4x33^2: 8 192 tris (smoke)"
    raise SystemExit(f"unknown workload {name}")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self, first_timed: int = 0):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # a very short timed region can fall between two nvidia-smi samples: then use every sample taken
        # since the sampler started (the untimed spin-up steps run the same step under the same load)
        window = "timed"
        lines = self.lines[first_timed:]
        if len(lines) < 3:
            lines, window = self.lines, "spin-up + timed (timed region shorter than 3 samples)"
        for ln in lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": window}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


# ----------------------------------------------------------------------------- reference arm
def time_reference(scene, steps, warmup):
    """The reference's own CPU implementation (oracle/_ref when it was built from /root/reference,
    else the C restatement) on the host cores: single-threaded, like the reference."""
    from oracle import port, ref
    x0, v0 = scene.x.copy(), scene.vel.copy()
    xn0 = x0 + scene.dt * v0
    times, pairs = [], []
    if ref.available():
        kind = "reference"
        r = ref.RefSolver(scene)
        for it in range(warmup + steps):
            r.set_state(x0, xn0, v0)
            r.assemble(scene.dt)
            r.record(False)
            # exactly the work clsn_resolve does: resolveCollision (dcollid.cpp:317-362) without reduceSuperelast and
            # without the computeImpactZone fail-safe that the reference enters after 5 unresolved CCD passes
            t0 = time.perf_counter()
            r.phase(ref.PH_AVG_VELOCITY)
            r.phase(ref.PH_PROXIMITY_DETECT)
            r.phase(ref.PH_APPLY)
            n_prox = r.num_callbacks()
            for _ in range(5):                       # MAX_ITER, dcollid.cpp:433
                n_true = r.phase(ref.PH_COLLISION_DETECT)
                r.phase(ref.PH_APPLY)
                if n_true == 0:
                    break
            r.phase(ref.PH_BOUNDARY)
            r.phase(ref.PH_FINAL_POSITION)
            r.phase(ref.PH_FINAL_VELOCITY)
            t1 = time.perf_counter()
            if it >= warmup:
                times.append(t1 - t0)
                pairs.append(r.num_callbacks() - n_prox)
    else:
        kind = "port"
        port.set_libm(port.LIBM_NATIVE)
        o = port.OracleSolver(scene, impact_zones=False, strain_limiting=False)
        for it in range(warmup + steps):
            o.set_state(x0, xn0)
            v = v0.copy()
            t0 = time.perf_counter()
            st = o.resolve(v)
            t1 = time.perf_counter()
            if it >= warmup:
                times.append(t1 - t0)
                pairs.append(sum(st[9:9 + st[1]]))
    return kind, times, pairs


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref = the unmodified reference sources) on the SAME
    workload as the CUDA arm -- config 4, 1 000 000 triangles.  The reference is single-threaded and needs ~8-10 minutes for
    that one step (404 s in the survey's probe), so exactly ONE step is timed whatever --steps says ("steps": 1
    in the line), without a warm-up step; --sample-reference falls back to a bounded sample of the same generator.
    The measurement does not depend on --gpus: on one box it is taken once and reused for the other N of a scaling run
    (cache file in the system temp dir, stated in `cpu_baseline.sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tempfile
    name = args.workload if not args.sample_reference else ("sample_ref" if args.sample == "sample" else args.sample)
    scene, desc = workload(name)
    cache = os.path.join(tempfile.gettempdir(), f"collision_b200_reference_arm_{name}.json")
    rec = None
    if os.path.exists(cache) and not args.no_cache:
        try:
            rec = json.load(open(cache))
            rec["cached"] = True
        except Exception:
            rec = None
    if rec is None:
        full = not args.sample_reference
        steps, warmup = (1, 0) if full else (args.steps, args.warmup)
        t0 = time.time()
        kind, times, pairs = time_reference(scene, steps, warmup)
        rec = {"kind": kind, "times": times, "pairs": pairs, "steps": steps, "warmup": warmup, "when": t0, "cached": False}
        try:
            json.dump(rec, open(cache, "w"))
        except Exception:
            pass
    total_t = sum(rec["times"])
    value = sum(rec["pairs"]) / total_t
    note = ("measured in this invocation" if not rec["cached"] else
            f"measured once on this box {time.time() - rec['when']:.0f} s ago by the same command and reused (independent of --gpus)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": rec["steps"],
        "warmup": rec["warmup"], "ms_per_step": 1e3 * total_t / len(rec["times"]), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "parallelism": "one host thread (the reference is single-threaded)",
                   "ccd_pairs_per_step": int(sum(rec["pairs"]) / len(rec["times"])),
                   "scope": "resolveCollision hot loop: avgVel, proximity pass, <=5 CCD passes, boundary, final position; "
                            "strain limiting and the impact-zone fail-safe excluded on both arms",
                   "note": "one step is all a run of a few minutes holds: requested --steps/--warmup are not honoured"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": rec["kind"],
                         "sample": f"{desc}; {len(rec['times'])} timed step(s), {total_t:.1f} s; {note}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- roofline bookkeeping
def algorithmic_bytes(scene, st):
    """SURVEY 8(d) per-unit figures x the units one step processed, per phase (bytes)."""
    V, T = scene.V, scene.T + scene.B
    passes = [st["proximity"]] + st["ccd"]
    P = sum(p["candidates"] for p in passes)
    Pt = sum(p["pairs_tested"] for p in passes)
    C = sum(p["contacts"] for p in passes)
    K = sum(p["contributions"] for p in passes)
    F = sum(p.get("features", 0) for p in passes)
    n = len(passes)
    Fc = sum(p.get("coplanar", 0) for p in passes)
    Fx = sum(p.get("exact_solves", 0) for p in passes[1:])
    F0 = passes[0].get("features", 0)
    Cc = sum(p["contacts"] for p in passes[1:])
    staged = Fx == Fc          # pipeline 0 solves every feature: exact_solves == coplanar
    roots_b = 24 * (F - F0) + 48 * Fc if staged else 24 * (F - F0) + 2 * 24 * Fx + 32 * Cc   # work list in; undecided list out + in; hit list out
    contact_b = 48 * Fc + 24 * F0 + 64 * K if staged else 32 * Cc + 24 * F0 + 64 * K          # records / hit list in, impulse records out
    bytes_ = {
        "avgvel": 72 * V,
        "build": (48 * V + 12 * T + 8 * T) + 64 * T + 2 * 16 * T,             # morton + 4-pass sort + element gather (only when the tree is rebuilt)
        "refit": n * (16 * T + 48 * V + 48 * T + 8 * T),                      # sorted elements, vertices once, exact leaf boxes, 64-B node per 8 leaves
        "traverse": n * (48 * T + 8 * T) + 8 * Pt,                            # leaf boxes + leaf-level nodes once, pairs out
        "cull": 8 * Pt + n * (48 * V + 12 * T) + 24 * F,                      # pairs in, vertex data once, work list out
        "roots": roots_b,
        "contact": contact_b,
        "reduce": 2 * 64 * K + n * 80 * V,                                    # records grouped + read, apply per vertex
        "finalize": (48 + 73) * V,
    }
    # Algorithmic FP64 operations (adds + multiplies of the reference's own expressions): coplanarity-cubic
    # coefficients 89, classifier 35, monic form + discriminant 25, positions at one time 24, PointToTri / EdgeToEdge ~115,
    # isCoplanar as the reference evaluates it 170 (libm calls not counted).
    Fbox = sum(p.get("box_survivors", 0) for p in passes[1:])
    Fccd = F - F0
    flops = {
        "cull": 124 * Fbox,
        "roots": (89 + 25 + 24 + 115) * Fccd + (170 + 115) * Fx,
        "contact": (24 + 115 + 60) * Cc + 115 * F0,
    }
    return bytes_, flops, dict(P=P, Pt=Pt, Fbox=sum(p.get("box_survivors", 0) for p in passes), F=F, Fcop=Fc, Fexact=Fx, C=C, K=K, passes=n)


# ----------------------------------------------------------------------------- CUDA arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from collision_b200.solver import CollisionSolver3d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the collision step has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    scene, desc = workload(args.workload)
    # the timed step is resolveCollision's hot loop; the impact-zone fail-safe (host-assisted, only entered when
    # 5 CCD passes leave collisions) is excluded on both arms, like strain limiting -- see config.scope
    solver = CollisionSolver3d(device=local, impact_zones=False, strain_limiting=False)
    CollisionSolver3d.set_params_from(scene.params)
    solver.assembleFromInterface(scene, scene.dt)
    if args.pipeline is not None:
        solver.set_pipeline(args.pipeline)
    # per-phase CUDA-event marks for the roofline figures: ON during the device-timed steps (they are part of what `value`
    # times), OFF -- the library's default -- during the end-to-end steps
    solver.set_phase_timing(not args.no_phase_marks)
    x_old = np.ascontiguousarray(scene.x)
    x_new = np.ascontiguousarray(scene.x_new())
    d_xo = torch.from_numpy(x_old).to(dev)
    d_xn = torch.from_numpy(x_new).to(dev)
    stepper = None
    if world > 1:
        if args.exchange == "library":
            # the in-library multi-GPU step (csrc/dist.cuh): same calls as on one GPU from here on
            from collision_b200.dist import enable_library_exchange
            enable_library_exchange(solver)
        else:
            from collision_b200.dist import DistributedSolver
            stepper = DistributedSolver(solver, mode=args.exchange)

    def one_step():
        solver.upload_device(d_xo.data_ptr(), d_xn.data_ptr())
        return stepper.resolve_device() if stepper else solver.resolve_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        solver.synchronize()

    # the metric's numerator is the reference-equivalent number of CCD narrow-phase callbacks of this
    # workload: counted once with the full traversal, then the timed steps may prune (identical results)
    solver.set_exact_stats(True)
    st_exact = one_step()
    solver.set_exact_stats(False)
    for _ in range(max(args.warmup, 3)):
        st = one_step()
    sampler = ClockSampler(local)
    sampler.start()
    # keep the GPU under load for ~0.8 s while nvidia-smi spins up (untimed).  The number of steps must be the SAME on every
    # rank -- a step contains collectives -- so it is derived from the slowest rank's step time, not from each rank's clock.
    t_one = time.perf_counter()
    one_step()
    t_one = torch.tensor([time.perf_counter() - t_one], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_one, op=dist.ReduceOp.MAX)
    n_spin = 0 if args.no_spin else max(1, min(400, int(0.8 / max(float(t_one.item()), 1e-3))))
    for _ in range(n_spin):
        one_step()
    barrier()
    n_before = len(sampler.lines)  # samples from here on fall inside the timed region
    solver.launch_count(reset=True)
    solver.timer_start()
    t0 = time.perf_counter()
    phase_ms = {}
    for _ in range(args.steps):
        st = one_step()
        for k, v in st.get("ms_phase", {}).items():
            phase_ms[k] = phase_ms.get(k, 0.0) + v
    ms = solver.timer_stop()
    log("per-pass stats of the last step:", json.dumps({"proximity": st["proximity"], "ccd": st["ccd"]}))
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = solver.launch_count()
    clocks = sampler.stop(n_before)

    # ---- e2e through the drop-in call with pinned host buffers (N = 1: the C ABI call; N > 1: host
    # upload + distributed step + host download)
    solver.set_phase_timing(False)
    h_xo = torch.from_numpy(x_old).pin_memory().numpy()
    h_xn = torch.from_numpy(x_new).pin_memory().numpy()
    h_out = torch.empty(x_new.shape, dtype=torch.float64).pin_memory().numpy()
    h_vel = torch.from_numpy(scene.vel.copy()).pin_memory().numpy()

    def one_step_host():
        if stepper is None:   # one GPU, or the in-library multi-GPU step: the drop-in call itself
            return solver.resolveCollision(h_xo, h_xn, h_vel, x_out=h_out)
        solver.upload(h_xo, h_xn)
        stepper.resolve_device()
        return solver.download()

    one_step_host()
    np.copyto(h_vel, scene.vel)
    barrier()
    te = time.perf_counter()
    for _ in range(args.steps):
        # h_vel is write-only for the library (vel = avgVel where has_collsn), so every step sees the same inputs
        # without a 12 MB host-side reset inside the timed region
        one_step_host()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - te)

    # ---- aggregate over ranks: time = max, pairs = every rank's slice
    # Two numerators ("candidate element pairs handed to the narrow phase"):
    #   processed             the pairs the timed steps really found and handed on (from CCD pass 2 on, pairs none of whose
    #                         points changed are skipped at traversal time: identical results, fewer pairs) -> `value`
    #   reference-equivalent  the callbacks the reference makes on the same step (counted once with the full traversal)
    ccd_pairs_ref = sum(p["candidates"] for p in st_exact["ccd"])
    ccd_pairs = sum(p["candidates"] for p in st["ccd"])
    lib_dist = world > 1 and stepper is None     # the library's statistics are already global then
    div = world if lib_dist else 1
    agg = torch.tensor([ms, e2e_ms, wall_ms, float(ccd_pairs) / div, float(ccd_pairs_ref) / div], dtype=torch.float64, device=dev)
    if world > 1:
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms, wall_ms = float(mx[0]), float(mx[1]), float(mx[2])
        ccd_pairs, ccd_pairs_ref = int(round(sm[3].item())), int(round(sm[4].item()))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    step_ms = ms / args.steps
    value = ccd_pairs / (step_ms * 1e-3)
    e2e_value = ccd_pairs / (e2e_ms / args.steps * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "parallelism": (
                       "one GPU" if world == 1 else
                       f"replicated mesh+tree, {world}-way query slices, owner-computes: " + (
                           "records stored into the owner's memory over NVLink from inside the kernels, NCCL all-reduce of the "
                           "pass counters + all-gather of avgVel, no host read-back (csrc/dist.cuh)" if lib_dist else
                           f"python-driven NCCL exchange ({args.exchange})")),
                   "l2": "no explicit flush: the step's working set (vertex state, BVH, pair and record buffers) is "
                         "several times the 126 MB L2", "ccd_passes": st["n_ccd_passes"],
                   "ccd_pairs_per_step": ccd_pairs, "ccd_pairs_reference_equivalent": ccd_pairs_ref,
                   "numerator": "CCD candidate pairs found and handed to the narrow phase by the timed steps; the reference "
                                "makes ccd_pairs_reference_equivalent callbacks on the same step (pairs untouched since the "
                                "previous pass are skipped here, results identical)",
                   "still_colliding": bool(st["still_colliding"]),
                   "pipeline": "fast path + exact solve of the undecided features" if any(
                       p.get("exact_solves", 0) != p.get("coplanar", 0) for p in st["ccd"]) else "staged exact solve",
                   "scope": "resolveCollision hot loop: avgVel, proximity pass, <=5 CCD passes, boundary, final position; "
                            "strain limiting and the impact-zone fail-safe excluded on both arms"},
        "step_ms": step_ms, "wall_ms_per_step": wall_ms / args.steps,
        "value_reference_equivalent": ccd_pairs_ref / (step_ms * 1e-3),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                # whole job: with the in-library multi-GPU step every rank uploads only its 1/N share of x_old / x_new (the ranks
                # all-gather the rest over NVLink) and downloads the complete result (x, avgVel, has_collsn)
                "h2d_bytes_per_step": int(2 * x_old.nbytes) * (1 if (world == 1 or lib_dist) else world),
                "d2h_bytes_per_step": int(2 * x_old.nbytes + scene.V) * world},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    # ---- roofline of the dominant phase, measured live (CUDA events between kernel groups)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    if world == 1 and sum(phase_ms.values()) > 0.0:   # (--no-phase-marks: no marks were recorded, no per-phase figures)
        bytes_, flops_, units = algorithmic_bytes(scene, st)
        per_step = {k: v / args.steps for k, v in phase_ms.items()}
        kernels = {}
        for k, b in bytes_.items():
            t = per_step.get(k, 0.0)
            if t > 1e-3 * step_ms:   # a phase that did not run in the timed steps (the tree is kept across steps) leaves ~0
                kernels[k] = {"ms": t, "share": t / step_ms, "alg_bytes": int(b), "gbs": b / (t * 1e-3) / 1e9,
                              "frac_hbm": b / (t * 1e-3) / 1e9 / peak}
        fp64_peak, fp64_src = 18.48e12, "fallback: 148 SMs x 64 lanes x 1.965 GHz (profiles/r2_fp64_peak.json, measured on this pool)"
        try:
            r = subprocess.run([os.path.join(ROOT, "tools", "fp64_peak")], capture_output=True, text=True, timeout=30)
            fp64_peak = float(json.loads(r.stdout.strip().splitlines()[-1])["fp64_inst_per_s_nofma"])
            fp64_src = "measured in this run (tools/fp64_peak: DADD/DMUL issue rate, --fmad=false)"
        except Exception:
            pass
        for k, f in flops_.items():
            if k in kernels:
                kernels[k]["alg_fp64_ops"] = int(f)
                kernels[k]["fp64_ops_per_s"] = f / (kernels[k]["ms"] * 1e-3)
                kernels[k]["frac_fp64"] = kernels[k]["fp64_ops_per_s"] / fp64_peak
        dom = max(kernels, key=lambda k: kernels[k]["ms"])
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        if dom in flops_:
            # the dominant phase is the FP64 narrow phase: its roofline is the DADD/DMUL issue rate (no FMA: every expression
            # rounds like the reference's), not HBM
            line["roofline"] = {"bound": "fp64", "kernel": dom, "achieved": kernels[dom]["fp64_ops_per_s"] / 1e12, "peak": fp64_peak / 1e12,
                                "unit": "TFLOP/s", "frac": kernels[dom]["frac_fp64"], "traffic": traffic.get(dom),
                                "peak_source": fp64_src, "hbm_frac": kernels[dom]["frac_hbm"],
                                "note": "algorithmic FP64 adds + multiplies of the reference's expressions (no FMA contraction) per "
                                        "second of the phase, against the measured FP64 issue peak; the HBM fraction of the same phase is hbm_frac"}
        else:
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                                "frac": kernels[dom]["frac_hbm"], "traffic": traffic.get(dom), "peak_source": peak_src}
        # the memory-bound passes the north star asks about: build / refit / impulse reduction against the HBM copy peak
        mem = {k: kernels[k] for k in ("refit", "reduce", "avgvel", "finalize") if k in kernels}
        if mem:
            worst = min(mem, key=lambda k: mem[k]["frac_hbm"])
            line["roofline_memory"] = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                                       "passes": {k: {"achieved": v["gbs"], "frac": v["frac_hbm"], "ms": v["ms"],
                                                      "traffic": traffic.get(k)} for k, v in mem.items()},
                                       "lowest": worst}
        line["kernels"] = kernels
        line["units"] = units
    # ---- the drop-in call with the API defaults (impact-zone fail-safe + strain limiting ON, as resolveCollision runs them):
    # config 4 still collides after the 5 CCD passes, so the reference -- and the host mirrors by default -- go on into
    # computeImpactZone; timed separately because the fail-safe is host-assisted and outside the north star's hot loop
    if world == 1 and not args.no_api_default:
        try:
            full = CollisionSolver3d(device=local, impact_zones=True, strain_limiting=True)
            full.assembleFromInterface(scene, scene.dt)
            np.copyto(h_vel, scene.vel)
            full.resolveCollision(h_xo, h_xn, h_vel, x_out=h_out)
            n_full = max(2, min(args.steps, 5))
            tf = time.perf_counter()
            for _ in range(n_full):
                full.resolveCollision(h_xo, h_xn, h_vel, x_out=h_out)
            tf = 1e3 * (time.perf_counter() - tf) / n_full
            fs = full.last_stats
            line["e2e_api_default"] = {"ms_per_step": tf, "value": ccd_pairs / (tf * 1e-3), "unit": UNIT, "steps": n_full,
                                       "impact_zone_iterations": fs["zone_iterations"], "zones": fs["zones"],
                                       "strain_sweeps": fs["strain_sweeps"],
                                       "note": "clsn_step_host with clsn_set_impact_zones(1) and clsn_set_strain_limiting(1): what the "
                                               "reference's resolveCollision() does on this input, host buffers, copies included"}
            full.close()
        except Exception as e:  # noqa: BLE001
            line["e2e_api_default"] = {"error": str(e)[:200]}
    # ---- CPU baseline on a bounded sample (rank 0, N = 1)
    if world == 1 and not args.no_cpu:
        sc_s, desc_s = workload(args.sample)
        kind, times, pairs = time_reference(sc_s, 1, 1)
        line["cpu_baseline"] = {"value": sum(pairs) / sum(times), "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"{desc_s}; 1 warm-up + 1 timed step, {sum(times):.1f} s",
                                "host_cores_available": os.cpu_count()}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config4")
    ap.add_argument("--sample", default="sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-spin", action="store_true", help="profiling runs: skip the untimed spin-up steps")
    ap.add_argument("--pipeline", type=int, default=None, help="CCD narrow-phase pipeline (default: the library's)")
    ap.add_argument("--no-phase-marks", action="store_true",
                    help="device-timed steps without the per-phase event marks (no kernels{} / roofline in the line)")
    ap.add_argument("--sample-reference", action="store_true",
                    help="--impl reference: time a bounded sample (16 K triangles) instead of the full workload")
    ap.add_argument("--no-cache", action="store_true", help="--impl reference: measure even if this box has a cached measurement")
    ap.add_argument("--no-api-default", action="store_true", help="skip the extra e2e figure with impact zones + strain limiting on")
    ap.add_argument("--exchange", default="library", choices=["library", "owner", "gather"],
                    help="multi-GPU exchange: inside the library (default) or the python-driven NCCL protocols")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
