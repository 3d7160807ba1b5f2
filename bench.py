#!/usr/bin/env python
"""Benchmark of the tncontract hot path on B200.

Metric (BASELINE.json): MPO.MPS apply + svd-compress sweeps/s at N=100, d=2,
chi=512, complex128 (config 3).  One STEP = one sweep through the public API:

    phi = contract_mps_mpo(psi, H)      # 100 fused apply+consolidate kernels
    phi.svd_compress(chi=512)           # 99 QR + R-absorb, 99 SVD + V/S-absorb

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* value : sweeps/s with psi and H already resident in HBM (device-timed with
          CUDA events, max over ranks).  For N > 1 every rank sweeps its own
          independent network (weak scaling, no data-path collective).
* e2e   : the same sweep with HOST inputs: site tensors copied from pinned host
          memory every step, the compressed MPS read back to the host.
* roofline: one extra profiled pass (tnb_profile_*: CUDA events around every
          kernel class) names the dominant kernel class and its achieved rate;
          the FP64 tensor peak is cuBLAS ZGEMM/DGEMM measured in this run
          (MEASURED_PEAKS.json has no FP64 figure), HBM peak from that file.
* --impl reference: the UNMODIFIED reference (baseline/_ref; the oracle port if
          that is missing) on all host threads: whole sweeps on the same inputs,
          nothing extrapolated (baseline/ref_arm.py).
* cpu_baseline: a bounded sample (8 bulk sites) of the same CPU path inside the GPU
          arm's run, scaled by the nominal flop profile and labelled as such.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MPO.MPS apply+svd-compress sweeps/s (N=100,d=2,chi=512)"
KCLASS = ["gemm", "jacobi_round", "qr_panel", "permute", "mps_mpo_site", "elementwise"]


# ----------------------------------------------------------------------------- workload
def bond_profile(n, d, chi):
    return [min(chi, d ** i, d ** (n - i)) for i in range(n + 1)]


def make_host_sites(n, d, chi, seed):
    """cfg 3 input: complex standard-normal site tensors with the
    bond profile a QR-canonised chi-capped MPS has."""
    rng = np.random.default_rng(seed)
    bonds = bond_profile(n, d, chi)
    return [(rng.standard_normal((d, bonds[i], bonds[i + 1])) +
             1j * rng.standard_normal((d, bonds[i], bonds[i + 1]))) for i in range(n)]


def tfi_w(J=1.0, h=0.5):
    I, Z, X = np.eye(2), np.diag([1.0, -1.0]), np.array([[0.0, 1.0], [1.0, 0.0]])
    W = np.zeros((3, 3, 2, 2), dtype=complex)
    W[0, 0], W[1, 0], W[2, 0], W[2, 1], W[2, 2] = I, Z, -h * X, -J * Z, I
    return W


def nominal_flops(n, d, chi, D):
    """LAPACK-model flop count of one sweep and of one bulk site."""
    b = bond_profile(n, d, chi)

    def qr(m, k):
        k2 = min(m, k)
        return 4 * 2 * (2 * m * k2 * k2 - 2.0 / 3.0 * k2 ** 3) if m >= k else 4 * 2 * (2 * k * k2 * k2 - 2.0 / 3.0 * k2 ** 3)

    def svd(m, k):
        a, c = max(m, k), min(m, k)
        return 4 * (6 * a * c * c + 20 * c ** 3)

    total, site = 0.0, []
    right = [1] * (n + 1)                       # bond dims after the truncating sweep
    for i in range(n, 0, -1):
        right[i - 1] = min(chi, d * right[i], D * b[i - 1])
    left = [1] * (n + 1)                        # bond dims after the QR sweep
    for i in range(n):
        left[i + 1] = min(d * left[i], D * b[i + 1])
    for i in range(n):
        f = 8.0 * b[i] * b[i + 1] * d * d * D * D                              # apply
        if i < n - 1:
            f += qr(d * left[i], D * b[i + 1])                                     # QR
            f += 8.0 * left[i + 1] * D * b[i + 1] * d * D * b[i + 2]               # R-absorb
        if i > 0:
            f += svd(d * right[i + 1], left[i])                                    # SVD (reversed sweep)
            f += 8.0 * right[i] * left[i] * d * left[i - 1]                        # V-absorb
        site.append(f)
        total += f
    return total, max(site)


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) >= 7 and r[3 + k] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def fp64_tensor_peak(torch, cplx):
    """cuBLAS ZGEMM / DGEMM 4096^3 burst on this GPU: the FP64 tensor-pipe denominator."""
    n = 4096
    dt = torch.complex128 if cplx else torch.float64
    a = torch.randn(n, n, dtype=dt, device="cuda")
    b = torch.randn(n, n, dtype=dt, device="cuda")
    for _ in range(2):
        torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return (8.0 if cplx else 2.0) * n ** 3 / (best * 1e-3) / 1e12


# ----------------------------------------------------------------------------- CPU arm
def _ref_arm():
    import importlib.util
    spec = importlib.util.spec_from_file_location("tnb_ref_arm", os.path.join(ROOT, "baseline", "ref_arm.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def mpo_host(n):
    Wfull = tfi_w()
    ws = [Wfull[2] if i == 0 else (Wfull[:, 0] if i == n - 1 else Wfull) for i in range(n)]
    wl = [["right", "physout", "physin"] if i == 0 else (["left", "physout", "physin"] if i == n - 1 else
                                                         ["left", "right", "physout", "physin"]) for i in range(n)]
    return ws, wl


def pinned_norm(args):
    """|phi| of the compressed state as the UNMODIFIED reference computed it for this exact workload (rank 0's
    inputs, seed 2; tests/golden/make_golden_fullsize.py) -- None for any other size."""
    if (args.sites, args.d, args.chi) != (100, 2, 512):
        return None
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "fullsize_cfg3.npz"))
        return float(np.real(z["comp.norm_right"]))
    except Exception:
        return None


def run_reference_arm(args):
    """The reference's own CPU implementation of the path (baseline/_ref, else the oracle port) on all host
    threads: WHOLE sweeps on the inputs of rank 0 of the GPU arm -- nothing is extrapolated.  A sweep takes
    minutes of CPU, so at most two are timed (one if the first took more than ~100 s) whatever --steps says, and
    there is no warm-up sweep; `steps` / `warmup` in the line are what actually ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ra = _ref_arm()
    n, d, chi = args.sites, args.d, args.chi
    ws, wl = mpo_host(n)
    kind, times, nrm, bonds = ra.full_sweep_seconds(make_host_sites(n, d, chi, seed=2), ws, wl, chi,
                                                    max_steps=min(args.steps, 2), budget_s=200.0)
    sweep_s = float(np.mean(times))
    val = 1.0 / sweep_s
    pin = pinned_norm(args)
    sample = "%d whole sweep(s) (contract_mps_mpo + svd_compress(chi=%d), N=%d) through %s, no warm-up sweep" % (
        len(times), chi, n, "the unmodified reference package (baseline/_ref, NumPy-2 shim)" if kind == "reference"
        else "the NumPy/LAPACK oracle port (baseline/_ref missing)")
    line = {"metric": METRIC, "value": val, "unit": "sweeps/s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": sweep_s * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic", "impl": "reference",
            "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": "sweeps/s", "cores": ra.host_threads(), "kind": kind,
                             "sample": sample},
            "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "result": {"bonds_max": int(max(bonds)), "norm": nrm,
                       "norm_rel_err_vs_pinned": (abs(nrm - pin) / pin) if pin else None}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "cfg3: TFI MPO (D=3) applied to random MPS N=%d d=%d chi=%d complex128, then "
                        "svd_compress(chi=%d)" % (args.sites, args.d, args.chi, args.chi),
            "sites": args.sites, "d": args.d, "chi": args.chi, "mpo_bond": 3,
            "l2": "inputs larger than L2 (each sweep streams > 7 GB of site tensors); no explicit flush",
            "parallelism": "1 network per GPU (independent sweeps, no collective on the data path)"}


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sites", type=int, default=100)
    ap.add_argument("--d", type=int, default=2)
    ap.add_argument("--chi", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the batched (cfg 4) sub-record")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import tncontract_b200 as tn
    from tncontract_b200 import _lib
    od = tn.onedim
    lib = _lib.load()

    n, d, chi = args.sites, args.d, args.chi
    host_sites = make_host_sites(n, d, chi, seed=2 + rank)
    pinned = [torch.from_numpy(a).pin_memory() for a in host_sites]
    pinned_labels = [["phys", "left", "right"]] * n
    w_host, w_labels = mpo_host(n)
    w_pinned = [torch.from_numpy(np.ascontiguousarray(w)).pin_memory() for w in w_host]

    def upload():
        psi = od.MatrixProductState([tn.Tensor(p, l) for p, l in zip(pinned, pinned_labels)], "left", "right", "phys")
        H = od.MatrixProductOperator([tn.Tensor(w, l) for w, l in zip(w_pinned, w_labels)], "left", "right",
                                     "physout", "physin")
        return psi, H

    def sweep(psi, H):
        phi = od.contract_mps_mpo(psi, H)
        phi.svd_compress(chi=chi)
        return phi

    psi, H = upload()
    psi.left_canonise(qr_decomposition=True, normalise=True)   # setup, not timed (cfg 3 definition)
    # the end-to-end leg starts from the same canonised state, held in pinned host memory
    pinned = [torch.from_numpy(np.ascontiguousarray(np.asarray(t.data))).pin_memory() for t in psi]
    pinned_labels = [list(t.labels) for t in psi]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        phi = sweep(psi, H)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = tn.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        phi = sweep(psi, H)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = tn.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    bonds = phi.bonddims()
    phi_norm = float(phi.norm(canonical_form="right"))

    # ---- end to end: host inputs in, host result out ------------------------------------
    h2d = sum(p.numel() * 16 for p in pinned) + sum(w.numel() * 16 for w in w_pinned)
    d2h = 0
    # pinned result buffers (the compressed MPS has the shapes of the device-timed result)
    out_pinned = [torch.empty(tuple(t.shape), dtype=torch.complex128).pin_memory() for t in phi]
    def e2e_step():
        p2, H2 = upload()
        out = sweep(p2, H2)
        res = [t.data.get(out=buf) for t, buf in zip(out, out_pinned)]
        torch.cuda.synchronize()
        return sum(r.numel() * 16 for r in res)

    e2e_step()   # one untimed pass: the allocator has seen the upload / download pattern
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d2h = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---- one profiled pass: device time by kernel class --------------------------------------
    lib.tnb_profile_enable(1)
    # Jacobi sweeps per factorisation of this pass, by the number of columns k (the library reports the count of
    # every tnb_svd_project call; the wrapper only records it)
    from tncontract_b200 import devarray as _dv
    sweeps_seen = []
    _orig_svd_project = _dv.svd_project
    def _recording_svd_project(a_):
        out_ = _orig_svd_project(a_)
        sweeps_seen.append((int(min(a_.shape[-2:])) if len(a_.shape) >= 2 else 1, int(_dv.last_svd_sweeps)))
        return out_
    _dv.svd_project = _recording_svd_project
    try:
        sweep(psi, H)
    finally:
        _dv.svd_project = _orig_svd_project
    torch.cuda.synchronize()
    import ctypes
    prof = {}
    for c, name in enumerate(KCLASS):
        pms, pw = ctypes.c_double(), ctypes.c_double()
        pl, ps = ctypes.c_longlong(), ctypes.c_longlong()
        lib.tnb_profile_get(c, ctypes.byref(pms), ctypes.byref(pw), ctypes.byref(pl), ctypes.byref(ps))
        prof[name] = {"ms": pms.value, "work": pw.value, "launches": pl.value}
    lib.tnb_profile_enable(0)

    # ---- batched path (cfg 4) sub-record: networks/s at this N, sharded over the ranks ---------------------
    batched_rec = None
    if not args.no_batched:
        import bench_batch
        batched_rec = bench_batch.measure(rank, world, networks_per_gpu=148, batch_size=148)

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    pin = pinned_norm(args)
    zpeak = fp64_tensor_peak(torch, True)
    top = max(prof, key=lambda k: prof[k]["ms"])
    p = prof[top]
    per_launch_ms = p["ms"] / max(p["launches"], 1)
    if top in ("gemm", "jacobi_round", "qr_panel"):
        achieved = p["work"] / (p["ms"] * 1e-3) / 1e12
        roof = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": zpeak, "unit": "TFLOP/s",
                "frac": achieved / zpeak, "traffic": None,
                "peak_source": "cuBLAS ZGEMM 4096^3 burst measured in this run (FP64 tensor pipe; "
                               "MEASURED_PEAKS.json has no FP64 figure)"}
    else:
        achieved = p["work"] / (p["ms"] * 1e-3) / 1e9
        roof = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src}
    if top == "jacobi_round":
        # The class work is the CONVENTIONAL count of a round (full 32x32 complex Gram + apply, 4 real
        # products per complex one).  The kernel forms only the upper triangle of the Hermitian Gram matrix
        # and uses 3-multiplication complex products, so it issues (30 + 48) / 128 of those DMMAs in
        # projection mode; both rates are reported.
        executed = 78.0 / 128.0
        roof["executed_fraction_of_algorithmic"] = executed
        roof["executed_dmma_tflops"] = achieved * executed
        roof["executed_dmma_frac_of_peak"] = achieved * executed / zpeak
    try:  # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tr = json.load(f).get(top)
        if tr:
            roof["traffic"] = tr["dram_bytes_per_launch"]
            roof["traffic_source"] = tr["source"]
    except (OSError, ValueError, KeyError):
        pass
    if top == "jacobi_round" and sweeps_seen:
        big = [n_ for k_, n_ in sweeps_seen if k_ >= 1024]
        hist = {}
        for n_ in big:
            hist[str(n_)] = hist.get(str(n_), 0) + 1
        roof["jacobi_sweeps"] = {"factorisations": len(sweeps_seen), "k1024_histogram": hist,
                                 "k1024_mean": (sum(big) / len(big)) if big else None}
    roof["avg_launch_ms"] = per_launch_ms
    roof["launches_per_step"] = p["launches"]
    step_ms_prof = sum(v["ms"] for v in prof.values())
    roof["share_of_step"] = p["ms"] / step_ms_prof if step_ms_prof else None
    gemm = prof["gemm"]
    total_f, _ = nominal_flops(n, d, chi, 3)
    line = {
        "metric": METRIC, "value": world * args.steps / (ms_max * 1e-3), "unit": "sweeps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": world * args.steps / (e2e_ms_max * 1e-3), "unit": "sweeps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        "kernel_classes": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                               "rate": (v["work"] / (v["ms"] * 1e-3) / (1e12 if k in ("gemm", "jacobi_round", "qr_panel") else 1e9))
                               if v["ms"] > 0 else None} for k, v in prof.items()},
        "fp64_tensor": {"zgemm_peak_tflops": zpeak,
                        "gemm_class_tflops": gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else None,
                        "sweep_nominal_tflops": total_f / (ms_max / args.steps * 1e-3) / 1e12,
                        "sweep_nominal_frac_of_peak": total_f / (ms_max / args.steps * 1e-3) / 1e12 / zpeak},
        "result": {"bonds_max": int(max(bonds)), "norm": phi_norm,
                   "norm_rel_err_vs_pinned": (abs(phi_norm - pin) / pin) if pin else None,
                   "pinned_by": "tests/golden/fullsize_cfg3.npz (unmodified reference, same inputs)" if pin else None},
    }
    if batched_rec is not None:
        line["batched"] = {k: batched_rec[k] for k in ("metric", "value", "unit", "n_gpus", "networks", "networks_per_gpu",
                                                       "batch", "ms", "scaling", "launches_per_network_rank0",
                                                       "ragged_group_fallbacks_rank0", "max_bond_after")}
        if not args.no_cpu_baseline and world == 1:
            from tncontract_b200 import batch as _b
            ra = _ref_arm()
            ha = _b.random_mps(3, 0, 0, 64, 4, 128, on_host=True)
            hb = _b.random_mps(3, 0, 1, 64, 4, 128, on_host=True)
            kind, tt = ra.network_seconds(ha, hb, 64, reps=2)
            line["batched"]["cpu_baseline"] = {"value": 1.0 / tt[-1], "unit": "networks/s", "cores": ra.host_threads(),
                                               "kind": kind, "sample": "network 0 of seed 3, second of two runs"}
    if not args.no_cpu_baseline and world == 1:
        # bounded sample (about 10-20 s of CPU): the work of 8 bulk sites at the bulk shapes, through the
        # reference's own tensor functions; a sweep is 99 site steps of which the flop profile makes
        # `total / bulk` bulk-site equivalents.  The whole-sweep CPU measurement is `--impl reference`.
        ra = _ref_arm()
        total, bulk = nominal_flops(n, d, chi, 3)
        kind, times = ra.bulk_site_seconds(tfi_w(), d, chi, 3, 9)
        per_site = float(np.mean(times[1:]))
        line["cpu_baseline"] = {"value": 1.0 / (per_site * total / bulk), "unit": "sweeps/s", "cores": ra.host_threads(),
                                "kind": kind,
                                "sample": "8 timed bulk sites (+1 warm-up) of the sweep (apply+consolidate, QR 3072x1536, "
                                          "R-absorb, SVD 1024x1536, truncation, V/S-absorb) via %s; %.3f s per site x %.1f "
                                          "bulk-site equivalents (nominal flop profile of the N=%d chain); whole sweeps "
                                          "are timed by --impl reference" %
                                          ("the unmodified reference (baseline/_ref)" if kind == "reference"
                                           else "the NumPy/LAPACK oracle port", per_site, total / bulk, n)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
