#!/usr/bin/env python
"""Headline benchmark: Cooper-Frye cell x species x momentum evaluations per second (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path (default)
    python bench.py --impl reference ...                     # the reference's own CPU (OpenMP) path, bounded sample
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # cells sharded over N GPUs + one NCCL all-reduce

Workload (config.workload) = BASELINE.json config 2 on config 5's surface: continuous spectra, all 444 SMASH species,
shipped 51 pT x 1 phi x 21 y grid, df_mode 2 (RTA Chapman-Enskog) with bulk + shear + baryon diffusion on THE synthetic
3+1D surface of `--cells` cells (default 10 M; is3d2_b200/synthetic.py bench_surface: blocks S-3D(1.25 M, seed 2024 + k)).
The surface is the same for every N: rank r integrates its contiguous block of cells (strong scaling).  One "step" = one
full Cooper-Frye pass over the surface: per-cell set-up kernel + spectra kernel + partial reduction on every GPU, then (N > 1)
ONE ncclAllReduce of the spectra issued by the product itself (is3d_comm_attach / is3d_spectra_device).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

# The C++ host layer keeps the reference's printf chatter ("Number of chosen particles = ...") on stdout.  The contract is
# ONE JSON line on stdout, so fd 1 is pointed at stderr for the whole run and the line is written to the saved descriptor.
_JSON_FD = None


def _claim_stdout() -> None:
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)

from is3d2_b200 import synthetic, workdir  # noqa: E402

# algorithmic FLOPs per integrand evaluation (SURVEY.md 8d / BASELINE.md 4; DESIGN.md restates the derivation)
F_ALG = {1: 165.0, 2: 194.0, 3: 265.0, 4: 265.0, 5: 265.0}
NS_SMASH, NPT, NPHI, NY = 444, 51, 1, 21


def bench_params(df_mode: int) -> dict:
    # PTB (df_mode 4) has no muB != 0 coefficient tables (reference DeltafData.cpp:480-484)
    b = 0 if df_mode == 4 else 1
    return dict(operation=1, mode=1, hrg_eos=2, dimension=3, df_mode=df_mode, include_baryon=b,
                include_bulk_deltaf=1, include_shear_deltaf=1, include_baryondiff_deltaf=b, regulate_deltaf=0, outflow=0)


def workload_name(df_mode: int, cells: int, n_gpus: int) -> str:
    terms = "bulk+shear" if df_mode == 4 else "bulk+shear+baryon diffusion"
    return (f"continuous spectra, all SMASH species (444), df_mode={df_mode} with {terms}, 51pT x 1phi x 21y, "
            f"one synthetic 3+1D surface of {cells} cells (S-3D blocks, seeds 2024+k) sharded over {n_gpus} GPU")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_port_evals_per_s(args, cells: int) -> tuple[float, float]:
    """Fallback CPU arm when oracle/_ref (the compiled reference) is not on the box: this repo's scalar restatement of
    the reference loop (oracle/cf_oracle.cpp, one thread) on `cells` cells of the same workload.  Returns (evals/s, s)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_api
    surf = synthetic.bench_surface(0, cells, baryon=True)
    root = tempfile.mkdtemp(prefix="is3d_port_")
    try:
        params = bench_params(args.df_mode)
        workdir.make_workdir(root, params, chosen="smash")
        prob = oracle_api.OracleProblem(root, params, surf)
        t0 = time.perf_counter()
        rc, _, _ = prob.spectra()
        sec = time.perf_counter() - t0
    finally:
        shutil.rmtree(root, ignore_errors=True)
    if rc != 0:
        raise RuntimeError(f"oracle port failed with status {rc}")
    return float(cells) * NS_SMASH * NPT * NPHI * NY / sec, sec


def reference_openmp_seconds(args, cells: int, runs: int) -> tuple[list[float], str, int]:
    """The unmodified reference's OpenMP build (oracle/_ref) on the first `cells` cells of the benchmark surface, all host
    threads; returns (seconds inside calculate_spectra per run, binary used, threads).  The -march=native build is tried
    first ("native" = the build container's CPU) and dropped for the plain -O3 one if this host's CPU
    rejects it."""
    cores = os.cpu_count() or 1
    surf = synthetic.bench_surface(0, cells, baryon=True)
    params = bench_params(args.df_mode)
    root = tempfile.mkdtemp(prefix="is3d_ref_")
    times, used = [], None
    try:
        workdir.make_workdir(root, params, chosen="smash")
        synthetic.write_mode1(os.path.join(root, "input", "surface.dat"), surf, baryon=True)
        env = dict(os.environ, OMP_NUM_THREADS=str(cores))
        for name in ("is3d_ref_omp_native", "is3d_ref_omp"):
            exe = os.path.join(REPO, "oracle", "_ref", name)
            if not os.access(exe, os.X_OK):
                continue
            try:
                times = []
                for _ in range(runs):
                    with open(os.path.join(root, "ref_stdout.log"), "w") as log:
                        subprocess.run([exe], cwd=root, stdout=log, stderr=subprocess.STDOUT, env=env, check=True)
                    times.append(float(open(os.path.join(root, "ref_dump", "timing.txt")).read().split()[0]))
                used = name
                break
            except (subprocess.CalledProcessError, OSError, ValueError) as ex:
                print(f"bench.py: {name} failed on this host ({ex!r}); trying the next reference build", file=sys.stderr)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    if used is None:
        raise RuntimeError("no runnable oracle/_ref OpenMP binary")
    return times, used, cores


def have_reference_binary() -> bool:
    return any(os.access(os.path.join(REPO, "oracle", "_ref", n), os.X_OK) for n in ("is3d_ref_omp_native", "is3d_ref_omp"))


def run_reference(args) -> None:
    """The reference's own CPU implementation (oracle/_ref, unmodified sources, OpenMP build) on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cells = args.ref_cells
    evals = float(cells) * NS_SMASH * NPT * NPHI * NY
    common = {"impl": "reference", "metric": "Cooper-Frye cell*species*momentum evals/s", "unit": "evals/s", "n_gpus": args.gpus,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "gpu_launches": 0}
    if not have_reference_binary():
        # the compiled reference did not travel: time the oracle port instead (kind = "port", one core)
        pc = max(50, cells // 10)
        vals = [oracle_port_evals_per_s(args, pc) for _ in range(max(1, min(args.steps, 2)))]
        value, sec = float(np.mean([v[0] for v in vals])), float(np.mean([v[1] for v in vals]))
        sample = f"first {pc} cells of the benchmark surface, oracle/cf_oracle.cpp (scalar restatement of the reference loop)"
        emit(dict(common, value=value, steps=len(vals), warmup=0, ms_per_step=sec * 1e3,
                  config={"workload": workload_name(args.df_mode, args.cells, args.gpus), "sample": sample},
                  cpu_baseline={"value": value, "unit": "evals/s", "cores": 1, "kind": "port", "sample": sample},
                  e2e={"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}))
        return
    times, used, cores = reference_openmp_seconds(args, cells, args.warmup + args.steps)
    sec = float(np.mean(times[args.warmup:]))
    value = evals / sec
    sample = (f"first {cells} cells of the benchmark surface (evals/s is linear in cells), oracle/_ref/{used} with {cores} OpenMP threads, "
              "timed inside calculate_spectra")
    emit(dict(common, value=value, steps=args.steps, warmup=args.warmup, ms_per_step=sec * 1e3,
              config={"workload": workload_name(args.df_mode, args.cells, args.gpus), "sample": sample,
                      "note": "3+1D OpenMP loop of the reference is racy (shared etaValues[0]); timing only"},
              cpu_baseline={"value": value, "unit": "evals/s", "cores": cores, "kind": "reference", "sample": sample},
              e2e={"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}))


def cpu_baseline(args) -> dict:
    """Bounded reference run beside the GPU number (rank 0, N = 1 only)."""
    if not have_reference_binary():
        pc = max(50, args.ref_cells // 10)
        value, sec = oracle_port_evals_per_s(args, pc)
        return {"value": value, "unit": "evals/s", "cores": 1, "kind": "port",
                "sample": f"first {pc} cells of the benchmark surface, oracle/cf_oracle.cpp on one core, {sec:.2f} s (oracle/_ref not on this box)"}
    cells = args.ref_cells
    times, used, cores = reference_openmp_seconds(args, cells, 1)
    evals = float(cells) * NS_SMASH * NPT * NPHI * NY
    return {"value": evals / times[0], "unit": "evals/s", "cores": cores, "kind": "reference",
            "sample": f"first {cells} cells of the benchmark surface, oracle/_ref/{used} with {cores} OpenMP threads, {times[0]:.2f} s inside calculate_spectra"}


SAMPLER_PARAMS = dict(operation=2, mode=1, hrg_eos=2, dimension=3, df_mode=3, include_baryon=0, include_bulk_deltaf=1,
                      include_shear_deltaf=1, include_baryondiff_deltaf=0, regulate_deltaf=0, outflow=0, oversample=1, fast=1,
                      test_sampler=0, sampler_seed=1, min_num_hadrons=1.0e12, max_num_samples=1000)


def sampler_bench(args, rank: int, world: int, local: int) -> dict:
    """BASELINE.json config 3: particle sampler, full SMASH HRG, df_mode 3 (PTM), 1000 oversampled events.  The surface is
    the concatenation of `sampler_cells`-cell blocks S-3D(seed 3024 + k, stress 0.3), one block per GPU (weak: the sampler
    shards cells and concatenates events, no collective; Philox streams keyed by the global cell index).  Timed end to end
    through the C ABI over `sampler_calls` calls after one warm-up call: surface resident, particle records copied back to
    (library-owned, pinned) host memory inside the timed region."""
    import torch
    import torch.distributed as dist

    from is3d2_b200 import HostSession, shard

    cells, nev, calls = args.sampler_cells, args.sampler_events, max(1, args.sampler_calls)
    surf = synthetic.s3d(cells, seed=3024 + rank, stress=0.3)
    root = tempfile.mkdtemp(prefix=f"is3d_smp_r{rank}_")
    res = {}
    try:
        workdir.make_workdir(root, SAMPLER_PARAMS, chosen="smash")
        with HostSession(root) as h:
            h.set_surface(surf)
            shard.set_global_thermo_averages(h)
            h.prepare()
            h.abi_set_surface(surf, global_offset=rank * cells)
            ntot, _ = h.abi_total_yield()
            def device_sample(nev_, copy=False):
                ptr, total, counts, st = h.abi_sample_device(nev_)          # list stays in HBM (device-resident consumer)
                return np.empty(int(total), dtype=np.uint8), counts, st, (lambda: None)

            for mode in ("full", "compact", "device"):
                sample = {"full": h.abi_sample, "compact": h.abi_sample_compact, "device": device_sample}[mode]
                rec_bytes = {"full": 104, "compact": 64, "device": 0}[mode]
                out = sample(nev, copy=False)                          # warm-up: same size, so the pinned list buffer is reused
                out[3]()
                secs, kms = [], []
                n_parts = proposals = 0
                for _ in range(calls):
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    parts, counts, st, release = sample(nev, copy=False)       # view of the library-owned pinned list
                    torch.cuda.synchronize()
                    secs.append(time.perf_counter() - t0)
                    kms.append(st.kernel_ms)
                    n_parts, proposals = len(parts), st.sampler_proposals
                    del parts
                    release()
                sec, dev_ms = float(np.median(secs)), float(np.median(kms))
                acc = torch.tensor([float(n_parts), float(proposals), sec, dev_ms, min(secs), max(secs)], dtype=torch.float64, device="cuda")
                if world > 1:
                    sums = acc.clone()
                    dist.all_reduce(sums)
                    dist.all_reduce(acc, op=dist.ReduceOp.MAX)
                    hadrons, props = float(sums[0]), float(sums[1])
                else:
                    hadrons, props = float(acc[0]), float(acc[1])
                sec, dev_ms, smin, smax = float(acc[2]), float(acc[3]), float(acc[4]), float(acc[5])
                res[mode] = {"hadrons_per_s": hadrons / sec, "seconds_median": sec, "seconds_min": smin, "seconds_max": smax,
                             "device_ms_median": dev_ms, "hadrons_per_s_device": hadrons / (dev_ms * 1e-3) if dev_ms > 0 else None,
                             "record_bytes": rec_bytes, "d2h_bytes": int(hadrons) * rec_bytes,
                             "d2h_gbs_of_wall": hadrons * rec_bytes / sec / 1e9, "hadrons": int(hadrons),
                             "proposals_per_s": props / sec, "acceptance": hadrons / max(props, 1.0)}
    finally:
        shutil.rmtree(root, ignore_errors=True)
    best = res["compact"]
    # what this box can move device -> pinned host memory when all ranks copy at once: the ceiling of any list delivery
    nbytes = max(int(best["d2h_bytes"] // world), 1 << 20)
    dev_buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host_buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host_buf.copy_(dev_buf)
    secs = []
    for _ in range(calls):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_buf.copy_(dev_buf, non_blocking=True)
        torch.cuda.synchronize()
        secs.append(time.perf_counter() - t0)
    tc = torch.tensor([float(np.median(secs))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    ceiling = nbytes * world / float(tc.item()) / 1e9
    del dev_buf, host_buf
    return {"metric": "sampled hadrons/s", "value": best["hadrons_per_s"], "unit": "hadrons/s", "calls": calls,
            "records": "value = 64-byte wire records delivered to host memory (is3d_sample_compact); beside it the 104-byte Sampled_Particle "
                       "records (is3d_sample) and the list left in HBM (is3d_sample_device: no PCIe transfer)",
            "compact": res["compact"], "full": res["full"], "device": res["device"],
            "d2h_ceiling_gbs": ceiling,
            "d2h_ceiling_note": f"plain cudaMemcpy of {nbytes} bytes per rank, HBM -> pinned host, all {world} ranks at once (aggregate GB/s)",
            "mean_yield_per_event_rank0": ntot,
            "workload": f"sampler, df_mode=3 PTM, fast=1, all SMASH species, {nev} events, S-3D(stress 0.3) {cells} cells per GPU x {world} GPU, "
                        "particle lists returned to host memory"}


def sampler_cpu_baseline(args) -> dict:
    """The reference's (serial; it has no parallel sampler) sample_dN_pTdpTdphidy on a bounded sample."""
    exe = os.path.join(REPO, "oracle", "_ref", "is3d_ref")
    if not os.access(exe, os.X_OK):
        return {"value": None, "unit": "hadrons/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    cells, nev = args.ref_sampler_cells, args.sampler_events
    surf = synthetic.s3d(cells, seed=3024, stress=0.3)
    root = tempfile.mkdtemp(prefix="is3d_cpu_smp_")
    try:
        workdir.make_workdir(root, dict(SAMPLER_PARAMS, test_sampler=1), chosen="smash")
        synthetic.write_mode1(os.path.join(root, "input", "surface.dat"), surf, baryon=False)
        with open(os.path.join(root, "ref_stdout.log"), "w") as log:
            subprocess.run([exe], cwd=root, stdout=log, stderr=subprocess.STDOUT, check=True)
        sec = float(open(os.path.join(root, "ref_dump", "timing.txt")).read().split()[0])
        ntot = float(np.fromfile(os.path.join(root, "ref_dump", "total_yield.bin"), dtype=np.float64)[0])
        n_ev = int(open(os.path.join(root, "ref_dump", "nevents.txt")).read().split()[0])
    finally:
        shutil.rmtree(root, ignore_errors=True)
    hadrons = ntot * n_ev
    return {"value": hadrons / sec, "unit": "hadrons/s", "cores": 1, "kind": "reference",
            "sample": f"{cells} cells x {n_ev} events of the same workload (mean {hadrons:.3g} hadrons = events x calculate_total_yield), "
                      f"serial reference, {sec:.2f} s inside calculate_spectra (histogram mode, no particle files)"}


def oracle_spot_check(args, h, surf: dict, avg) -> dict:
    """Checker leg (rank 0, before any communicator is attached): the first `check_cells` cells of the benchmarked surface
    through the SAME context and kernels against the CPU oracle (oracle/cf_oracle.cpp, pinned to the unmodified reference)."""
    n = args.check_cells
    try:
        sys.path.insert(0, os.path.join(REPO, "tests"))
        import harness
        import oracle_api
        sl = {k: np.ascontiguousarray(v[:n]) for k, v in surf.items()}
        h.abi_set_surface(sl, global_offset=0)
        got, _ = h.abi_spectra()
        root = tempfile.mkdtemp(prefix="is3d_check_")
        try:
            params = bench_params(args.df_mode)
            workdir.make_workdir(root, params, chosen="smash")
            # same surface-averaged thermodynamics as the benchmarked context (they fix the PTB tables of df_mode 4)
            rc, want, _ = oracle_api.OracleProblem(root, params, sl, famod_chain=0,
                                                   after_surface=lambda s: s.set_thermo_averages(avg)).spectra()
        finally:
            shutil.rmtree(root, ignore_errors=True)
        if rc != 0:
            return {"cells": n, "ok": False, "error": f"oracle status {rc}"}
        tol = harness.RTOL
        try:
            worst = harness.assert_spectra_close(got, want, rtol=tol, what="bench spot check")
            return {"cells": n, "max_rel_err": worst, "tolerance": tol, "ok": True, "against": "oracle/cf_oracle.cpp on the surface's first cells"}
        except AssertionError as e:
            return {"cells": n, "tolerance": tol, "ok": False, "error": str(e)[:300]}
    except Exception as e:  # noqa: BLE001 -- the checker must not take the measurement down
        return {"cells": n, "ok": None, "error": f"checker unavailable: {e!r}"[:300]}


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from is3d2_b200 import HostSession

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this repository has no CPU compute path")
    torch.cuda.set_device(local)
    os.environ["IS3D_DEVICE"] = str(local)
    os.environ.pop("IS3D_DEVICES", None)                    # one process per GPU here; the group API is the other way in
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    params = bench_params(args.df_mode)
    root = tempfile.mkdtemp(prefix=f"is3d_bench_r{rank}_")
    workdir.make_workdir(root, params, chosen="smash")
    h = HostSession(root)
    try:
        line = _measure_ours(args, h, world, rank, local)
        if rank == 0:
            emit(line)
    finally:
        # tear-down order matters: every torch tensor that touched the context's stream or NCCL must be released before
        # the process group and the stream go away (a tensor freed afterwards aborts the rank with "context is destroyed")
        import gc
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        h.close()                                           # destroys the context and its communicator
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        shutil.rmtree(root, ignore_errors=True)


def attach_product_communicator(h, world: int, rank: int) -> None:
    """The all-reduce of the path lives in the product (is3d_comm_attach -> ncclAllReduce inside is3d_spectra[_device]);
    torch.distributed only ships the 128-byte NCCL id from rank 0 to the other ranks."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    buf = (C.c_char * 128)()
    if rank == 0:
        st = h.lib.is3d_comm_unique_id(buf)
        if st != 0:
            raise RuntimeError("is3d_comm_unique_id: " + h.lib.is3d_comm_last_error().decode())
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, src=0)
    ident = bytes(t.cpu().tolist())
    h._check(h.lib.is3d_comm_attach(h.ctx, ident, world, rank), "is3d_comm_attach")


def _library_margin(h) -> float:
    """is3d_params.negligible_margin the benchmarked context was created with (the host layer reads IS3D_NEGLIGIBLE_MARGIN)."""
    import ctypes as C

    from is3d2_b200 import capi
    p = capi.Params()
    h.lib.is3d_default_params(C.byref(p))
    return float(os.environ.get("IS3D_NEGLIGIBLE_MARGIN", p.negligible_margin))


def _measure_ours(args, h, world: int, rank: int, local: int):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from is3d2_b200 import sassinfo, shard

    # ONE surface, whatever the number of GPUs (strong scaling): rank r integrates the contiguous cell block [b, e)
    G = args.cells
    b, e = shard.cell_range(G, rank, world)
    cells = e - b
    surf = synthetic.bench_surface(b, e, baryon=True)
    h.set_surface(surf)                                     # host layer: this block's thermodynamic sums ...
    avg = shard.set_global_thermo_averages(h)               # ... -> averages of the WHOLE surface (set-up; torch.distributed)
    h.prepare()
    shape = h.spectra_shape()
    total = int(np.prod(shape))
    evals_global = float(G) * total

    check = oracle_spot_check(args, h, surf, avg) if (rank == 0 and args.check_cells > 0) else None
    if world > 1:
        attach_product_communicator(h, world, rank)

    ext = torch.cuda.ExternalStream(h.lib.is3d_stream(h.ctx), device=torch.device("cuda", local))
    # pinned host copies (e2e path) and resident device copies (value path)
    host_cols = {k: torch.from_numpy(v).pin_memory() for k, v in surf.items()}
    del surf
    dev_cols = {k: t.cuda(non_blocking=True) for k, t in host_cols.items()}
    torch.cuda.synchronize()
    out_dev = torch.zeros(total, dtype=torch.float64, device="cuda")
    h.abi_set_surface_device({k: t.data_ptr() for k, t in dev_cols.items()}, cells, global_offset=b)

    def step():
        # set-up kernel + spectra kernel + partial reduction + (N > 1) the product's own ncclAllReduce of the spectra
        return h.abi_spectra_device(out_dev.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(ext):
        for _ in range(args.warmup):
            step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kernel_ms, launches, skipped, executed, executed_pair, dropped, reruns = 0.0, 0, 0, 0, 0, 0, 0
        ev0.record(ext)
        for _ in range(args.steps):
            st = step()
            kernel_ms += st.kernel_ms
            launches += st.kernel_launches
            skipped, executed, executed_pair = st.cells_skipped, st.evals_executed, st.pair_evals_executed
            dropped, reruns = st.evals_dropped, reruns + st.prune_reruns
        ev1.record(ext)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = ev0.elapsed_time(ev1)
        value_out = out_dev.cpu().numpy().copy()
        collectives = int(h.lib.is3d_comm_collectives(h.ctx))

        # ---- end to end through the C ABI with HOST buffers: H2D of the 25 columns + compute (+ all-reduce) + D2H ----
        host_np = {k: t_.numpy() for k, t_ in host_cols.items()}

        def e2e_step():
            h.abi_set_surface(host_np, global_offset=b)
            return h.abi_spectra()[0]

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()                                                       # untimed: first-use allocations
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            spec = e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps

    # max over ranks of the two times, sums of the per-rank counters
    t = torch.tensor([ms, e2e_s, kernel_ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([float(skipped), float(executed), float(cells), float(executed_pair), float(dropped), float(reruns)],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c)
    ms_total, e2e_s, kernel_ms = (float(v) for v in t.tolist())
    skipped_all, executed_all, cells_all, pair_all, dropped_all, reruns_all = (float(v) for v in c.tolist())
    assert int(cells_all) == G

    fp64_peak = h.abi_fp64_peak()
    sampler = None if args.no_sampler else sampler_bench(args, rank, world, local)
    if rank != 0:
        return None

    ms_step = ms_total / args.steps
    kern_s = kernel_ms / args.steps * 1e-3                               # slowest rank's dominant-kernel time per step
    value = evals_global / (ms_step * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    # ---- roofline of the dominant kernel, from what THIS run executed ----
    # executed FP64-pipe work = class-evaluations the kernel ran (is3d_stats.evals_executed: padding slots and idle thread
    # columns included) x FP64-pipe instructions per class-evaluation read from the SASS of the loaded library x 2 flops
    # two launches per pass: single classes, and charge-conjugate pairs (two class-evaluations from one exponential)
    label = {1: "df_spectra_kernel<1,1,0,0,4,0>", 2: "df_spectra_kernel<2,1,0,0,4,0>"}.get(args.df_mode)
    label_pair = {1: "df_spectra_kernel<1,1,0,0,4,1>", 2: "df_spectra_kernel<2,1,0,0,4,1>"}.get(args.df_mode)
    mix = mix_pair = None
    try:
        kern = sassinfo.library_info()["kernels"]
        mix, mix_pair = (kern.get(label), kern.get(label_pair)) if label else (None, None)
    except Exception as ex:  # noqa: BLE001
        print(f"bench.py: SASS scan failed: {ex!r}", file=sys.stderr)
    achieved = frac = per_eval = per_eval_pair = None
    if mix and executed_all > 0 and (mix_pair or pair_all == 0):
        per_eval = mix["fp64"] / mix["evals_per_trip"]
        per_eval_pair = mix_pair["fp64"] / mix_pair["evals_per_trip"] if mix_pair else 0.0
        fp64_instr = (executed_all - pair_all) * per_eval + pair_all * per_eval_pair
        # all ranks run side by side: per-GPU rate = (executed on the slowest rank ~ total / world) / its kernel time
        achieved = fp64_instr / world * 2.0 / kern_s / 1e12
        frac = achieved / fp64_peak
    # DRAM bytes of one launch: only from an `ncu --set full` capture of the SAME inner loop (listing hash) and the same launch
    traffic = traffic_src = None
    try:
        cap = json.load(open(os.path.join(REPO, "profiles", "ncu_k1_headline.json")))
        have = {label: mix["listing_sha256"] if mix else None, label_pair: mix_pair["listing_sha256"] if mix_pair else None}
        if mix and cap.get("listing_sha256") == have and cap.get("cells_per_launch"):
            per_cell = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) / cap["cells_per_launch"]
            traffic = per_cell * min(cells, 4 << 20)
            traffic_src = (f"ncu dram__bytes_read.sum + dram__bytes_write.sum of the pair + single launches of one pass, per cell ({per_cell:.1f} B) "
                           "x cells of one pass, profiles/ncu_k1_headline.json (same inner-loop SASS as the loaded library)")
    except (OSError, KeyError, ValueError):
        pass
    bytes_alg = cells * 25 * 8.0 + total * 8.0
    line = {
        "metric": "Cooper-Frye cell*species*momentum evals/s", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.df_mode, G, world), "global_cells": G, "cells_per_gpu": cells if world == 1 else f"{G // world}..{-(-G // world)}",
                   "evals_per_step": evals_global, "l2_policy": "inputs larger than L2 (cell packs: 256 B x cells per pass)",
                   "parallelism": (f"one surface, cells sharded x{world} in contiguous blocks; the product all-reduces {total} doubles "
                                   f"(ncclAllReduce inside is3d_spectra_device, {collectives} issued by rank 0)") if world > 1 else "single GPU",
                   "species_classes": "evaluations are counted per species (444); species with identical (mass, sign, baryon "
                                      "number) share one integrand, computed once and scaled by each species' degeneracy "
                                      "(193 classes for the SMASH list) -- every species' bins are delivered"},
        "e2e": {"value": evals_global / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(G * 25 * 8),
                "d2h_bytes_per_step": int(total * 8 * world), "seconds_per_step": e2e_s,
                "equals_value_path": bool(np.array_equal(spec.reshape(-1), value_out))},
        "gpu_launches": int(launches),
        "cells_skipped_frac": skipped_all / G,
        "negligible_items": {
            "margin": _library_margin(h), "class_evals_dropped_per_step": dropped_all,
            "dropped_frac_of_thread_slots": dropped_all / (dropped_all + executed_all) if executed_all else None,
            "reruns_in_timed_steps": int(reruns_all),
            "what": "(cell, y, phi) items whose every exponent (u.p - b mu_B)/T over a block of momentum columns is >= 680 (Bose/Fermi "
                    "factor < 1e-295, the reference's exp overflows at 709.8) or exceeds the block row's smallest possible exponent by "
                    "more than `margin` are dropped before the momentum loop; the library sums a rigorous bound of every dropped term "
                    "and tests it against each finished bin (bound <= 1e-13 |bin|), repeating the call without the margin if any bin "
                    "fails (a rerun is inside the timed region). --negligible-margin 0 measures without the margin"},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": frac,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": (label + " + " + label_pair) if label else "feqmod_spectra_kernel", "kernel_ms_per_step": kernel_ms / args.steps,
                     "what": "executed FP64-pipe instructions x 2 flops / kernel time / live DFMA peak, per GPU",
                     "fp64_instr_per_class_eval": per_eval, "fp64_instr_per_class_eval_in_pair_slots": per_eval_pair,
                     "inner_loop_instructions": mix["instructions"] if mix else None,
                     "inner_loop_instructions_pair_kernel": mix_pair["instructions"] if mix_pair else None,
                     "inner_loop_sass_sha256": mix["listing_sha256"] if mix else None,
                     "class_evals_executed_per_step": executed_all, "class_evals_in_pair_slots_per_step": pair_all,
                     "species_evals_delivered_per_step": evals_global * (1.0 - skipped_all / G),
                     "peak_source": "DFMA micro-benchmark run live by is3d_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 entry)",
                     "reference_operation_count": {"flops_per_eval": F_ALG[args.df_mode], "tflops_equivalent": evals_global / world * F_ALG[args.df_mode] / kern_s / 1e12,
                                                   "note": "SURVEY.md 8d's provisional count of the REFERENCE's loop per species-evaluation; not a roofline fraction "
                                                           "(species classes, folded polynomials and table-driven exp execute far fewer instructions)"},
                     "hbm_gbs_algorithmic": bytes_alg / kern_s / 1e9, "hbm_peak_gbs_measured": peaks.get("hbm_gbs")},
        "check": check,
        "clocks": clocks,
    }
    if sampler is not None:
        line["sampler"] = sampler
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
        if sampler is not None:
            line["sampler"]["cpu_baseline"] = sampler_cpu_baseline(args)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--negligible-margin", type=float, default=None,
                    help="is3d_params.negligible_margin of the benchmarked context (default: the library's, 80; 0 = off)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=10_000_000,
                    help="cells of THE benchmark surface (BASELINE.json config 5: 10 M), shared by all GPUs (strong scaling)")
    ap.add_argument("--check-cells", type=int, default=64, help="cells of the oracle spot check (0 = off)")
    ap.add_argument("--sampler-calls", type=int, default=5)
    ap.add_argument("--df-mode", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="2 = the headline workload; 1, 3, 4, 5 time the other df corrections on the same surface")
    ap.add_argument("--ref-cells", type=int, default=5000, help="cells of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sampler", action="store_true", help="skip the sampler (hadrons/s) section")
    ap.add_argument("--sampler-cells", type=int, default=100_000)
    ap.add_argument("--sampler-events", type=int, default=1000)
    ap.add_argument("--ref-sampler-cells", type=int, default=3000)
    args = ap.parse_args()
    _claim_stdout()
    if args.negligible_margin is not None:
        os.environ["IS3D_NEGLIGIBLE_MARGIN"] = repr(float(args.negligible_margin))      # read by the host layer at prepare()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
