#!/usr/bin/env python
"""Headline benchmark: Cooper-Frye cell x species x momentum evaluations per second (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path (default)
    python bench.py --impl reference ...                     # the reference's own CPU (OpenMP) path, bounded sample
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # cells sharded over N GPUs + one NCCL all-reduce

Workload (config.workload): continuous spectra, all 444 SMASH species, shipped 51 pT x 1 phi x 21 y grid,
df_mode 2 (RTA Chapman-Enskog) with bulk + shear + baryon diffusion on the seeded synthetic 3+1D surface S-3D
(SURVEY.md 8d), `--cells-per-gpu` cells per GPU (weak scaling; 1.25 M x 8 GPUs = the 10 M-cell surface of
BASELINE.json config 5).  One "step" = one full Cooper-Frye pass over the rank's cells (per-cell set-up kernel +
spectra kernel + partial reduction) followed, for N > 1, by the all-reduce of the spectra array.
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
    return (f"continuous spectra, all SMASH species (444), df_mode={df_mode} with {terms}, "
            f"51pT x 1phi x 21y, synthetic 3+1D surface S-3D, {cells} cells per GPU x {n_gpus} GPU")


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
    surf = synthetic.s3d(cells, seed=2024, baryon=True)
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


def run_reference(args) -> None:
    """The reference's own CPU implementation (oracle/_ref, unmodified sources, OpenMP build) on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    exe = os.path.join(REPO, "oracle", "_ref", "is3d_ref_omp")
    cores = os.cpu_count() or 1
    cells = args.ref_cells
    if not os.access(exe, os.X_OK):
        # the compiled reference did not travel: time the oracle port instead (kind = "port", one core)
        pc = max(50, cells // 10)
        vals = [oracle_port_evals_per_s(args, pc) for _ in range(max(1, min(args.steps, 2)))]
        value, sec = float(np.mean([v[0] for v in vals])), float(np.mean([v[1] for v in vals]))
        sample = f"{pc}-cell S-3D sample (seed 2024) of the same workload, oracle/cf_oracle.cpp (scalar restatement of the reference loop)"
        line = {"impl": "reference", "metric": "Cooper-Frye cell*species*momentum evals/s", "value": value, "unit": "evals/s",
                "n_gpus": args.gpus, "steps": len(vals), "warmup": 0, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.df_mode, args.cells_per_gpu, args.gpus), "sample": sample},
                "cpu_baseline": {"value": value, "unit": "evals/s", "cores": 1, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return
    surf = synthetic.s3d(cells, seed=2024, baryon=True)
    params = bench_params(args.df_mode)
    root = tempfile.mkdtemp(prefix="is3d_ref_")
    try:
        workdir.make_workdir(root, params, chosen="smash")
        synthetic.write_mode1(os.path.join(root, "input", "surface.dat"), surf, baryon=True)
        env = dict(os.environ, OMP_NUM_THREADS=str(cores))
        times = []
        for i in range(args.warmup + args.steps):
            with open(os.path.join(root, "ref_stdout.log"), "w") as log:
                subprocess.run([exe], cwd=root, stdout=log, stderr=subprocess.STDOUT, env=env, check=True)
            t = float(open(os.path.join(root, "ref_dump", "timing.txt")).read().split()[0])
            if i >= args.warmup:
                times.append(t)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    evals = float(cells) * NS_SMASH * NPT * NPHI * NY
    sec = float(np.mean(times))
    value = evals / sec
    sample = f"{cells}-cell prefix-sized S-3D sample (seed 2024) of the same workload, timed inside calculate_spectra"
    line = {"impl": "reference", "metric": "Cooper-Frye cell*species*momentum evals/s", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.df_mode, args.cells_per_gpu, args.gpus), "sample": sample,
                       "note": "3+1D OpenMP loop of the reference is racy (shared etaValues[0]); timing only"},
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def cpu_baseline(args) -> dict:
    """Bounded reference run beside the GPU number (rank 0, N = 1 only)."""
    exe = os.path.join(REPO, "oracle", "_ref", "is3d_ref_omp")
    cores = os.cpu_count() or 1
    if not os.access(exe, os.X_OK):
        pc = max(50, args.ref_cells // 10)
        value, sec = oracle_port_evals_per_s(args, pc)
        return {"value": value, "unit": "evals/s", "cores": 1, "kind": "port",
                "sample": f"{pc} cells of the same S-3D workload, oracle/cf_oracle.cpp on one core, {sec:.2f} s (oracle/_ref not on this box)"}
    cells = args.ref_cells
    surf = synthetic.s3d(cells, seed=2024, baryon=True)
    root = tempfile.mkdtemp(prefix="is3d_cpu_")
    try:
        workdir.make_workdir(root, bench_params(args.df_mode), chosen="smash")
        synthetic.write_mode1(os.path.join(root, "input", "surface.dat"), surf, baryon=True)
        env = dict(os.environ, OMP_NUM_THREADS=str(cores))
        with open(os.path.join(root, "ref_stdout.log"), "w") as log:
            subprocess.run([exe], cwd=root, stdout=log, stderr=subprocess.STDOUT, env=env, check=True)
        sec = float(open(os.path.join(root, "ref_dump", "timing.txt")).read().split()[0])
    finally:
        shutil.rmtree(root, ignore_errors=True)
    evals = float(cells) * NS_SMASH * NPT * NPHI * NY
    return {"value": evals / sec, "unit": "evals/s", "cores": cores, "kind": "reference",
            "sample": f"{cells} cells of the same S-3D workload, reference OpenMP build with {cores} threads, {sec:.2f} s inside calculate_spectra"}


SAMPLER_PARAMS = dict(operation=2, mode=1, hrg_eos=2, dimension=3, df_mode=3, include_baryon=0, include_bulk_deltaf=1,
                      include_shear_deltaf=1, include_baryondiff_deltaf=0, regulate_deltaf=0, outflow=0, oversample=1, fast=1,
                      test_sampler=0, sampler_seed=1, min_num_hadrons=1.0e12, max_num_samples=1000)


def sampler_bench(args, rank: int, world: int, local: int) -> dict:
    """BASELINE.json config 3: particle sampler, full SMASH HRG, df_mode 3 (PTM), 1000 oversampled events, on this
    rank's block of a synthetic 3+1D surface.  Timed end to end through is3d_sample (host surface already on the
    device; particle records copied back to host memory inside the timed region).  No collective: ranks sample
    disjoint cell blocks with Philox streams keyed by the global cell index."""
    import torch
    import torch.distributed as dist

    from is3d2_b200 import HostSession, shard

    cells, nev = args.sampler_cells, args.sampler_events
    surf = synthetic.s3d(cells, seed=3024 + rank, stress=0.3)
    root = tempfile.mkdtemp(prefix=f"is3d_smp_r{rank}_")
    try:
        workdir.make_workdir(root, SAMPLER_PARAMS, chosen="smash")
        with HostSession(root) as h:
            h.set_surface(surf)
            shard.set_global_thermo_averages(h)
            h.prepare()
            h.abi_set_surface(surf, global_offset=rank * cells)
            ntot, _ = h.abi_total_yield()
            h.abi_sample(nev)                                      # warm-up: same size, so the pinned list buffer is reused
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            parts, counts, st, release = h.abi_sample(nev, copy=False)       # view of the library-owned pinned list
            torch.cuda.synchronize()
            sec = time.perf_counter() - t0
            n_parts = len(parts)
            del parts
            release()
    finally:
        shutil.rmtree(root, ignore_errors=True)
    acc = torch.tensor([float(n_parts), float(st.sampler_proposals), sec, st.kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        sums = acc.clone()
        dist.all_reduce(sums)
        dist.all_reduce(acc, op=dist.ReduceOp.MAX)
        hadrons, proposals, sec, kms = float(sums[0]), float(sums[1]), float(acc[2]), float(acc[3])
    else:
        hadrons, proposals, sec, kms = (float(v) for v in acc.tolist())
    return {"metric": "sampled hadrons/s", "value": hadrons / sec, "unit": "hadrons/s", "hadrons": int(hadrons),
            "proposals_per_s": proposals / sec, "acceptance": hadrons / max(proposals, 1.0), "seconds": sec,
            "device_ms": kms, "d2h_bytes": int(hadrons) * 104, "mean_yield_per_event_rank0": ntot,
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


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from is3d2_b200 import HostSession, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this repository has no CPU compute path")
    torch.cuda.set_device(local)
    os.environ["IS3D_DEVICE"] = str(local)
    os.environ.setdefault("IS3D_FAMOD_CHAIN", "0")          # df_mode 5: chain-free (shardable) initial guesses
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
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        h.close()
        shutil.rmtree(root, ignore_errors=True)


def _measure_ours(args, h, world: int, rank: int, local: int):
    import torch
    import torch.distributed as dist

    from is3d2_b200 import shard

    cells = args.cells_per_gpu
    surf = synthetic.s3d(cells, seed=2024 + rank, baryon=True)         # rank's shard of the surface
    # thermodynamic averages (only the sampler uses them) from a small prefix: the host loop is O(cells) python-free C++
    h.set_surface({k: v[:1000] for k, v in surf.items()})
    shard.set_global_thermo_averages(h)
    h.prepare()
    shape = h.spectra_shape()
    total = int(np.prod(shape))
    evals_rank = float(cells) * total

    ext = torch.cuda.ExternalStream(h.lib.is3d_stream(h.ctx), device=torch.device("cuda", local))
    # pinned host copies (e2e path) and resident device copies (value path)
    host_cols = {k: torch.from_numpy(v).pin_memory() for k, v in surf.items()}
    dev_cols = {k: t.cuda(non_blocking=True) for k, t in host_cols.items()}
    torch.cuda.synchronize()
    out_dev = torch.zeros(total, dtype=torch.float64, device="cuda")
    h.abi_set_surface_device({k: t.data_ptr() for k, t in dev_cols.items()}, cells, global_offset=rank * cells)

    def step():
        st = h.abi_spectra_device(out_dev.data_ptr())
        if world > 1:
            dist.all_reduce(out_dev)                                      # the one collective of the path (NCCL / NVLink)
        return st

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
        kernel_ms, launches = 0.0, 0
        ev0.record(ext)
        for _ in range(args.steps):
            st = step()
            kernel_ms += st.kernel_ms
            launches += st.kernel_launches
        ev1.record(ext)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())

        # ---- end to end through the C ABI with HOST buffers: H2D of the 25 columns + compute + D2H of the spectra ----
        # N = 1: is3d_set_surface + is3d_spectra (host in, host out).  N > 1: is3d_set_surface + is3d_spectra_device,
        # one NCCL all-reduce of the device array, D2H into pinned memory -- the sequence INTEGRATION.md prescribes.
        host_np = {k: t_.numpy() for k, t_ in host_cols.items()}
        out_host = torch.empty(total, dtype=torch.float64).pin_memory()

        def e2e_step():
            h.abi_set_surface(host_np, global_offset=rank * cells)
            if world > 1:
                h.abi_spectra_device(out_dev.data_ptr())
                dist.all_reduce(out_dev)
                out_host.copy_(out_dev, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return out_host.numpy()
            return h.abi_spectra()[0]

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()                                                       # untimed: first-use allocations
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            spec = e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())

    fp64_peak = h.abi_fp64_peak()
    sampler = None if args.no_sampler else sampler_bench(args, rank, world, local)
    if rank == 0:
        ms_step = ms_total / args.steps
        value = evals_rank * world / (ms_step * 1e-3)
        kern_s = kernel_ms / args.steps * 1e-3
        achieved = evals_rank * F_ALG[args.df_mode] / kern_s / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        bytes_alg = cells * 25 * 8.0 + total * 8.0
        # DRAM bytes and FP64-pipe utilisation of ONE df_spectra_kernel launch from the committed `ncu --set full` capture of
        # this command at the default size (profiles/r01_ncu_k1_headline.json); null for any other configuration
        traffic = traffic_src = fp64_pct = executed = None
        try:
            cap = json.load(open(os.path.join(REPO, "profiles", "r01_ncu_k1_headline.json")))
            if cap["config"] == {"df_mode": args.df_mode, "cells_per_gpu": cells}:
                traffic = float(cap["dram_bytes_read"] + cap["dram_bytes_write"])
                traffic_src = "ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_ncu_k1_headline.json"
                fp64_pct = cap["fp64_pipe_active_pct"]
                # executed work: FP64-pipe instructions of the SASS inner loop x 2 flops, per class slot actually evaluated
                executed = (evals_rank * cap["class_slots"] / cap["species"]) * cap["fp64_instr_per_class_eval_sass"] * 2.0 / kern_s / 1e12
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": "Cooper-Frye cell*species*momentum evals/s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.df_mode, cells, world), "global_cells": cells * world,
                       "evals_per_step": evals_rank * world, "l2_policy": "inputs larger than L2 (cell packs: 256 B x cells per pass)",
                       "parallelism": f"cells sharded x{world}, one NCCL all-reduce of {total} doubles" if world > 1 else "single GPU",
                       "species_classes": "evaluations are counted per species (444); species with identical (mass, sign, baryon "
                                          "number) share one integrand, computed once and scaled by each species' degeneracy "
                                          "(193 classes for the SMASH list) -- every species' bins are delivered"},
            "e2e": {"value": evals_rank * world / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(cells * 25 * 8),
                    "d2h_bytes_per_step": int(total * 8)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "df_spectra_kernel" if args.df_mode <= 2 else "feqmod_spectra_kernel", "kernel_ms_per_step": kernel_ms / args.steps,
                         "fp64_pipe_active_pct_ncu": fp64_pct,
                         "executed_tflops": executed, "frac_executed": (executed / fp64_peak) if executed else None,
                         "flops_per_eval_algorithmic": F_ALG[args.df_mode],
                         "peak_source": "DFMA micro-benchmark run live by is3d_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 entry)",
                         "hbm_gbs_algorithmic": bytes_alg / kern_s / 1e9, "hbm_peak_gbs_measured": peaks.get("hbm_gbs")},
            "clocks": clocks,
        }
        if sampler is not None:
            line["sampler"] = sampler
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
            if sampler is not None:
                line["sampler"]["cpu_baseline"] = sampler_cpu_baseline(args)
        return line
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells-per-gpu", type=int, default=1_250_000)
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
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
