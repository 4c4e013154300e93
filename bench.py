#!/usr/bin/env python
"""bench.py -- likelihood+gradient evaluations/s of the fused population likelihood.

    python bench.py --gpus N --steps K --warmup W [--workload cfg3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one fused evaluation of log L and d log L / d Lambda over the whole synthetic catalog
(all events x samples + all found injections).  Default workload: BASELINE.json configs[2]
("B-spline full model, 300 events x 10k samples, 1e8 injections") -- the configuration the
north-star's ">= 70 % of HBM roofline on 1 B200 / >= 85 % scaling at 8 GPUs" is quoted on; its
plan (7.4 GB) is far larger than L2, so consecutive timed steps cannot be served from cache.
Multi-GPU: injections sharded by (m1, q) piece bucket, whole events dealt round-robin, one NCCL
all-gather of ~4 KB partial records per step; total work is fixed => "scaling": "strong".

PyTorch is used only as plumbing (NCCL all-gather, CUDA events, device scratch tensors); every
kernel in the timed region is from libgwi.so.  ``--impl reference`` times the CPU arm: the plain-C
oracle port of the reference algorithm (oracle/c/gwi_oracle.c) on all host cores, on a bounded
sample (the reference itself needs JAX, which this image does not have).
"""

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "likelihood+grad evals/sec"
UNIT = "evals/s"
ALG_BYTES_PER_SAMPLE = 64  # 8 fp64 columns read once per evaluation


def _config(name, family, E, S, I, n_params, chains, world, shard_by, l2_policy):
    """The `config` object of BOTH arms (identical keys: the driver compares them)."""
    from gwinferno_b200 import workloads

    return {
        "workload": f"{name}: BASELINE.json configs[{workloads.WORKLOADS[name][0]}]", "model": family, "E": E, "S": S, "I": I,
        "n_params": n_params, "samples_per_eval": E * S + I, "chains_per_step": chains,
        "parallelism": ("single" if world == 1 else (f"shard{world}: injections by (m1, q) piece bucket, whole events round-robin, records exchanged through NVLink peer memory by libgwi"
                                                     if shard_by == "bucket" else f"shard{world}: injections by index range")),
        "l2_policy": l2_policy,
        "neff_grad": False, "likelihood": "marginalize_selection=False, min_neff_cut=True (reference defaults)",
    }


def _l2_policy(name):
    from gwinferno_b200 import workloads

    _, _, E, S, I = workloads.WORKLOADS[name]
    plan_gb = 72.0 * (E * S + I) / 1e9
    return "inputs larger than L2 (plan ~%.2f GB in total)" % plan_gb if plan_gb > 0.26 else "inputs fit in L2; new Lambda every step, no flush"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.3)  # nvidia-smi buffers its first lines; give it time to flush at least a few samples
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def _cpu_arm(name, seconds_per_step, steps, warmup):
    """The CPU implementation of the path: the plain-C oracle (oracle/c/gwi_oracle.c, a port of the
    reference algorithm working from raw coordinates; the reference itself needs JAX, which is not
    installed) on ALL host cores, on a bounded sample of the workload -- every event of the first
    ``n_ev`` and the first ``n_inj`` found injections, sized from a calibration run so that one step
    takes about ``seconds_per_step`` -- extrapolated linearly in samples to the full workload.
    Each step evaluates a different Lambda and includes the likelihood glue (analysis.py:257-319)."""
    from gwinferno_b200 import lowering, synthetic, workloads
    from oracle import c_oracle, popmodel

    cfg_idx, family, E, S, I = workloads.WORKLOADS[name]
    cfg = 2 if name == "cfg4" else int(name[3:])
    cores = os.cpu_count() or 1
    n_full = E * S + I

    def build(n_ev, n_inj):
        pe_all = synthetic.make_events(E, S, 1000 * cfg + 0)
        zpe = (float(pe_all["redshift"].min()), float(pe_all["redshift"].max()))
        pe = {k: np.ascontiguousarray(v[:n_ev]) for k, v in pe_all.items()}
        inj = synthetic.make_injections(I, 1000 * cfg + 1, 0, n_inj)
        z_range = (max(zpe[0], synthetic.ZLO), min(zpe[1], synthetic.ZHI))
        weights, params_fn = workloads.build_model(family, pe, inj, z_range=z_range)
        low, lam, _ = workloads.lower_workload(weights, params_fn, pe, inj)
        return low, weights, params_fn, pe, n_ev * S + n_inj

    def one(low, lam, n_ev):
        ev = c_oracle.evaluate(low.spec, low.pe_cols, low.inj_cols, 4.0 * I, lam, want_jac=True, want_neff_jac=False, n_threads=cores)
        return popmodel.hierarchical_log_likelihood(ev, n_ev, min_neff_cut=True)[0]

    # calibration on a small sample
    n_ev0, n_inj0 = min(E, 2 * cores), min(I, 200_000)
    low, weights, params_fn, pe, n0 = build(n_ev0, n_inj0)
    lam = lowering.flatten_params(weights(pe, True, params_fn(0)), low.spec.n_params)
    one(low, lam, n_ev0)  # warm-up: cosmology table, page faults
    t0 = time.perf_counter()
    one(low, lam, n_ev0)
    rate = n0 / (time.perf_counter() - t0)  # samples per second on all cores
    frac = min(1.0, rate * seconds_per_step / n_full)
    n_ev, n_inj = max(1, int(round(E * frac))), max(1000, int(round(I * frac)))
    low, weights, params_fn, pe, n_sample = build(n_ev, n_inj)
    times = []
    for s_ in range(warmup + steps):
        lam = lowering.flatten_params(weights(pe, True, params_fn(s_)), low.spec.n_params)
        t0 = time.perf_counter()
        one(low, lam, n_ev)
        times.append(time.perf_counter() - t0)
    t = float(np.sum(times[warmup:]))
    value = (steps / t) * n_sample / n_full
    return {
        "value": value, "unit": UNIT, "cores": cores, "kind": "port", "n_params": int(low.spec.n_params),
        "sample": f"{n_ev} events x {S} samples + {n_inj} injections ({n_sample} of {n_full} samples), {1e3 * t / steps:.0f} ms per sampled step on {cores} threads, "
                  "extrapolated linearly in samples; plain-C port of the reference algorithm (oracle/c/gwi_oracle.c; the reference needs JAX, not installed)",
        "samples_per_s": n_sample * steps / t,
    }, t, n_sample, n_full


def cpu_baseline(name):
    """cpu_baseline object of the GPU arm's line: about 10 s of CPU work on rank 0."""
    return _cpu_arm(name, seconds_per_step=2.0, steps=4, warmup=1)[0]


def parity_at_size(name, scale, lam, res, world):
    """{rel_log_l, rel_grad}: GPU result `res` (header + gradient) vs oracle/c/gwi_oracle.c on the whole catalog."""
    from gwinferno_b200 import capi, workloads
    from oracle import c_oracle, popmodel

    t0 = time.perf_counter()
    pe, inj, const, z_range = workloads.shard_catalog(name, 0, 1, scale=scale, all_reduce_minmax=lambda lo, hi: (lo, hi))
    weights, params_fn = workloads.build_model(const["family"], pe, inj, z_range=z_range)
    low, _, _ = workloads.lower_workload(weights, params_fn, pe, inj)
    cores = os.cpu_count() or 1
    ev = c_oracle.evaluate(low.spec, low.pe_cols, low.inj_cols, const["total_inj"], lam, want_jac=True, want_neff_jac=False, n_threads=cores)
    l_o, g_o, _ = popmodel.hierarchical_log_likelihood(ev, const["E"], min_neff_cut=True)
    g = np.asarray(res[capi.GWI_LIKE_HEADER:])
    return {"rel_log_l": float(abs(res[0] - l_o) / abs(l_o)), "rel_grad": float(np.max(np.abs(g - g_o)) / np.max(np.abs(g_o))), "log_l_oracle": float(l_o),
            "tolerance": {"rel_log_l": 1e-10, "rel_grad": 1e-8}, "oracle": f"oracle/c/gwi_oracle.c, whole catalog, {cores} threads", "n_gpus": world,
            "seconds": time.perf_counter() - t0}


def nuts_block(device, n_chains=16):
    """The second half of BASELINE.json's metric: NUTS ESS/s on configs[1] (70 x 4000 + 5e5 injections, 164 free parameters),
    the whole sampler loop in native code (csrc/nuts.cpp: multinomial NUTS, windowed adaptation, dense mass matrix -- NumPyro's
    NUTS(dense_mass=True) in spirit) around gwi_loglike_host; priors of examples/simple_bspline_example.py.  The CPU row is the
    same chain priced at the plain-C oracle's evaluation rate on this host (measured on a few evaluations, not run in full)."""
    from gwinferno_b200 import nuts, pipeline, workloads
    from gwinferno_b200.likelihood import PopulationLikelihood
    from oracle import c_oracle

    pe, inj, const, z_range = workloads.shard_catalog("cfg2", 0, 1)
    weights, params_fn = workloads.build_model(const["family"], pe, inj, z_range=z_range)
    low, lam0, p0 = workloads.lower_workload(weights, params_fn, pe, inj)
    eng = PopulationLikelihood(low, const["total_inj"], device=device)
    blocks = pipeline.bspline_prior_blocks(low.slots_for, p0)
    dim = low.spec.n_params - 1
    theta0 = 0.1 * np.random.default_rng(0).standard_normal(dim)
    n_warm, n_samp, flags = 1000, 800, 7
    t0 = time.perf_counter()
    samples, info = nuts.nuts_native(eng, blocks, theta0, n_warm, n_samp, Nobs=const["E"], seed=0, max_depth=8, flags=flags)
    wall = time.perf_counter() - t0
    eng.model.close()
    ess = np.array([nuts.effective_sample_size(samples[:, i]) for i in range(samples.shape[1])])
    # the same sampler, K chains advanced together: one batched evaluation (gwi_loglike_batch_host) per round of leapfrog
    # steps -- the reference's MCMC(chain_method="vectorized") on one GPU.  ESS = sum of the per-chain ESS.
    chains_obj = None
    if n_chains > 1:
        hint = max(1, n_chains // 2)  # two alternating groups of chains: one on the GPU, one doing its host arithmetic
        engK = PopulationLikelihood(low, const["total_inj"], device=device, batch_hint=hint)
        th0 = 0.1 * np.random.default_rng(1).standard_normal((n_chains, dim))
        t0 = time.perf_counter()
        sK, infK = nuts.nuts_native_chains(engK, blocks, th0, n_warm, n_samp, Nobs=const["E"], seed=100, max_depth=8, flags=flags)
        wallK = time.perf_counter() - t0
        engK.model.close()
        essK = np.array([sum(nuts.effective_sample_size(sK[c, :, i]) for c in range(n_chains)) for i in range(dim)])
        rhat = np.array([nuts.split_rhat(sK[:, :, i]) for i in range(dim)])
        t_samp = max(i["sampling_seconds"] for i in infK)
        chains_obj = {"n_chains": n_chains, "warmup": n_warm, "samples_per_chain": n_samp,
                      "batch_hint": hint,
                      "driver": "gwi_nuts_sample_posterior_chains: one host thread per chain, batch_hint chains per gwi_loglike_batch_host call (two alternating groups: the GPU evaluates one while the other does its host arithmetic)",
                      "ess_min": float(essK.min()), "ess_median": float(np.median(essK)), "ess_min_per_s": float(essK.min() / t_samp),
                      "ess_median_per_s": float(np.median(essK) / t_samp), "ess_min_per_wall_s": float(essK.min() / wallK),
                      "grad_evals_per_s": float(sum(i["leapfrogs_sampling"] for i in infK) / t_samp), "split_rhat_max": float(np.nanmax(rhat)),
                      "mean_accept": float(np.mean([i["mean_accept"] for i in infK])), "step_size": [float(i["step_size"]) for i in infK],
                      "sampling_seconds": float(t_samp), "wall_s": wallK}
    cores = os.cpu_count() or 1
    c_oracle.evaluate(low.spec, low.pe_cols, low.inj_cols, const["total_inj"], lam0, want_jac=True, want_neff_jac=False, n_threads=cores)
    t0 = time.perf_counter()
    n_cpu = 5
    for _ in range(n_cpu):
        c_oracle.evaluate(low.spec, low.pe_cols, low.inj_cols, const["total_inj"], lam0, want_jac=True, want_neff_jac=False, n_threads=cores)
    cpu_rate = n_cpu / (time.perf_counter() - t0)
    gpu_rate = info["leapfrogs_sampling"] / info["sampling_seconds"]
    return {"workload": "cfg2: BASELINE.json configs[1]", "dim": dim, "warmup": n_warm, "samples": n_samp, "max_tree_depth": 8,
            "sampler": "csrc/nuts.cpp: multinomial NUTS + windowed adaptation + dense mass (flags 7), likelihood+gradient via gwi_loglike_host",
            "ess_min": float(ess.min()), "ess_median": float(np.median(ess)), "ess_min_per_s": float(ess.min() / info["sampling_seconds"]),
            "ess_median_per_s": float(np.median(ess) / info["sampling_seconds"]), "grad_evals_per_s": float(gpu_rate), "mean_accept": info["mean_accept"],
            "step_size": info["step_size"], "wall_s": wall, "chains": chains_obj,
            "cpu": {"grad_evals_per_s": float(cpu_rate), "cores": cores, "kind": "port (oracle/c/gwi_oracle.c), priced not run: the same chain at this evaluation rate",
                    "ess_min_per_s": float(ess.min() / (info["leapfrogs_sampling"] / cpu_rate)), "ess_median_per_s": float(np.median(ess) / (info["leapfrogs_sampling"] / cpu_rate))}}


def run_reference(args):
    """CPU arm (``--impl reference``): same metric, config and unit as the GPU arm, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gwinferno_b200 import workloads

    name = args.workload
    _, family, E, S, I = workloads.WORKLOADS[name]
    t0 = time.perf_counter()
    # bounded steps: the whole run ends within a few minutes whatever --steps is
    per_step = max(0.25, min(3.0, 150.0 / max(1, args.steps + args.warmup)))
    base, t, n_sample, n_full = _cpu_arm(name, seconds_per_step=per_step, steps=args.steps, warmup=args.warmup)
    value = base["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps * n_full / n_sample, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(name, family, E, S, I, base["n_params"], workloads.N_CHAINS.get(name, 1), args.gpus, "bucket", _l2_policy(name)),
        "host": f"{base['cores']} host threads (CPU arm: no GPU in this line)",
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "setup_s": time.perf_counter() - t0 - t,
    }
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from gwinferno_b200 import capi, lowering, workloads
    from gwinferno_b200.likelihood import PopulationLikelihood

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def allreduce_minmax(lo, hi):
        t = torch.tensor([lo, -hi], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t[0]), float(-t[1])

    name = args.workload
    t_setup = time.perf_counter()
    if args.emulate_world > 1 and world == 1:
        # tuning aid: time ONE rank's shard of a W-way run on a single GPU (no collective)
        pe, inj, const, z_range = workloads.shard_catalog(name, 0, args.emulate_world, scale=args.scale, all_reduce_minmax=lambda lo, hi: (lo, hi), shard_by=args.shard_by)
    else:
        pe, inj, const, z_range = workloads.shard_catalog(name, rank, world, scale=args.scale, all_reduce_minmax=allreduce_minmax, shard_by=args.shard_by)
    weights, params_fn = workloads.build_model(const["family"], pe, inj, z_range=z_range)
    low, lam0, _ = workloads.lower_workload(weights, params_fn, pe, inj)
    P = low.spec.n_params
    t_gen = time.perf_counter() - t_setup
    eng = PopulationLikelihood(low, const["total_inj"], device=local, need_neff_grad=False, chunk_steps=args.chunk_steps, n_deep=args.n_deep,
                               batch_hint=workloads.N_CHAINS.get(name, 1), catalog_on_device=args.catalog_on_device)
    t_plan = time.perf_counter() - t_setup - t_gen  # (with --catalog-on-device this includes the up-front column upload; gwi_model_create alone is in setup_s)
    info = eng.info()
    mdl = eng.model
    Nobs = const["E"]
    n_steps = args.warmup + args.steps
    chains = workloads.N_CHAINS.get(name, 1)  # cfg4: every step evaluates `chains` Lambda vectors
    if chains > 1 and world > 1:
        raise SystemExit("the chain-batched workload shards trivially (disjoint chain ranges per GPU); run it with --gpus 1")
    lams = np.stack([lowering.flatten_params(weights(pe, True, params_fn(s)), P) for s in range(n_steps)])
    if chains > 1:
        rng = np.random.default_rng(4)
        lam_chain = torch.from_numpy(lams[0][None, :] * (1.0 + 0.02 * rng.standard_normal((chains, P)))).cuda()
        out_chain = torch.zeros(chains * (capi.GWI_LIKE_HEADER + P), dtype=torch.float64, device="cuda")
    lam_dev = torch.from_numpy(lams).cuda()
    lam_pin = torch.from_numpy(lams).pin_memory()
    out = torch.zeros(capi.GWI_LIKE_HEADER + P, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    if world > 1:
        # the exchange belongs to libgwi (peer-memory pushes + flags inside its last kernel); torch.distributed only
        # carries the 80-byte IPC handles once, here, and the barriers around the timed region
        handles = [None] * world
        dist.all_gather_object(handles, mdl.comm_local_handle(world))
        mdl.comm_connect(handles, rank)
        dist.barrier()

    def step(i, lam_ptr=None):
        ptr = lam_dev[i].data_ptr() if lam_ptr is None else lam_ptr
        if chains > 1:
            mdl.loglike_batch_ptr(lam_chain.data_ptr(), chains, out_chain.data_ptr(), Nobs, stream=stream)
        elif world == 1:
            mdl.loglike(ptr, out.data_ptr(), Nobs, stream=stream)
        else:
            mdl.loglike_sharded(ptr, out.data_ptr(), Nobs, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None  # samples clocks through warm-up + timed region
    for i in range(args.warmup):
        step(i)
    barrier()
    mdl.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.flush_l2:
        # cold numbers for the L2-resident configurations: a 512 MB write between steps evicts the plan
        # from the 126 MB L2; only the steps themselves are timed (one event pair per step)
        scrub = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        pairs = []
        for i in range(args.warmup, n_steps):
            scrub.fill_(i & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step(i)
            b.record()
            pairs.append((a, b))
        barrier()
        ms = float(sum(a.elapsed_time(b) for a, b in pairs))
        del scrub
    else:
        e0.record()
        for i in range(args.warmup, n_steps):
            step(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    kt = mdl.stream_times_ms(min(64, args.steps))
    mdl.set_timing(False)
    res = (out_chain[: capi.GWI_LIKE_HEADER + P] if chains > 1 else out).cpu().numpy()
    per_rank = None
    if world > 1:
        # per-rank spread (the exchange makes every rank wait for the slowest one): own timed loop, own stream kernel, own shard
        mine = {"rank": rank, "ms_per_step": ms / args.steps, "kernel_ms": float(np.mean(kt)) if len(kt) else None,
                "samples": int(info["n_samples_pe"] + info["n_samples_inj"])}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    # keep the same work running (untimed, every rank the same number of steps) until the 100 ms
    # clock sampler has seen the GPU under this load
    n_extra = int(min(5000, max(1, 0.7 / max(1e-6, ms / args.steps * 1e-3))))
    for k in range(n_extra):
        step(args.warmup + k % args.steps)
    barrier()
    clocks = sampler.stop() if sampler else None
    # ---- end-to-end: host Lambda in, host (log L, gradient) out, every step ------------------
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, n_steps):
        if chains > 1:
            d = lam_chain.cpu().pin_memory().to("cuda", non_blocking=True)
            mdl.loglike_batch_ptr(d.data_ptr(), chains, out_chain.data_ptr(), Nobs, stream=stream)
            host = out_chain.cpu()
        elif world == 1:
            head, grad = mdl.loglike_host(lams[i], Nobs)
        else:
            d = lam_pin[i].to("cuda", non_blocking=True)
            step(i, d.data_ptr())
            host = out.cpu()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t[0])
    # ---- the same through the Python drop-in (hierarchical_likelihood on lazy weights), single GPU ----
    e2e_python = None
    if world == 1 and chains == 1:
        from gwinferno_b200 import likelihood as L

        pw, iw = weights(pe, True, params_fn(0)), weights(inj, False, params_fn(0))
        keys, pattern = lowering._structure(pw, iw)
        L._ENGINES[(keys, pattern, float(const["total_inj"]), False, int(local))] = eng  # this bench's resident plan, not a second one
        n_py = min(args.steps, 20)
        t0 = time.perf_counter()
        for i in range(n_py):
            p_i = params_fn(args.warmup + i)
            r_py = L.hierarchical_likelihood(weights(pe, True, p_i), weights(inj, False, p_i), const["total_inj"], Nobs, 1.0)
        t_py = time.perf_counter() - t0
        e2e_python = {"value": n_py / t_py, "unit": UNIT, "steps": n_py, "api": "gwinferno_b200.likelihood.hierarchical_likelihood (model objects -> lazy weights -> one libgwi call; per-event sites read back)",
                      "log_l": r_py.log_likelihood}
        L._ENGINES.clear()  # the engine is closed by this function, not by the cache
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_local = info["n_samples_pe"] + info["n_samples_inj"]
    n_total = const["E"] * const["S"] + const["I"]
    peak, peak_src = _peaks()
    k_ms = float(np.mean(kt)) if len(kt) else float("nan")
    achieved = n_local * chains * ALG_BYTES_PER_SAMPLE / (k_ms * 1e-3) / 1e9  # a batched launch streams the plan once per chain (from L2)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{name}_n{world}")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": chains * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(name, const["family"], const["E"], const["S"], const["I"], P, chains, world, args.shard_by,
                          "L2 flushed between steps (512 MB write; one CUDA-event pair per step)" if args.flush_l2 else _l2_policy(name)),
        "samples_per_s": n_total * chains * args.steps / (ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel": "stream_kernel", "kernel_ms": k_ms, "alg_bytes_per_launch": n_local * chains * ALG_BYTES_PER_SAMPLE,
                     "actual_bytes_per_launch": info["bytes_per_eval"], "kernel_share_of_step": k_ms / (ms / args.steps)},
        "e2e": {"value": chains * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * P * chains, "d2h_bytes_per_step": 8 * (capi.GWI_LIKE_HEADER + P) * chains,
                "api": "gwi_loglike_host (C-ABI, host buffers: Lambda H2D, evaluation, log L + gradient D2H, synchronise)" if world == 1 and chains == 1 else "pinned Lambda H2D + libgwi device call + result D2H"},
        "e2e_python": e2e_python,
        "gpu_launches": info["kernel_launches_per_eval"] * args.steps,  # a chain batch launches every kernel once
        "clocks": clocks,
        "result": {"log_l": float(res[0]), "passed": float(res[1]), "status": float(res[7])},
        "setup_s": {"generate": t_gen, "plan_build_and_upload": t_plan, "plan_on_device": info["plan_on_device"], "gwi_model_create": info["plan_seconds"]},
        "per_rank": per_rank,
        "plan": {k: info[k] for k in ("n_valid_pe", "n_valid_inj", "n_padded", "n_chunks", "n_stream_columns", "n_spline_dims", "n_deep", "grid_blocks", "block_threads")},
        # tuning switches in effect (all default off: 0 / empty in the product configuration)
        "switches": {"active": info["active_switches"], "library": os.path.basename(capi.LIB_PATH),
                     "env": {k: v for k, v in os.environ.items() if k.startswith(("GWI_TUNE_", "GWI_EXP_", "GWI_FUSED", "GWI_GRAPH", "GWI_SPLIT", "GWI_PLAN"))}},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(name)
    if not args.no_cpu_baseline and chains == 1:
        # parity at the size that was timed: the last timed step's log L and gradient against the plain-C oracle on the
        # WHOLE catalog (rank 0 regenerates it when the run is sharded), same Lambda
        line["parity_at_size"] = parity_at_size(name, args.scale, lams[n_steps - 1], res, world)
    if world == 1 and not args.no_cpu_baseline and not args.no_nuts:
        line["nuts"] = nuts_block(local, args.nuts_chains)
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = None


def _protect_stdout():
    """stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 on their own
    (NCCL prints its version banner there when NCCL_DEBUG is set in the environment) are sent to
    stderr instead, and the result line goes to the original stdout."""
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(text):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink S and I (testing only; the reported config then differs from BASELINE's)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (cpu_baseline, parity_at_size, nuts): tuning runs")
    ap.add_argument("--no-nuts", action="store_true", help="skip the NUTS ESS/s block")
    ap.add_argument("--nuts-chains", type=int, default=16, help="chains of the multi-chain NUTS run (advanced together, batched evaluations); 1 = skip it")
    ap.add_argument("--flush-l2", action="store_true", help="cold numbers: evict L2 between steps (for the L2-resident configurations)")
    ap.add_argument("--emulate-world", type=int, default=1, help="tuning aid: run rank 0's shard of a W-way partition on one GPU")
    ap.add_argument("--shard-by", default="bucket", choices=["bucket", "index"], help="multi-GPU partition of the found injections")
    ap.add_argument("--catalog-on-device", action="store_true",
                    help="place the sample columns in device memory before gwi_model_create (a caller whose arrays already live on the GPU): the plan build then moves nothing over PCIe")
    ap.add_argument("--n-deep", type=int, default=-1, help="tuning experiment: spline dims with lane-private accumulators (-1 = auto)")
    ap.add_argument("--chunk-steps", type=int, default=0, help="tuning experiment: samples per lane per chunk (0 = auto)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
