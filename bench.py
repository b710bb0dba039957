#!/usr/bin/env python
"""Headline benchmark: utterances/sec of the reverse-diffusion enhancement path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this framework (B200, native CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host CPU (oracle port)

A "step" = one pass of the hot path over one batch: ScoreModel.enhance for B = 8 utterances per GPU of
32 640 samples (-> [B,1,256,256] complex spectrograms), N = 30 predictor-corrector steps with
reverse_diffusion + ald x1 (60 network evaluations), ncsnpp 27.8 M with seeded random weights
(init_scale = 1.0; the default init is degenerate).  Weak scaling: B is per GPU.

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM (CUDA events, max over ranks);
`e2e` goes through ScoreModel.enhance_batch with host buffers (pinned H2D + D2H inside the timed region);
`roofline` is the live CUDA-event measurement of the convolution kernels inside the timed steps;
`cpu_baseline` times the oracle port of the reference algorithm on this box's host cores on a bounded
sample.  Multi-GPU: one process per GPU (torchrun), one NCCL broadcast of the weight blob at load, no
per-step collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Workloads = BASELINE.json `configs` (SURVEY.md section 8d restates them).  The default line is configs[1]; the others are
# selected with --config and are what BASELINE.md section 4 quotes.  gflop / gb: algorithmic work per utterance per network
# evaluation.
WORKLOADS = {
    "headline": dict(tag="BASELINE configs[1]", backbone="ncsnpp", B_per_gpu=8, L=32640, F=256, T=256, N=30, corrector="ald",
                     corrector_steps=1, predictor="reverse_diffusion", snr=0.5, gflop=624.8, gb=0.879, params="27.8M"),
    "4s": dict(tag="literal 4 s row of SURVEY 8(d)", backbone="ncsnpp", B_per_gpu=8, L=64000, F=256, T=512, N=30, corrector="ald",
               corrector_steps=1, predictor="reverse_diffusion", snr=0.5, gflop=1251.8, gb=1.648, params="27.8M"),
    "large8s": dict(tag="BASELINE configs[3]", backbone="ncsnpplarge", B_per_gpu=32, L=128000, F=256, T=1024, N=50, corrector="ald",
                    corrector_steps=1, predictor="reverse_diffusion", snr=0.5, gflop=2131.0, gb=3.721, params="65.6M"),
    "sweep": dict(tag="BASELINE configs[4]", backbone="ncsnpp", B_per_gpu=16, L=32640, F=256, T=256, N=30, corrector="ald",
                  corrector_steps=1, predictor="reverse_diffusion", snr=0.5, gflop=624.8, gb=0.879, params="27.8M"),
    "storm": dict(tag="StoRM mode (SURVEY 8(f)-1)", backbone="ncsnpp", B_per_gpu=8, L=32640, F=256, T=256, N=30, corrector="none",
                  corrector_steps=1, predictor="reverse_diffusion", snr=0.5, gflop=624.8, gb=0.879, params="2 x 27.8M"),
}
WORKLOAD = WORKLOADS["headline"]
GFLOP_PER_UTT_NFE = WORKLOAD["gflop"]
GB_PER_UTT_NFE = WORKLOAD["gb"]


def nfe_of(W) -> int:
    return W["N"] * ((W["corrector_steps"] if W["corrector"] != "none" else 0) + 1)


def workload_config(world: int, name: str = "headline", nfe: int = None) -> dict:
    """Same `config` object in both arms (native / --impl reference)."""
    W = WORKLOADS[name]
    nfe = nfe_of(W) if nfe is None else nfe
    return dict(workload=f"{W['tag']}: {W['backbone']} {W['params']}, B={W['B_per_gpu']}/GPU, {W['L']} samples -> {W['F']}x{W['T']} complex "
                         f"STFT, N={W['N']} PC ({W['predictor']}+{W['corrector']}x{W['corrector_steps']}, {nfe} NFE), random-init weights "
                         "init_scale=1.0", name=name,
                global_batch=world * W["B_per_gpu"], parallelism=f"dp{world} (utterance sharding, one NCCL weight broadcast)",
                l2="working set per step (activations >= 3 GB) >> 126 MB L2; no explicit flush", nfe=nfe,
                tflop_per_step=world * W["B_per_gpu"] * nfe * W["gflop"] / 1e3)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d.get("bf16_tflops"), bf16_sustained=d.get("bf16_tflops_sustained"), hbm=d.get("hbm_gbs"),
                    src="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    power_w_max=max(pw) if pw else None, samples=len(self.rows))


def usable_cpus() -> int:
    """Host cores this process may really use: affinity mask and cgroup CPU quota, not os.cpu_count()."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(-(-int(q) // int(per)))))
    except Exception:  # noqa: BLE001
        pass
    return n


def pick_cpu_threads() -> int:
    """Thread count that makes the CPU oracle fastest on this box (oversubscribing a quota-limited container
    slows torch's conv by an order of magnitude, so 'all threads' is calibrated, not assumed)."""
    import torch
    import torch.nn.functional as F
    n = usable_cpus()
    cands = sorted({c for c in (n, n // 2, 64, 32, 16, 8) if 1 <= c <= n}, reverse=True)
    x = torch.randn(1, 128, 128, 128)
    w = torch.randn(128, 128, 3, 3)
    best, best_t = cands[-1], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        F.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            F.conv2d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def synth_batch(B, L, first_id):
    import torch
    from oracle.storm_oracle import synth_utterance
    return torch.stack([synth_utterance(first_id + b, L) for b in range(B)])


# ---------------------------------------------------------------------------------------------------------
def cpu_oracle_rate(threads, nfe_target, W=None, pc_steps=1):
    """utterances/sec of the reference algorithm (oracle port) on the host: time `pc_steps` PC steps of ONE utterance of the
    workload's shape and scale linearly in NFE (every step has identical shapes)."""
    import torch
    from oracle import storm_oracle as O
    from storm_b200 import ncsnpp_spec
    W = W or WORKLOAD
    torch.set_num_threads(threads)
    ocfg = O.named_cfg(W["backbone"])
    cfg = ncsnpp_spec.NCSNppConfig(nf=ocfg.nf, ch_mult=ocfg.ch_mult, num_res_blocks=ocfg.num_res_blocks, attn_resolutions=ocfg.attn_resolutions)
    sd = ncsnpp_spec.random_state_dict(cfg, seed=0, init_scale=1.0)
    gen = torch.Generator().manual_seed(0)
    Y = torch.complex(torch.randn(1, 1, W["F"], W["T"], generator=gen), torch.randn(1, 1, W["F"], W["T"], generator=gen)) * 0.3
    score = O.make_score_fn(sd, ocfg)
    with torch.no_grad():
        t0 = time.perf_counter()
        _, nfe = O.pc_sample(score, Y, O.OUVE(), pc_steps, snr=W["snr"])
        dt = time.perf_counter() - t0
    per_nfe = dt / nfe
    return 1.0 / (per_nfe * nfe_target), dt, nfe


def run_reference(args):
    """--impl reference: the reference's algorithm on the host CPU (the reference is Python and cannot travel to
    the GPU box, so this is the oracle port pinned to it by tests/golden)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = WORKLOADS[args.config]
    threads = pick_cpu_threads()
    nfe_target = nfe_of(W)
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt, nfe = cpu_oracle_rate(threads, nfe_target, W)
        if i >= args.warmup:
            rates.append((r, dt))
    value = sum(r for r, _ in rates) / len(rates)
    sample = (f"B=1, {W['F']}x{W['T']}, 1 PC step (2 NFE) per bench step, utt/s extrapolated linearly in NFE to {nfe_target} NFE; the CPU arm "
              "runs ONE utterance, it does not batch or shard")
    cfg = workload_config(args.gpus, args.config)
    cfg["reference_config"] = f"B=1, 1 PC step of the {W['F']}x{W['T']} shape on the host CPU, extrapolated linearly to {nfe_target} NFE"
    line = dict(metric="utterances/sec", value=value, unit="utt/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * sum(dt for _, dt in rates) / len(rates), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=cfg,
                cpu_baseline=dict(value=value, unit="utt/s", cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit="utt/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def path_traffic():
    """DRAM bytes of ONE whole network evaluation (every kernel), from the committed ncu capture of this build
    (profiles/r2_path_traffic.json, written by tools/path_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`)."""
    tp = os.path.join(ROOT, "profiles", "r2_path_traffic.json")
    if os.path.exists(tp):
        return json.load(open(tp))
    return None


def run_native(args):
    import torch
    import torch.distributed as dist
    from storm_b200 import ScoreModel, StochasticRegenerationModel, ncsnpp_spec, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    name = args.config
    W = WORKLOADS[name]
    B, L = W["B_per_gpu"], W["L"]
    storm = name == "storm"
    if storm:
        model = StochasticRegenerationModel(backbone_denoiser=W["backbone"], backbone_score=W["backbone"], sde="ouve", init_scale=1.0, seed=42)
        nets = model._nets()
    else:
        model = ScoreModel(backbone=W["backbone"], sde="ouve", init_scale=1.0, seed=42)
        nets = [model.dnn]
    # ---- weights: rank 0 owns the seeded state, ONE NCCL broadcast of the flat blob per network, no later collective
    if rank == 0:
        for i, net in enumerate(nets):
            net.load_state_dict(ncsnpp_spec.random_state_dict(net.cfg, seed=i, init_scale=1.0))
    model.cuda()
    for net in nets:
        parallel.broadcast_weights(net, src=0)      # no-op at world == 1
    eng = nets[-1].engine()                          # the score network carries ~all of the work
    engines = [n.engine() for n in nets]
    sp = model.data_module.spec_config()
    ids = list(range(rank * B, rank * B + B))
    y_host = synth_batch(B, L, rank * B).pin_memory()
    y_dev = y_host.cuda()
    kw = dict(predictor=W["predictor"], corrector=W["corrector"], N=W["N"], corrector_steps=W["corrector_steps"], snr=W["snr"])
    nfe_step = nfe_of(W) + (1 if storm else 0)       # StoRM: + one denoiser evaluation

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sweep_rows = None
    if storm:
        def step_dev():
            return model.enhance_batch(y_dev, seed=42, utt_ids=ids, return_nfe=True, **kw)
    elif name == "sweep":
        def step_dev():
            return sweep_once(model, y_dev, ids, W)
    else:
        sc = model._sampler_cfg(W["predictor"], W["corrector"], W["N"], W["corrector_steps"], W["snr"])

        def step_dev():
            return eng.enhance(y_dev, sc, sp, seed=42, utt_ids=ids)

    for _ in range(args.warmup):
        step_dev()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for e in engines:
        e.reset_launch_count()
    # ---- `value`: device-resident inputs, CUDA events, NO per-launch instrumentation inside the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        res = step_dev()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(e.launch_count() for e in engines)
    if name == "sweep":
        sweep_rows = res
        out = None
    else:
        out, nfe = res
        assert torch.isfinite(out).all() and nfe == nfe_of(W)

    # ---- end to end through the public API with host buffers (pinned H2D + D2H inside the timed region)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if name == "sweep":
            r = sweep_once(model, y_host, ids, W, host=True)
        else:
            r = model.enhance_batch(y_host, seed=42, utt_ids=ids, **kw)
            assert r.device.type == "cpu"
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop() if rank == 0 else None

    # ---- separate profiled pass (two CUDA events around every conv launch): the live roofline measurement of the
    # dominant kernel class; its instrumentation never touches `value`
    for e in engines:
        e.set_profile(True)
    barrier()
    p0 = time.perf_counter()
    step_dev()
    barrier()
    prof_ms = (time.perf_counter() - p0) * 1e3
    prof = {"tcgen05": dict(launches=0, ms=0.0, flops=0.0), "simt": dict(launches=0, ms=0.0, flops=0.0)}
    for e in engines:
        pe = e.get_profile()
        for k in prof:
            for kk in prof[k]:
                prof[k][kk] += pe[k][kk]
        e.set_profile(False)

    if world > 1:
        tt = torch.tensor([ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    total_utts = world * B * args.steps
    value = total_utts / (ms * 1e-3)
    e2e_value = total_utts / (e2e_ms * 1e-3)
    dom = "tcgen05" if prof["tcgen05"]["flops"] >= prof["simt"]["flops"] else "simt"
    dp = prof[dom]
    achieved = dp["flops"] / (dp["ms"] * 1e-3) / 1e12 if dp["ms"] > 0 else 0.0
    conv_share = (prof["tcgen05"]["ms"] + prof["simt"]["ms"]) / prof_ms
    peak = pk["bf16_sustained"]
    # operand split of the tensor-core convolutions (storm_b200/csrc/conv_tc.cu): default = one fp16 pass + two fp8 correction
    # passes at twice the rate (2 passes); STORM_B200_TC_PREC=0 = bf16 hi/lo split (3 passes)
    passes = 3 if os.environ.get("STORM_B200_TC_PREC", "1") == "0" else 2
    split = "f16 x f16 + two e4m3 x e5m2 correction passes, fp32 accumulation" if passes == 2 else "bf16 hi/lo x 3, fp32 accumulation"
    pt = path_traffic()
    traffic, traffic_note = None, "no ncu capture committed for this build"
    if pt and name == "headline":
        traffic = pt.get("conv_dram_bytes_per_launch")
        traffic_note = pt.get("note")
    roofline = dict(bound="tensor", kernel=f"implicit-GEMM conv3x3/1x1 ({dom})", achieved=achieved, peak=peak, unit="TFLOP/s",
                    frac=achieved / peak if peak else None, traffic=traffic, traffic_note=traffic_note,
                    path_dram_bytes_per_utt_nfe=(pt or {}).get("path_dram_bytes_per_utt_nfe") if name == "headline" else None,
                    path_algorithmic_bytes_per_utt_nfe=W["gb"] * 1e9,
                    peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']})",
                    note=(f"algorithmic fp32-equivalent FLOPs (2*MACs) / CUDA-event time of the conv launches, measured in a SEPARATE profiled "
                          f"pass of one step (the timed `value` steps carry no instrumentation); the tcgen05 kernel spends {passes} "
                          f"16-bit-equivalent tensor passes per algorithmic MAC for fp32-grade results ({split}), so the tensor pipe runs at "
                          f"{passes}x this rate")
                    if dom == "tcgen05" else "fp32 SIMT FFMA kernel; tensor-core (tcgen05) path not used for this run",
                    issued_frac=(passes * achieved / peak) if (peak and dom == "tcgen05") else None,
                    launches=dp["launches"], avg_launch_ms=dp["ms"] / max(1, dp["launches"]), conv_share_of_step=conv_share,
                    by_engine=prof)
    if world == 1 and not args.no_cpu:      # the CPU baseline is a per-box figure: timed on rank 0 of the single-GPU run only
        threads = pick_cpu_threads()
        cpu_rate, cpu_dt, cpu_nfe = cpu_oracle_rate(threads, nfe_of(W), W)
        cpu_baseline = dict(value=cpu_rate, unit="utt/s", cores=threads, kind="port",
                            sample=f"oracle port, B=1 {W['F']}x{W['T']}, 1 PC step ({cpu_nfe} NFE) in {cpu_dt:.1f}s, scaled linearly to {nfe_of(W)} NFE")
    else:
        cpu_baseline = dict(value=None, unit="utt/s", cores=0, kind="port", sample="measured in the N=1 run only")
    cfg = workload_config(world, name, nfe_step)
    line = dict(metric="utterances/sec", value=value, unit="utt/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                dtype_detail=f"fp32 in/out; tensor-core products as {split} (fp32-grade; tolerance 1e-3 vs the reference)" if dom == "tcgen05" else "fp32 FFMA",
                data="synthetic", config=cfg, roofline=roofline, cpu_baseline=cpu_baseline,
                e2e=dict(value=e2e_value, unit="utt/s", h2d_bytes_per_step=world * B * L * 4, d2h_bytes_per_step=world * B * L * 4,
                         ms_per_step=e2e_ms / args.steps),
                gpu_launches=launches, clocks=clk,
                path_tflops=(world * B * nfe_step * args.steps * W["gflop"] / 1e3 / (ms * 1e-3)) if name != "sweep" else None)
    if sweep_rows is not None:
        sweep_rows = list(sweep_rows) + [rk45_row(model, y_dev, W)]
        line["sweep"] = sweep_rows
        line["config"]["workload"] = (f"{W['tag']}: sampler sweep, ncsnpp 27.8M, B={B}, 256x256: PC (reverse_diffusion+ald) and fixed-step "
                                      "probability flow at N in {10,20,30,50} (8 rows, 330 evaluations per utterance); `value` = utterances of all "
                                      "8 rows / time of one whole sweep; per-row rates in `sweep`, where the reference's scipy RK45 PF-ODE "
                                      "(one utterance, timed once) is the ninth row")
        line["value"] = world * B * 8 * args.steps / (ms * 1e-3)              # 8 fixed-step rows per sweep
        line["e2e"]["value"] = world * B * 8 * args.steps / (e2e_ms * 1e-3)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def rk45_row(model, y_dev, W):
    """The reference's PF-ODE sampler (scipy RK45 on the host, rtol = atol = 1e-5, sampling/__init__.py:119-129) on ONE utterance
    (model.py:224 forces one utterance per solve).  Every right-hand side is one native drift evaluation; the state crosses PCIe
    per evaluation as in the reference.  Timed once per bench invocation (wall clock, synchronised), outside the timed steps."""
    import torch
    from storm_b200 import _native
    sp = model.data_module.spec_config()
    eng = model.dnn.engine()
    yd = y_dev[:1]
    norm = _native.wave_norm(yd)
    T_pad = _native.padded_frames(_native.num_frames(yd.shape[1], sp))
    Y = _native.stft(yd, norm, sp, T_pad).unsqueeze(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sampler = model.get_ode_sampler(Y, N=30, minibatch=1, seed=42)
    sample, nfes = sampler()
    wav = eng.istft(sample.squeeze(1), norm, yd.shape[1], sp).cpu()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    return dict(sampler="pf_rk45_scipy", N=None, nfe=int(nfes[0]), ms=ms, utt_per_s=1.0 / (ms * 1e-3), batch=1,
                note="host-driven adaptive solver, one utterance per solve, timed once")


def sweep_once(model, y, ids, W, host=False):
    """BASELINE configs[4]: PC vs probability-flow ODE at N in {10, 20, 30, 50} on one batch; every row is timed on the device.
    PC = reverse_diffusion + ald (2N evaluations); PF = the same predictor on RSDE.discretize(probability_flow=True), no
    corrector, no noise (N evaluations).  The reference's RK45 row is measured separately (rk45_row)."""
    import torch
    from storm_b200 import _native
    rows = []
    dev = model._device()
    yd = y.to(dev, non_blocking=True) if host else y
    sp = model.data_module.spec_config()
    eng = model.dnn.engine()
    B = yd.shape[0]

    def timed(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        b.synchronize()
        return r, a.elapsed_time(b)

    for N in (10, 20, 30, 50):
        sc = model._sampler_cfg("reverse_diffusion", "ald", N, 1, W["snr"])
        (out, nfe), ms = timed(lambda: eng.enhance(yd, sc, sp, seed=42, utt_ids=ids))
        if host:
            out = out.cpu()
        rows.append(dict(sampler="pc", N=N, nfe=nfe, ms=ms, utt_per_s=B / (ms * 1e-3)))
    norm = _native.wave_norm(yd)
    T_pad = _native.padded_frames(_native.num_frames(yd.shape[1], sp))
    Y = _native.stft(yd, norm, sp, T_pad).unsqueeze(1)
    for N in (10, 20, 30, 50):
        s = model.sde
        sc = _native.make_sampler_config("reverse_diffusion", "none", N, 0, 0.0, model.t_eps, s.theta, s.sigma_min, s.sigma_max,
                                         probability_flow=True)
        (res, ms) = timed(lambda: eng.istft(eng.pc_sample(Y, sc, seed=42, utt_ids=ids)[0].squeeze(1), norm, yd.shape[1], sp))
        if host:
            res = res.cpu()
        rows.append(dict(sampler="pf_fixed_step", N=N, nfe=N, ms=ms, utt_per_s=B / (ms * 1e-3)))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(WORKLOADS), help="headline = BASELINE configs[1] (the driver's line)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (A/B runs of kernel variants)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
