#!/usr/bin/env python
"""bench.py — IQ Gsamples/s sensed (window + FFT + band energy + K-frame average + ANN) on N B200s.

Workload (BASELINE.json configs[1]): 1x B200 batched sensing, 3 channels + noise floor, 1024-pt FFT,
Hann window, 64-frame Welch average, ANN predict, 1e9 complex-float samples (8 GB) resident in HBM.
A "step" is one pass of the fused kernel over that batch (15258 decision groups, ONE kernel launch).
With --gpus N every rank owns its own 1e9-sample shard of the capture (weak scaling, groups are
independent, no collective on the data path).

Timing: W >= 3 warm-up steps; K steps timed with CUDA events on the launching stream, bracketed by
barrier + synchronize, max over ranks.  The input (8 GB) is 63x larger than L2, nothing is flushed.
  value     Gsamples/s, inputs already resident in HBM
  e2e       same metric through the C-ABI host path (crn_sense_batch_host): pinned HOST buffer,
            host->device copies and device->host result readback inside the timed region
  roofline  8 B/sample algorithmic bytes / measured kernel time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port (reference algorithm, options of this workload) on the host cores, on a
            bounded sample of the same IQ
`--impl reference` times the reference's CPU algorithm alone (host cores, bounded sample per step).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

NFFT, NAVG = 1024, 64
TOTAL_SAMPLES = 10 ** 9
METRIC = "IQ Gsamples/s sensed (FFT+band energy+ANN)"
UNIT = "Gsamples/s"
WORKLOAD = "configs[1]: 1xB200 batched sensing, 3 channels+NF, 1024-pt FFT, Hann, 64-frame Welch average + ANN predict, 1e9 complex-float samples"


def _oracle():
    """bench.py may execute oracle/ only for the cpu_baseline leg and --impl reference."""
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    import oracle
    return oracle


WORKLOADS = {
    # name: (description, builder(crn) -> (cfg, ngroups, streams))
    "config2": WORKLOAD,
    "wideband": "configs[2]: wideband sweep, 8192-pt FFT, 64 equal sub-channels, 100 MHz equivalent rate, energy-detection features, 64-frame average, 1e9 complex-float samples",
    "multiradio": "configs[3]: multi-radio, 4096 independent sensing streams (simulated CORNET nodes), 2048-pt FFT, 64-frame Welch average + ANN, one decision per stream (537e6 samples), streams sharded over the GPUs by stream id (strong scaling)",
    "multiradio_weak": "configs[3] shape, weak scaling: every GPU its own 4096 sensing streams, 2048-pt FFT, 64-frame Welch average + ANN",
    "sc16": "configs[1] fed in the USRP wire format: 1024-pt FFT, Hann, 64-frame Welch average + ANN, 1e9 samples as int16 (I,Q) pairs (4 B/sample), converted on the GPU",
    "refexact": "configs[0] on the GPU: reference-exact mode, 512-pt FFT, no window, |X| averaged over 10 frames, (sum)^2 features + ANN, 1e9 complex-float samples",
}
ACTIVE = {"name": "config2"}


def workload_config(crn, name=None):
    """Returns (cfg, decision groups per GPU).  The default - and the only bench line the contract asks for -
    is BASELINE configs[1]; the others are the remaining BASELINE configs, selectable with --workload.
    `crn` is whoever supplies the config structs: the product's ctypes mirror (our arm) or the oracle's own
    (reference arm, which must not load the product library); both offer the same four constructors."""
    w = name or ACTIVE["name"]
    if w == "wideband":
        cfg = crn.config_wideband(8192, 64, 64)
        return cfg, TOTAL_SAMPLES // cfg.group_samples
    if w in ("multiradio", "multiradio_weak"):
        cfg = crn.config_welch(2048, 64)
        return cfg, 4096
    if w == "refexact":
        cfg = crn.config_reference()
        return cfg, TOTAL_SAMPLES // cfg.group_samples
    cfg = crn.config_welch(NFFT, NAVG)
    if w == "sc16":
        cfg.iq_format = 1  # CRN_IQ_SC16
    return cfg, TOTAL_SAMPLES // cfg.group_samples  # 15258 full decisions, remainder dropped


def synth_cfg(crn, cfg):
    # Markov PU as documented (README.md:70-74), 64 decisions per dwell, SNR 10 dB, seed 12
    return crn.synth_config(cfg.group_samples, dwell_groups=64, snr_db=10.0, seed=12, hop_mode=0)


def base_config(extra=None):
    c = {"workload": WORKLOADS[ACTIVE["name"]], "nfft": NFFT, "navg": NAVG, "window": "hann(liquid, symmetric)",
         "detector": "|X|^2", "bands": "NF,CH1,CH2,CH3 (reference bin plan x2)", "ann": "4-5-3 logistic, fp64",
         "samples_per_gpu": None, "l2": "input 8 GB >> 126 MB L2, no flush needed",
         "synthetic": "OFDM PU (64 sc, cp16, taper4, 1.4->13 MS/s) hopping 833/835/838 MHz by the documented Markov matrix + AWGN, SNR 10 dB, seed 12"}
    if extra:
        c.update(extra)
    return c


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).
    NVML is polled every 2 ms from a thread (the timed region is tens of milliseconds); falls back to
    `nvidia-smi -lms` when the NVML binding is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, index):
        self.index = index
        self.rows = []      # (time, sm_mhz, power_w, reason_bits)
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        getr = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.rows.append((time.time(), float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)),
                                  n.nvmlDeviceGetPowerUsage(self.h) / 1e3, int(getr(self.h))))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 9:
                bits = 0
                for bit, v in zip((0x8, 0x40, 0x20, 0x4), r[5:9]):
                    if v.lower().startswith("active"):
                        bits |= bit
                try:
                    self.max_mhz = float(r[2])
                    self.rows.append((time.time(), float(r[1]), float(r[3]), bits))
                except ValueError:
                    pass

    def stop(self, t0, t1):
        if self.nvml is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        time.sleep(0.01)
        self.stop_flag = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        inside = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        sm = sorted(r[1] for r in inside)
        bits = 0
        for r in inside:
            bits |= r[3]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "samples": len(inside),
                "power_w_max": max(r[2] for r in inside),
                "reasons": sorted(name for bit, name in self.REASONS.items() if bits & bit),
                "source": "nvml 2 ms poll" if self.nvml else "nvidia-smi -lms 20"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy, of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def cpu_leg(crn, cfg, iq_host, budget_s, threads=None):
    """Oracle port on the host cores on a bounded sample of the same IQ: about `budget_s` seconds of
    all-core work (whole passes over the first groups of the capture).  Returns the cpu_baseline dict."""
    oracle = _oracle()
    nthreads = threads or oracle.port().crn_oracle_max_threads()
    gs = cfg.group_samples
    eps = 2 if cfg.iq_format == crn.IQ_SC16 else 1   # array elements per sample (int16 pairs vs complex64)
    have = iq_host.size // (gs * eps)
    probe = min(have, max(4 * nthreads, 32))
    t = oracle.time_port(cfg, iq_host[: probe * gs * eps], probe, nthreads)
    rate = probe * gs / max(t, 1e-9)
    n = int(min(have, max(probe, rate * budget_s // gs)))
    passes = max(1, int(round(budget_s / (n * gs / rate))))
    secs = sum(oracle.time_port(cfg, iq_host[: n * gs * eps], n, nthreads) for _ in range(passes))
    # the reference engine is single-threaded per radio: one thread on ~1 s of the same work
    n1 = int(max(1, min(have, rate / max(nthreads, 1) // gs)))
    t1 = oracle.time_port(cfg, iq_host[: n1 * gs * eps], n1, 1)
    return {"value": passes * n * gs / secs / 1e9, "unit": UNIT, "cores": nthreads, "kind": "port",
            "single_thread_value": n1 * gs / max(t1, 1e-9) / 1e9,
            "sample": "%d pass(es) over the first %d of %d decision groups (%d samples per pass) of the same synthetic "
                      "capture, %.1f s of CPU work on %d threads; reference algorithm restated in C (oracle/crn_oracle.c, "
                      "liquid-style radix-2 fp32 FFT; liquid-dsp/FFTW unavailable), one engine state per thread, gcc -O2"
                      % (passes, n, TOTAL_SAMPLES // gs, n * gs, secs, nthreads),
            "seconds": secs}


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores.  Nothing of the product is on this path:
    the workload comes from the oracle's own config statement (oracle/crn_oracle_config.c) and libcrnsense.so is
    never loaded into this process (checked below)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    oracle = _oracle()
    cfg, ngroups = workload_config(oracle)
    gs = cfg.group_samples
    nthreads = oracle.port().crn_oracle_max_threads()
    sc = synth_cfg(oracle, cfg)
    # bounded sample per step: ~budget seconds of all-core CPU work, whole run within a few minutes
    total_steps = args.steps + args.warmup
    budget = min(4.0, max(0.25, 150.0 / max(total_steps, 1)))
    probe_groups = max(nthreads, 8)
    iq, _ = oracle.synth(sc, probe_groups * gs)
    t = oracle.time_port(cfg, iq, probe_groups, nthreads)
    n = int(max(probe_groups, min(2048, (probe_groups * budget / max(t, 1e-9)))))
    if n > probe_groups:
        iq, _ = oracle.synth(sc, n * gs)
    # two builds of the same C statement: -O2 (what the tests check against) and -O3 -march=native built on this
    # host; the line reports the faster one
    builds = [("gcc -O2", None)]
    if oracle.native() is not None:
        builds.append(("gcc -O3 -march=native (built on this host)", oracle.native()))
    runs = {}
    for name, lib in builds:
        for _ in range(args.warmup):
            oracle.time_port(cfg, iq, n, nthreads, lib)
        runs[name] = sum(oracle.time_port(cfg, iq, n, nthreads, lib) for _ in range(args.steps))
    best = min(runs, key=runs.get)
    tot = runs[best]
    value = args.steps * n * gs / tot / 1e9
    sample = ("each step = first %d decision groups (%d samples) of the synthetic capture, all %d host threads, "
              "reference algorithm restated in C with this workload's options (the unmodified engine is fixed at "
              "N=512/K=10/no window and cannot express it); liquid-style radix-2 fp32 FFT (liquid-dsp absent; FFTW3f %s "
              "on this host); reported build: %s; all builds: %s"
              % (n, n * gs, nthreads, "installed but unused" if oracle.fftw_installed() else "not installed", best,
                 ", ".join("%s %.3f GS/s" % (k, args.steps * n * gs / v / 1e9) for k, v in runs.items())))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config({"samples_per_step": n * gs}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    # the unmodified reference engine in its own (only) mode, for context
    if oracle.ref() is not None:
        rcfg = oracle.config_reference()
        rsc = oracle.synth_config(rcfg.group_samples, dwell_groups=64, snr_db=10.0, seed=12)
        rn = nthreads * 2000
        riq, _ = oracle.synth(rsc, rn * rcfg.group_samples)
        sec, nd = oracle.time_ref(riq, 512, rn * 10, nthreads)
        line["reference_engine_native_mode"] = {
            "value": rn * rcfg.group_samples / sec / 1e9, "unit": UNIT, "cores": nthreads, "kind": "reference",
            "sample": "unmodified CE_Predictive_Node.cpp (oracle/_ref), N=512 K=10 rect |X|, %d decisions, one engine per thread" % nd}
    with open("/proc/self/maps") as fh:
        mapped = sorted({ln.split()[-1] for ln in fh if ln.rstrip().endswith(".so") and ROOT in ln})
    line["native_so_mapped"] = [os.path.relpath(m, ROOT) for m in mapped]
    assert not any("libcrnsense" in m for m in mapped), "the reference arm must not load the product library"
    print(json.dumps(line), flush=True)
    return 0


def kernel_traffic(kernel_name):
    """dram bytes per launch of this kernel from the committed ncu --set full captures (profiles/traffic.json is
    keyed by kernel name: a capture of one kernel says nothing about another), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get("kernels", {}).get(kernel_name, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class Ctx:
    """What every measurement of this process shares: ranks, device, peak."""
    pass


def measure(ctx, name, steps, warmup, env=None, keep=False, parity_groups=64):
    """Device-resident timing of ONE workload on this rank's GPU (+ parity spot check against the oracle, outside the
    timed region).  A step = one fused launch over the rank's batch + the device->host read of its results
    (features, MLP outputs, decisions, masks) into pinned memory; the timed region runs from the first kernel launch
    to the last feature readback, the readback of step i overlapping the kernel of step i+1 (two result
    sets, a copy stream).  Returns a dict; with keep=True the tensors stay alive in it for the e2e leg."""
    import numpy as np
    import torch
    crn, cdist, world, rank, local_rank, dev = ctx.crn, ctx.cdist, ctx.world, ctx.rank, ctx.local_rank, ctx.dev
    cfg, ngroups_all = workload_config(crn, name)
    strong = (name == "multiradio")
    first, ngroups = (cdist.shard_groups(ngroups_all, world, rank) if strong else (rank * ngroups_all, ngroups_all))
    gs = cfg.group_samples
    nsamp = ngroups * gs
    stream = torch.cuda.current_stream().cuda_stream
    saved = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        sensor = crn.Sensor(cfg, device=local_rank)   # CRN_* development switches are read once, at create
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    info = sensor.kernel_info()
    d_iq = torch.empty(nsamp, 2, dtype=torch.float32, device=dev)
    d_state = torch.empty(ngroups, dtype=torch.int32, device=dev)
    if name in ("multiradio", "multiradio_weak"):   # one decision group per stream; this rank's streams by stream id
        crn.synth_generate_streams(synth_cfg(crn, cfg), d_iq, first, ngroups, gs, d_state, local_rank, stream)
    else:                                            # rank r owns samples [r*nsamp, (r+1)*nsamp) of one long capture
        crn.synth_generate(synth_cfg(crn, cfg), d_iq, rank * nsamp, nsamp, d_state, local_rank, stream)
    sample_bytes = 8
    if cfg.iq_format == crn.IQ_SC16:   # quantise the capture to the 16-bit wire format, in place chunks
        sample_bytes = 4
        d16 = torch.empty(nsamp, 2, dtype=torch.int16, device=dev)
        step_q = 1 << 26
        for i0 in range(0, nsamp, step_q):
            d16[i0:i0 + step_q] = (d_iq[i0:i0 + step_q] * 32768.0).round_().clamp_(-32768, 32767).to(torch.int16)
        d_iq = d16
        torch.cuda.synchronize()
    # two result sets: the results of step i are read back (copy stream) while the kernel of step i+1 runs
    def result_set():
        return (torch.empty(ngroups, cfg.nbands, dtype=torch.float32, device=dev),
                torch.empty(ngroups, 3, dtype=torch.float64, device=dev),
                torch.empty(ngroups, dtype=torch.int32, device=dev),
                torch.empty(ngroups, dtype=torch.int64, device=dev))
    d_sets = [result_set(), result_set()]
    h_sets = [[torch.empty_like(t, device="cpu").pin_memory() for t in d] for d in d_sets]
    d2h_bytes = sum(t.numel() * t.element_size() for t in h_sets[0])
    main_stream = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    copied = [None, None]   # event: this set's previous readback has finished

    def step(i, ev_kernel):
        """fused launch of step i into result set i % 2, then its readback queued on the copy stream"""
        b = i & 1
        if copied[b] is not None:
            main_stream.wait_event(copied[b])      # the set is free again (it was, long ago)
        sensor.sense_device(d_iq, ngroups, d_sets[b][0], d_sets[b][1], d_sets[b][2], d_sets[b][3], stream)
        ev_kernel.record(main_stream)
        copy_stream.wait_event(ev_kernel)
        with torch.cuda.stream(copy_stream):
            for h, d in zip(h_sets[b], d_sets[b]):
                h.copy_(d, non_blocking=True)
            copied[b] = torch.cuda.Event()
            copied[b].record(copy_stream)

    nwarm = max(warmup, 3)
    for i in range(nwarm):
        step(i, torch.cuda.Event())
    ctx.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25 if keep else 0.05)
    launches0 = sensor.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 2)]
    ctx.barrier()
    t_wall0 = time.time()
    ev[0].record(main_stream)
    for i in range(steps):
        step(nwarm + i, ev[i + 1])              # ev[i+1]: kernel i done; kernels run back to back on the main stream
    main_stream.wait_event(copied[(nwarm + steps - 1) & 1])
    main_stream.wait_event(copied[(nwarm + steps) & 1])
    ev[steps + 1].record(main_stream)           # every result of every step is in pinned host memory
    ctx.barrier()
    t_wall1 = time.time()
    gpu_launches = sensor.launches - launches0
    kern_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    total_ms = ev[0].elapsed_time(ev[steps + 1])
    last = (nwarm + steps - 1) & 1
    d_feat, d_ann, d_dec, d_mask = d_sets[last]
    h_res = h_sets[last]
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    total_ms_max = cdist.max_over_ranks(total_ms, dev)
    total_samples = (ngroups_all * gs) if strong else world * nsamp
    value = total_samples * steps / (total_ms_max * 1e-3) / 1e9

    # ---- parity spot check of this very batch against the oracle (outside every timed region) ----------
    pick = sorted(set(int(round(i * (ngroups - 1) / max(parity_groups - 1, 1))) for i in range(min(parity_groups, ngroups))))
    idx = torch.tensor(pick, device=dev)
    rows = d_iq.view(ngroups, gs, 2)[idx].cpu().numpy()
    iq_pick = rows.ravel() if cfg.iq_format == crn.IQ_SC16 else rows.view(np.complex64).ravel()
    oracle = _oracle()
    of, oa, od, om = oracle.sense_port(cfg, iq_pick, nthreads=oracle.port().crn_oracle_max_threads())
    gf, ga, gd = h_res[0].numpy()[pick], h_res[1].numpy()[pick], h_res[2].numpy()[pick]
    gm = h_res[3].numpy()[pick].astype(np.uint64)
    parity = {"groups_checked": len(pick), "of_groups": ngroups,
              "feat_max_rel": float((abs(gf - of) / abs(of)).max()),
              "ann_max_abs": float(abs(ga - oa).max()), "decisions_equal": bool((gd == od).all()),
              "masks_equal": bool((gm == om).all()),
              "checker": "oracle/crn_oracle.c on the same IQ, groups spread evenly over the rank's batch; no tolerance floor"}
    peak, peak_src = ctx.peak
    kernel_ms = sum(kern_ms) / len(kern_ms)
    achieved = nsamp * sample_bytes / (kernel_ms * 1e-3) / 1e9
    out = {"name": name, "workload": WORKLOADS[name], "value": value, "ms_per_step": total_ms_max / steps,
           "scaling": "strong" if strong else "weak", "steps": steps, "gpu_launches": int(gpu_launches),
           "kernel": info["name"], "kernel_ms": kernel_ms, "kernel_ms_minmax": [min(kern_ms), max(kern_ms)],
           "achieved_gbs": achieved, "frac": achieved / peak, "samples_per_gpu": nsamp, "groups_per_gpu": ngroups,
           "total_samples": total_samples, "sample_bytes": sample_bytes, "d2h_bytes_per_step": d2h_bytes,
           "parity_check": parity, "clocks": clocks, "kernel_info": info, "nfft": cfg.nfft, "navg": cfg.navg,
           "nbands": cfg.nbands, "peak": peak, "peak_source": peak_src}
    if env:
        out["env"] = env
    if keep:
        out["_keep"] = dict(cfg=cfg, sensor=sensor, d_iq=d_iq, d_feat=d_feat, d_ann=d_ann, d_dec=d_dec, d_mask=d_mask,
                            ngroups=ngroups, nsamp=nsamp, gs=gs, stream=stream)
    else:
        sensor.close()
        del d_iq, d_feat, d_ann, d_dec, d_mask, d_state, d_sets
        torch.cuda.empty_cache()
    return out


# the other BASELINE configs, measured after the headline workload (5 steps each, own captures, outside its timed
# region) so that the driver-run line carries them too: name -> (workload, environment switches read at crn_create)
OTHER_WORKLOADS = {
    "configs[0] reference-exact on the GPU": ("refexact", None),
    "configs[1] with every bin live (no band-plan pruning)": ("config2", {"CRN_NO_PRUNE": "1"}),
    "configs[2] wideband 8192-pt x 64 sub-channels": ("wideband", None),
    "configs[3] multi-radio 4096 streams x 2048-pt, sharded by stream id": ("multiradio", None),
}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import crn_b200 as crn
    import importlib
    cdist = importlib.import_module("crn_b200.dist")

    ctx = Ctx()
    ctx.crn, ctx.cdist = crn, cdist
    world = ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = ctx.rank = int(os.environ.get("RANK", "0"))
    local_rank = ctx.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libcrnsense has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = ctx.dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ctx.barrier = barrier
    ctx.peak = measured_peak()

    main = measure(ctx, ACTIVE["name"], args.steps, args.warmup, keep=True)
    k = main.pop("_keep")
    cfg, sensor, d_iq, d_feat, d_dec, d_mask = k["cfg"], k["sensor"], k["d_iq"], k["d_feat"], k["d_dec"], k["d_mask"]
    ngroups, nsamp, stream = k["ngroups"], k["nsamp"], k["stream"]
    sample_bytes = main["sample_bytes"]

    # optional occupancy exchange (north_star): outside the data path and the timed region
    counts = [cdist.shard_groups(4096, world, r)[1] for r in range(world)] if main["scaling"] == "strong" else None
    all_dec = cdist.gather_occupancy(d_dec, counts)  # every rank sees the whole capture's decisions
    dec_hist = cdist.occupancy_histogram(d_dec)
    assert all_dec.numel() == (sum(counts) if counts else world * ngroups)
    # cooperative fusion of the ranks' occupancy masks (8 B per decision over NCCL + one fusion kernel), also outside
    if counts is None or len(set(counts)) == 1:
        fused = cdist.fuse_across_ranks(d_mask, max(cfg.nbands, 3), crn.FUSE_OR, stream)
        torch.cuda.synchronize()
        assert fused.numel() == ngroups and (world > 1 or torch.equal(fused, d_mask))

    # ---- end to end through the C-ABI host path: pinned host IQ -> H2D -> kernel -> D2H results --------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_iq = torch.empty(nsamp, 2, dtype=d_iq.dtype, pin_memory=True)
    h_iq.copy_(d_iq)
    torch.cuda.synchronize()
    res = (crn.Result * ngroups)()
    ptr = C.c_void_p(h_iq.data_ptr())
    sensor.sense_host_raw(ptr, ngroups, res)  # warm-up: allocates the staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sensor.sense_host_raw(ptr, ngroups, res)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_sec = cdist.max_over_ranks(t1 - t0, dev)
    hf, ha, hd, _ = crn.results_to_arrays(res, cfg.nbands)
    # the host path launches 128-group chunks whose groups are split over several CTAs: equal to fp32 rounding
    e2e_ok = bool(np.allclose(hf, d_feat.cpu().numpy(), rtol=2e-6, atol=0) and np.array_equal(hd, d_dec.cpu().numpy()))
    e2e = {"value": main["total_samples"] * e2e_steps / e2e_sec / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": nsamp * sample_bytes, "d2h_bytes_per_step": ngroups * (cfg.nbands * 4 + 3 * 8 + 4 + 8),
           "steps": e2e_steps, "path": "crn_sense_batch_host (C-ABI), pinned host IQ, 64 MiB double-buffered chunks; wall clock, max over ranks",
           "matches_device_path": e2e_ok}

    # ---- CPU baseline on rank 0 (N = 1 only): bounded sample of the same capture -----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        iq_host = h_iq.numpy().ravel() if cfg.iq_format == crn.IQ_SC16 else h_iq.numpy().view(np.complex64).ravel()
        cpu = cpu_leg(crn, cfg, iq_host, args.cpu_seconds)
    sensor.close()
    del h_iq, d_iq, k
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs on the same box, same process (every rank takes part) ---------------
    others = {}
    if ACTIVE["name"] == "config2" and not args.no_others:
        for label, (wname, env) in OTHER_WORKLOADS.items():
            r = measure(ctx, wname, args.other_steps, 3, env=env)
            others[label] = {"gsamples_s": r["value"], "scaling": r["scaling"], "kernel": r["kernel"], "kernel_ms": r["kernel_ms"],
                             "ms_per_step": r["ms_per_step"], "steps": r["steps"], "frac": r["frac"],
                             "achieved_gbs": r["achieved_gbs"], "samples_per_gpu": r["samples_per_gpu"],
                             "total_samples": r["total_samples"], "traffic": kernel_traffic(r["kernel"]),
                             "parity_check": r["parity_check"], "clocks": r["clocks"], "workload": r["workload"]}
            if env:
                others[label]["env"] = env

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = ctx.peak
    roofline = {"bound": "hbm", "achieved": main["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": main["frac"],
                "traffic": kernel_traffic(main["kernel"]), "kernel": main["kernel"], "peak_source": peak_src,
                "algorithmic_bytes_per_launch": nsamp * sample_bytes, "kernel_ms": main["kernel_ms"],
                "note": "%d B per complex sample read once; kernel_ms = CUDA events around the launch alone; the feature "
                        "write-back (%d B per launch) is not counted in the bytes; `value` additionally includes the "
                        "device->host read of the results" % (sample_bytes, main["d2h_bytes_per_step"])}
    line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": main["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config({"samples_per_gpu": nsamp, "groups_per_gpu": ngroups, "kernel": main["kernel_info"],
                                   "nfft": cfg.nfft, "navg": cfg.navg, "nbands": cfg.nbands,
                                   "timed_region": "per step: one fused launch + device->host copy of features, MLP outputs, decisions and masks"}),
            "clocks": main["clocks"], "e2e": e2e, "gpu_launches": main["gpu_launches"], "roofline": roofline,
            "cpu_baseline": cpu, "parity_check": main["parity_check"], "decision_histogram": dec_hist.cpu().tolist(),
            "ms_per_step_minmax": main["kernel_ms_minmax"], "other_workloads": others}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_sweep(args):
    """BASELINE configs[4]: FFT size 256..8192 x batch 1e3..1e6 frames (Welch K=64 + ANN; N=256 uses the 16-channel
    energy plan), device-resident, every rank its own batch (weak), max over ranks; the reference CPU path (oracle
    port, all host threads, bounded sample) timed beside every FFT size on rank 0.  ONE JSON line with all rows."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import crn_b200 as crn
    import importlib
    cdist = importlib.import_module("crn_b200.dist")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --sweep: no CUDA device; libcrnsense has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream().cuda_stream
    peak, peak_src = measured_peak()
    max_samples = 2 * 10 ** 9                       # 16 GB of IQ per GPU at most (8192 x 1e6 is capped)
    d_iq = torch.empty(max_samples, 2, dtype=torch.float32, device=dev)
    crn.synth_generate(crn.synth_config(65536, dwell_groups=64, snr_db=10.0, seed=12), d_iq, rank * max_samples,
                       max_samples, None, local_rank, stream)
    torch.cuda.synchronize()
    rows = []
    for n in (256, 512, 1024, 2048, 4096, 8192):
        cfg = crn.config_welch(n, NAVG) if n >= 512 else crn.config_wideband(n, NAVG, 16)
        gs = cfg.group_samples
        cpu = None
        if rank == 0 and not args.no_cpu:          # reference CPU path beside this FFT size (bounded sample)
            take = min(max_samples, 64 * 10 ** 6)
            iq_host = d_iq[:take].cpu().numpy().view(np.complex64).ravel()
            cpu = cpu_leg(crn, cfg, iq_host, min(args.cpu_seconds, 2.0))
        with crn.Sensor(cfg, device=local_rank) as sensor:
            info = sensor.kernel_info()
            for frames in (10 ** 3, 10 ** 4, 10 ** 5, 10 ** 6):
                ng = max(1, min(frames // cfg.navg, max_samples // gs))
                d_feat = torch.empty(ng, cfg.nbands, dtype=torch.float32, device=dev)
                d_ann = torch.empty(ng, 3, dtype=torch.float64, device=dev)
                d_dec = torch.empty(ng, dtype=torch.int32, device=dev)
                for _ in range(max(args.warmup, 3)):
                    sensor.sense_device(d_iq, ng, d_feat, d_ann, d_dec, None, stream)
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    sensor.sense_device(d_iq, ng, d_feat, d_ann, d_dec, None, stream)
                e1.record()
                torch.cuda.synchronize()
                ms = cdist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
                gsps = world * ng * gs / ms / 1e6
                rows.append({"nfft": n, "frames_per_gpu": ng * cfg.navg, "capped": ng * cfg.navg < frames,
                             "samples_per_gpu": ng * gs, "ms_per_launch": ms, "gsamples_s": gsps,
                             "frac_of_hbm_peak": gsps * 8 / (peak * world), "kernel": info["name"],
                             "cpu_gsamples_s": cpu["value"] if cpu else None, "cpu_cores": cpu["cores"] if cpu else None})
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "scaling": "weak",
                          "data": "synthetic", "dtype": "f32", "hbm_peak_gbs": peak, "peak_source": peak_src,
                          "config": {"workload": "configs[4]: FFT-size sweep 256-8192 points x batch 1e3-1e6 frames, "
                                                 "64-frame Welch average + ANN, beside the reference CPU path on host cores"},
                          "sweep": rows}), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configs appended to the default line")
    ap.add_argument("--other-steps", type=int, default=5)
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[4]: FFT size x batch frames, one JSON line with all rows")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS),
                    help="default = BASELINE configs[1] (the contract's bench line); others = remaining BASELINE configs")
    args = ap.parse_args()
    ACTIVE["name"] = args.workload
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 behind
    # Python's back (NCCL prints its version banner there under NCCL_DEBUG=VERSION) are sent to stderr.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        return run_reference(args)
    if args.sweep:
        return run_sweep(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
