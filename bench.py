#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metric (input Mbp/s of the BWT + sampled SA/ISA build).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

A step = one pass of the hot path (K1 decode -> K2 sort -> K3 extract -> [K5 gap, K6 merge] ->
K4 dictionary -> K7 SA/ISA walk) over one synthetic input.  `value`: inputs already resident in
HBM; `e2e`: the same through the C ABI with pinned HOST buffers (H2D of the input file bytes,
D2H of BWT + anchors + SA + ISA inside the timed region).  The CPU oracle is executed only for
the `cpu_baseline` leg and by `--impl reference`.  The metric's second half, LF-steps/s, is
`lf_steps_per_s` (K7's dictionary, 2^20 chains x 256 dependent steps on the GPU) next to
`cpu_baseline.lf_steps_per_s` (the reference's bwttestdecodespeed instrument restated: 8 interleaved
chains on one host thread over the BWT of the CPU sample; measured outside the timed build).
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
sys.path.insert(0, ROOT)

METRIC = "input Mbp/s BWT+SSA build"
UNIT = "Mbp/s"


def build_params(workload):
    # cfg1 is the reference's CPU-runnable BWT-only case; the others build the full sampled SA/ISA
    if workload == "cfg1":
        return dict(sasamplingrate=32, isasamplingrate=262144, bwtonly=True)
    return dict(sasamplingrate=32, isasamplingrate=262144, bwtonly=False)


def config_dict(args, n, itype, extra=None):
    from bwtb3m_b200.workloads import CONFIGS
    d = {
        "workload": "%s: %s" % (args.workload, CONFIGS[args.workload][3]),
        "n_symbols": int(n),
        "inputtype": itype,
        "sasamplingrate": 32,
        "isasamplingrate": 262144,
        "bwtonly": 1 if args.workload == "cfg1" else 0,
        "scale": args.scale,
        "l2": "flushed between timed steps (256 MiB write); cfg3/cfg5 inputs are also far larger than L2",
    }
    if extra:
        d.update(extra)
    return d


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append(line.strip())
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]  # the upper half of the samples = under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_oracle_run(itype, filebytes, nsyms_limit, params, threads, want_lf=False):
    """Times the CPU oracle (restated reference algorithm) on a prefix of the workload."""
    from oracle import oracle as orc
    orc.build()
    if itype in ("pac", "pacterm"):
        from bwtb3m_b200.workloads import pac_file_from_packed
        pac = pac_file_from_packed(filebytes, nsyms_limit)  # prefix of the workload as its own .pac file
        t0 = time.perf_counter()
        t = orc.decode_pac(pac.tobytes(), term=(itype == "pacterm"))
    else:
        t0 = time.perf_counter()
        t = np.ascontiguousarray(filebytes[:nsyms_limit])
    n = t.size
    nblocks = orc.default_numblocks(n, 2 << 30, threads)  # reference defaults: mem=2 GiB
    rate = 64 if params["bwtonly"] else 4096  # anchors every 4096 positions keep all host threads busy in the SSA walk
    bwt, pp, st = orc.b3m(t, nblocks=nblocks, rate=rate, nthreads=threads)
    if not params["bwtonly"]:
        orc.ssa(bwt, pp, params["sasamplingrate"], params["isasamplingrate"], nthreads=threads)
    dt = time.perf_counter() - t0
    if want_lf:
        # LF-steps/s on the host, outside the timed build: the reference's instrument (bwttestdecodespeed.cpp:67-97),
        # 8 interleaved dependent chains on one thread over the BWT of the sample, started at anchor ranks
        lf = orc.lf_speed(bwt, pp[:, 0], tpar=8, maxsteps=1 << 21)
        return dt, n, nblocks, lf
    return dt, n, nblocks


def file_level_leg(args, itype, data, nsym, params):
    """The reference's own contract (/root/reference/src/bwtb3m.cpp:62-65): input FILE in, output FILES out, through
    b3m_compute_bwt (K8 run-length Huffman encoding on the device, then the writes) -- and, for pacterm, the BWA
    export b3m_to_bwa on top (/root/reference/src/bwtb3mtobwa.cpp:29).  Files live on tmpfs (/dev/shm) when it has
    room, so the figure is host memory + PCIe + device, not a disk."""
    import shutil
    import tempfile
    from bwtb3m_b200 import files
    need = 6 * data.size + 16 * nsym // 32 + (1 << 30)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need else tempfile.gettempdir()
    if shutil.disk_usage(base).free < need:
        return {"skipped": "no room for the files (%d bytes needed under %s)" % (need, base)}
    d = tempfile.mkdtemp(prefix="b3m_bench_", dir=base)
    try:
        fn = os.path.join(d, "in.dat")
        data.tofile(fn)
        out = os.path.join(d, "out.bwt")
        runs = []
        for _ in range(2):  # the first run also pays the CUDA context and the first pinned allocation
            t0 = time.perf_counter()
            r = files.compute_bwt(fn, inputtype=itype, outputfilename=out, sasamplingrate=params["sasamplingrate"],
                                  isasamplingrate=params["isasamplingrate"], bwtonly=params["bwtonly"], tmpprefix=os.path.join(d, "tmp"))
            runs.append((time.perf_counter() - t0, r["seconds_total"], r["seconds_device"]))
        wall, tot, dev = min(runs)
        sizes = {suf: os.path.getsize(out[:-4] + suf) for suf in (".bwt", ".hist", ".sa", ".isa", ".preisa") if os.path.exists(out[:-4] + suf)}
        res = {"value": nsym / wall / 1e6, "unit": UNIT, "seconds": wall, "seconds_device": dev, "dir": base,
               "call": "b3m_compute_bwt: %s file -> %s" % (itype, " + ".join(sorted(sizes))),
               "bytes_in": int(data.size), "bytes_out": int(sum(sizes.values())), "first_run_seconds": runs[0][0]}
        if itype == "pacterm" and not params["bwtonly"]:
            t0 = time.perf_counter()
            files.to_bwa(out, os.path.join(d, "bwa.bwt"), os.path.join(d, "bwa.sa"))
            tb = time.perf_counter() - t0
            res["to_bwa_seconds"] = tb
            res["value_with_bwa_export"] = nsym / (wall + tb) / 1e6
        return res
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  gt1/bwtb3m cannot be
    built here (libmaus2 is absent), so this arm times the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bwtb3m_b200 import workloads
    from oracle import oracle as orc
    itype, data, nsym = workloads.make(args.workload, args.scale)
    params = build_params(args.workload)
    threads = os.cpu_count() or 1
    sample = min(nsym, args.ref_sample)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_oracle_run(itype, data, sample, params, threads)
    tot, syms = 0.0, 0
    for _ in range(args.steps):
        dt, n, nblocks = cpu_oracle_run(itype, data, sample, params, threads)
        tot += dt
        syms += sample
    v = syms / tot / 1e6
    cfg = config_dict(args, sample, itype)
    if sample < nsym:
        # the CPU arm cannot run the full workload inside the bench's time budget: its line names what it processed
        cfg["workload"] = "first %d symbols of %s" % (sample, cfg["workload"])
        cfg["full_workload_n_symbols"] = int(nsym)
        cfg["note"] = ("bounded sample: the restated reference's throughput FALLS with n (see cpu_baseline.curve of the GPU arm), "
                       "so a ratio against the full-size GPU line understates the CPU time of the full workload")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "first %d symbols of the workload per step, %d blocks; restated reference (libmaus2 unavailable)" % (sample, nblocks)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bwtb3m_b200 import Engine, multigpu, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    itype, data, nsym = workloads.make(args.workload, args.scale)
    params = build_params(args.workload)
    host_in = torch.from_numpy(data).pin_memory()
    dev_in = host_in.cuda(non_blocking=False)
    stream = torch.cuda.Stream()
    eng = Engine(local, stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    state = {"drv": None, "strategy": "single"}

    def build(host_sa_ptr=0, host_bwa_ptr=0):
        if world == 1:
            # e2e leg: the sampled SA and BWA's packed BWT land in the caller's pinned buffers while the last sorting step still runs
            eng.build(numblocks=args.numblocks, host_sa_ptr=host_sa_ptr, host_bwa_ptr=host_bwa_ptr, gapmode=args.gapmode, **params)
        else:
            state["drv"], res = multigpu.build_distributed(eng, local_blocks=args.numblocks, driver=state["drv"], strategy=args.strategy, **params)
            state["strategy"] = res["strategy"]

    def step_device():
        eng.load_device(dev_in.data_ptr(), dev_in.numel(), itype)
        build()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step_device()
        info = eng.info()
        # ---- value: device-resident input, CUDA events on the engine's stream, max over ranks ----
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.15)
        sync_all()
        l0 = eng.info()["launches"]
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            ev[k][0].record(stream)
            step_device()
            ev[k][1].record(stream)
        sync_all()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        l1 = eng.info()["launches"]
        time.sleep(0.1)
        sampler.stop()
        info = eng.info()
        launches = int(l1 - l0)
        if world > 1:
            red = torch.tensor([total_ms, float(launches)], dtype=torch.float64, device="cuda")
            mx = red.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(red, op=dist.ReduceOp.SUM)
            total_ms, launches = float(mx[0].item()), int(red[1].item())

        # ---- e2e: pinned host input -> results in pinned host buffers (rank 0), wall clock ----
        n = info["n"]
        out = None
        # pacterm builds deliver what the reference pipeline ends in (bwtb3m -> bwtb3mtobwa): BWA's packed
        # BWT words + sampled SA (+ ISA, anchors); other inputs deliver the BWT as one byte per symbol
        as_bwa = itype == "pacterm" and not params["bwtonly"]
        nwords = (n - 1 + 15) >> 4
        if rank == 0:
            out = {
                "bwt": (torch.empty(nwords, dtype=torch.int32) if as_bwa else torch.empty(n, dtype=torch.uint8)).pin_memory(),
                "preisa": torch.empty(2 * info["npreisa"], dtype=torch.int64).pin_memory(),
                "sa": torch.empty(max(info["nsa"], 1), dtype=torch.int64).pin_memory(),
                "isa": torch.empty(max(info["nisa"], 1), dtype=torch.int64).pin_memory(),
            }

        # N > 1: the input file is uploaded over every rank's PCIe link (1/N each, all-gather over NVLink) and the two
        # large results leave the same way (multigpu.load_distributed / fetch_distributed); the host buffers of the
        # results are ONE page-locked shared-memory region mapped by all ranks
        io_state = {}
        shared = None
        if world > 1 and as_bwa:
            try:
                shared = {"bwt": multigpu.SharedHost(4 * nwords, "bwa", rank, world, width=4), "sa": multigpu.SharedHost(8 * max(info["nsa"], 1), "sa", rank, world, width=8)}
            except RuntimeError as ex:  # raised on every rank alike (no room under /dev/shm): results leave through rank 0
                shared = None
                if rank == 0:
                    print("bench: %s; the e2e leg fetches through rank 0" % ex, file=sys.stderr)

        def step_e2e():
            if world == 1:
                eng.load_host_ptr(host_in.data_ptr(), host_in.numel(), itype)
            else:
                multigpu.load_distributed(eng, host_in, itype, io_state)
            build(out["sa"].data_ptr() if (rank == 0 and info["nsa"] and world == 1) else 0, out["bwt"].data_ptr() if (rank == 0 and as_bwa and world == 1) else 0)
            if shared is not None and state["strategy"] == "shard":
                multigpu.fetch_distributed(eng, state["drv"], shared["bwt"].ptr(), shared["sa"].ptr() if info["nsa"] else 0,
                                           out["preisa"].data_ptr() if rank == 0 else 0, out["isa"].data_ptr() if (rank == 0 and info["nisa"]) else 0)
            elif rank == 0:
                if as_bwa:
                    eng.fetch_bwa(out_ptr=out["bwt"].data_ptr())
                eng.fetch_ptrs(0 if as_bwa else out["bwt"].data_ptr(), out["preisa"].data_ptr(),
                               out["sa"].data_ptr() if info["nsa"] else 0, out["isa"].data_ptr() if info["nisa"] else 0)

        if shared is not None and info["nsa"] and isinstance(state["drv"], dict):
            state["drv"]["stream_sa_host"] = shared["sa"].ptr()  # the SA samples leave during every rank's finish kernel
        step_e2e()
        sync_all()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            step_e2e()
        sync_all()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            mx = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            e2e_s = float(mx[0].item())
        h2d = int(host_in.numel())  # N > 1: the sum over the ranks' slices
        # where the end-to-end time goes: one more step with a device synchronisation and a barrier behind every phase (untimed above)
        e2e_phases = None
        if world > 1:
            def phase(fn):
                sync_all()
                t = time.perf_counter()
                fn()
                sync_all()
                return 1e3 * (time.perf_counter() - t)
            ph = [phase(lambda: multigpu.load_distributed(eng, host_in, itype, io_state)), phase(lambda: build())]
            if shared is not None and state["strategy"] == "shard":
                ph.append(phase(lambda: multigpu.fetch_distributed(eng, state["drv"], shared["bwt"].ptr(), shared["sa"].ptr() if info["nsa"] else 0,
                                                                   out["preisa"].data_ptr() if rank == 0 else 0, out["isa"].data_ptr() if (rank == 0 and info["nisa"]) else 0)))
            mx = torch.tensor(ph, dtype=torch.float64, device="cuda")
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            e2e_phases = dict(zip(("load", "build", "fetch"), [round(float(x), 3) for x in mx.tolist()]))
            if len(ph) == 3:
                # inside the fetch: CUDA events on every rank's stream
                state["drv"]["fetch_timeline"] = []
                ph.append(phase(lambda: multigpu.fetch_distributed(eng, state["drv"], shared["bwt"].ptr(), shared["sa"].ptr() if info["nsa"] else 0,
                                                                   out["preisa"].data_ptr() if rank == 0 else 0, out["isa"].data_ptr() if (rank == 0 and info["nisa"]) else 0)))
                tl = state["drv"].pop("fetch_timeline")
                mine = {tl[i][0]: tl[i - 1][1].elapsed_time(tl[i][1]) for i in range(1, len(tl))}
                alltl = [None] * world
                dist.all_gather_object(alltl, mine)
                e2e_phases["fetch_timeline"] = {k: {"rank0": round(alltl[0].get(k, 0.0), 3), "max": round(max(d.get(k, 0.0) for d in alltl), 3)} for k in mine}
        e2e_check = None
        if shared is not None and state["strategy"] == "shard":
            # what landed in the shared host buffers equals what rank 0's engine returns through the plain fetch calls
            sync_all()
            if rank == 0:
                w_ref, _, _, _ = eng.fetch_bwa()
                e2e_check = bool(np.array_equal(shared["bwt"].t.numpy().view(np.uint32)[:nwords], w_ref))
                del w_ref
                if info["nsa"]:
                    sa_ref = eng.fetch(bwt=False, preisa=False, isa=False)["sa"]
                    e2e_check = e2e_check and bool(np.array_equal(shared["sa"].t.numpy().view(np.uint64)[:info["nsa"]], sa_ref))
                    del sa_ref
        d2h = int((4 * nwords if as_bwa else n) + 16 * info["npreisa"] + 8 * info["nsa"] + 8 * info["nisa"])

        if isinstance(state["drv"], dict):
            state["drv"].pop("stream_sa_host", None)  # the device-timed steps below deliver nothing to the host
        # ---- roofline of the dominant kernel: per-kernel CUDA events on separate profiled steps ----
        eng.set_profile(True)
        nprof = 2
        for k in range(nprof):
            flush.fill_(k)
            step_device()
        sync_all()
        kt = eng.kernel_times()
        eng.set_profile(False)
        # N > 1: the slowest rank per kernel, and the phases of one position-sharded build on every rank (CUDA events)
        kt_max, build_timeline = None, None
        if world > 1:
            allkt = [None] * world
            dist.all_gather_object(allkt, {k: v["ms"] / nprof for k, v in kt.items()})
            kt_max = {k: round(max(d.get(k, 0.0) for d in allkt), 4) for k in sorted(set().union(*allkt))}
            direct = state["drv"].get("direct") if isinstance(state["drv"], dict) else None
            if direct is not None and state["strategy"] == "shard":
                direct.timeline = []
                sync_all()
                step_device()
                sync_all()
                tl = direct.timeline
                direct.timeline = None
                mine = {tl[i][0]: tl[i - 1][1].elapsed_time(tl[i][1]) for i in range(1, len(tl))}
                alltl = [None] * world
                dist.all_gather_object(alltl, mine)
                if mine:
                    build_timeline = {k: {"rank0": round(alltl[0].get(k, 0.0), 3), "max": round(max(d.get(k, 0.0) for d in alltl), 3),
                                          "min": round(min(d.get(k, 0.0) for d in alltl), 3)} for k in mine}
        lf_ms = None
        k8 = None
        if rank == 0:
            lf_ms, _ = eng.lf_bench(1 << 20, 256)
            if world == 1 and not args.no_file_level:
                # K8 alone: run-length Huffman encoding of the BWT on the device, payload written to tmpfs
                import tempfile
                tdir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
                tfn = os.path.join(tdir, "b3m_bench_k8_%d.bwt" % os.getpid())
                try:
                    eng.set_profile(True)
                    t0 = time.perf_counter()
                    eng.write_bwt(tfn)
                    tw = time.perf_counter() - t0
                    k8t = eng.kernel_times()
                    eng.set_profile(False)
                    kms = sum(v["ms"] for k, v in k8t.items() if k.startswith("rl_"))
                    kb = sum(v["bytes"] for k, v in k8t.items() if k.startswith("rl_"))
                    k8 = {"kernels_ms": {k: round(v["ms"], 4) for k, v in k8t.items()}, "rl_kernels_ms": kms, "algorithmic_bytes": kb,
                          "achieved_gbs": kb / (kms * 1e-3) / 1e9 if kms else None, "write_bwt_seconds": tw, "file_bytes": os.path.getsize(tfn)}
                finally:
                    if os.path.exists(tfn):
                        os.remove(tfn)
    # ---- N > 1: what was timed equals a single-GPU build of the same input, bit for bit ----
    parity = None
    if world > 1:
        with torch.cuda.stream(stream):
            step_device()
            sync_all()
            if rank == 0:
                parity = {"against": "single-GPU build on rank 0, same input", "strategy": state["strategy"]}
                ref = None
                try:
                    multi = eng.fetch()
                    # room for a second engine on rank 0's device: the bench's own scratch goes first
                    flush = None
                    io_state.clear()
                    torch.cuda.empty_cache()
                    ref = Engine(local)
                    ref.load_device(dev_in.data_ptr(), dev_in.numel(), itype)
                    ref.build(numblocks=1, preisarate=info["preisarate"], **params)
                    one = ref.fetch()
                    for k in sorted(one):
                        parity[k] = bool(np.array_equal(multi[k], one[k]))
                    parity["ok"] = all(v for k, v in parity.items() if k not in ("against", "strategy"))
                    del multi, one
                except Exception as ex:  # reported, not fatal: the timing above stands, the check did not run
                    parity["ok"] = None
                    parity["error"] = str(ex)[:300]
                finally:
                    if ref is not None:
                        ref.close()
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    dom = max(kt.items(), key=lambda kv: kv[1]["ms"]) if kt else (None, None)
    roof = None
    if dom[0]:
        name, r = dom
        achieved = r["bytes"] / (r["ms"] * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                ent = tj.get(args.workload, {}).get(name) if args.scale == 1.0 else None
                if ent:
                    traffic = ent.get("dram_bytes_per_launch")
            except Exception:
                pass
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "launches_timed": r["launches"],
                "avg_launch_ms": r["ms"] / r["launches"], "algorithmic_bytes_per_launch": r["bytes"] / r["launches"],
                "share_of_step": r["ms"] / sum(x["ms"] for x in kt.values())}

    # ---- CPU baseline: the oracle port on this box's host cores (N=1 only) ----
    threads = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu and world == 1:
        sample = min(nsym, args.cpu_sample)
        dt, nn, nblocks, cpu_lf = cpu_oracle_run(itype, data, sample, params, threads, want_lf=True)
        cpu = {"value": sample / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port", "seconds": dt,
               "sample": "first %d symbols of the workload, %d blocks, full pipeline; restated reference (libmaus2 unavailable)" % (sample, nblocks),
               "lf_steps_per_s": cpu_lf, "lf_instrument": "bwttestdecodespeed restated: 8 interleaved chains, 1 thread, BWT of the sample"}
        # throughput against n: the direction of the bias of every bounded CPU sample is on record
        curve = [{"n_symbols": int(sample), "value": sample / dt / 1e6, "seconds": dt}]
        for tok in [x for x in args.cpu_curve.split(",") if x.strip()]:
            k = min(nsym, int(tok))
            if any(c["n_symbols"] == k for c in curve):
                continue
            dtk, _, _ = cpu_oracle_run(itype, data, k, params, threads)
            curve.append({"n_symbols": int(k), "value": k / dtk / 1e6, "seconds": dtk})
        cpu["curve"] = sorted(curve, key=lambda c: c["n_symbols"])

    file_level = None
    if world == 1 and not args.no_file_level:
        try:
            file_level = file_level_leg(args, itype, data, nsym, params)
        except Exception as ex:  # the leg must never take the bench line down
            file_level = {"failed": str(ex)[:300]}
    if k8 is not None:
        peak_k8, _ = measured_peak_gbs()
        if k8.get("achieved_gbs"):
            k8["frac_of_hbm_peak"] = k8["achieved_gbs"] / peak_k8

    total_s = total_ms * 1e-3
    line = {
        "metric": METRIC, "value": nsym * args.steps / total_s / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(args, nsym, itype, {"numblocks": info["numblocks"], "preisarate": info["preisarate"],
                                                  "parallelism": ("single GPU" if world == 1 else
                                                                  "text replicated, %d suffix key ranges, every rank stores its BWT rows / samples into rank 0's HBM over NVLink (CUDA IPC peer stores from the sorting kernels), one all-reduce (NCCL) as vote and fence" % world if state["strategy"] == "shard" else
                                                                  "text replicated, %d block ranges, merge tree over NCCL" % world)}),
        "clocks": sampler.summary(),
        "e2e": {"value": nsym * e2e_steps / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                "outputs": ("BWA packed BWT words (b3m_engine_fetch_bwa)" if as_bwa else "BWT bytes") + " + anchors + sampled SA + sampled ISA",
                "path": ("single GPU: results stream out during the last sorting kernel" if world == 1 else
                         "every rank uploads 1/N of the input and sends 1/N of the BWA words and SA samples to ONE shared page-locked host buffer over its own PCIe link" if e2e_check is not None else
                         "every rank uploads 1/N of the input; results leave through rank 0"),
                "check": e2e_check, "phases_ms": e2e_phases,
                "numa": ({"gpu_node_rank0": shared["bwt"].node, "policy_set_rank0": bool(shared["bwt"].policy)} if shared is not None else None)},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "phases_ms": {k[3:]: round(info[k], 4) for k in info if k.startswith("ms_")},
        "counters": {k: info[k] for k in ("sort_rounds", "radix_passes", "radix_bytes", "sort_active_sum", "walk_lf_steps",
                                          "walk_chains", "gap_lf_steps", "max_lcpnext")},
        "kernels_ms_per_step": {k: round(v["ms"] / nprof, 4) for k, v in kt.items()},
        "lf_steps_per_s": (1 << 20) * 256 / (lf_ms * 1e-3),
    }
    if kt_max is not None:
        line["kernels_ms_per_step_max_over_ranks"] = kt_max
    if build_timeline is not None:
        line["build_timeline_ms"] = build_timeline
    if file_level is not None:
        line["file_level"] = file_level
    if k8 is not None:
        line["k8_rl_encode"] = k8
    if parity is not None:
        line["parity_check"] = parity
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only; invalid as a bench value)")
    ap.add_argument("--numblocks", type=int, default=1)
    ap.add_argument("--gapmode", default="auto", choices=["auto", "atomic", "list"], help="how K5 counts gap arrays (multi-block builds; include/b3m.h B3M_GAP_*)")
    ap.add_argument("--strategy", default="auto", choices=["auto", "shard", "merge"], help="multi-GPU decomposition (bwtb3m_b200.multigpu)")
    ap.add_argument("--cpu-sample", type=int, default=256_000_000, help="symbols of the workload the cpu_baseline leg processes")
    ap.add_argument("--ref-sample", type=int, default=128_000_000, help="symbols per step of --impl reference (about 7 s of CPU work per step)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-file-level", action="store_true", help="skip the file-in / files-out leg (b3m_compute_bwt on tmpfs) and the K8 timing")
    ap.add_argument("--cpu-curve", default="32000000,1000000000", help="further prefix sizes the cpu_baseline leg times (throughput against n); empty: none")
    ap.add_argument("--e2e-steps", type=int, default=5, help="steps of the end-to-end (host buffers) leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
