#!/usr/bin/env python3
"""bench.py -- CheckResources decisions/sec of the B200 evaluator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...       (N > 1, one rank per GPU)

A *step* is one pass of the hot path (rule-table scan + CEL condition evaluation) over one batch of
synthetic requests: workload C2 of SURVEY.md 8(d) = BASELINE.json configs[1] (10 resource policies x 8
actions, 2 derived roles with CEL on request.resource.attr, 2^20 requests = 8 388 608 decisions / step /
GPU).  Weak scaling: every rank evaluates its own 2^20-request shard (no data-path collective), then the
packed decision bitmaps are all-gathered over NCCL so every rank holds the whole result.

  value     whole-job decisions/s with the request columns already resident in HBM (CUDA events, max over
            ranks; successive steps rotate over distinct batches whose total footprint exceeds L2)
  e2e       the same metric through the host-buffer C-ABI call cgpu_check (pinned host columns -> H2D ->
            kernel -> D2H -> effect bytes), i.e. what engine.Check would pay per call
  roofline  algorithmic bytes / launch (73 B x 2^20) over the mean kernel duration vs the
            measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  oracle/c/check_ref.c (a plain-C port of the reference algorithm) on all host cores

--impl reference times that CPU port alone (the Go toolchain needed for the reference's own engine is not
in this image).
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

METRIC = "checkresources_decisions_per_sec"
UNIT = "decisions/s"
NOW_NS = 1_700_000_000_000_000_000


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run_nvml(self) -> bool:
        """NVML in-process: ~1 ms per sample (an nvidia-smi process takes ~100 ms, longer than a short timed region)."""
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ids = [x.strip() for x in vis.split(",") if x.strip()]
                if self.gpu < len(ids) and ids[self.gpu].isdigit():
                    idx = int(ids[self.gpu])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = str(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = [pynvml.nvmlClocksThrottleReasonHwSlowdown, pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, pynvml.nvmlClocksThrottleReasonSwPowerCap]
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001 -- no NVML: fall back to nvidia-smi
            return False
        while not self._stop.is_set():
            try:
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), mx, "0"] +
                                    ["Active" if r & b else "Not Active" for b in bits])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.002)
        return True

    def _run(self):
        if self._run_nvml():
            return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for nm, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def get_workload(name):
    from cerbos_b200 import workloads as W
    return W.WORKLOADS[name]()


def shard_fields(w, n, shard):
    """Fields of requests [shard*n, (shard+1)*n) of the workload's stream."""
    return w.fields(n, start=shard * n)


def cpu_port_rate(w, ft, enc, seconds=10.0, n=None, threads=None):
    """Decisions/s of the C port (oracle/c/check_ref.c) on all host cores over a bounded sample."""
    from oracle import cref
    import ctypes
    n = n or min(w.default_n, 1 << 20)
    threads = threads or (os.cpu_count() or 1)
    b = w.columns(w.fields(n), enc)
    cols = [np.ascontiguousarray(c) for c in b.columns]
    ptrs = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    sizes = (ctypes.c_size_t * len(cols))(*[c.nbytes for c in cols])
    bb = cref._Batch(b.n, b.max_actions, NOW_NS, 0, ptrs, sizes, len(cols))
    out = np.zeros((b.n, b.max_actions), dtype=np.uint8)
    buf = ctypes.create_string_buffer(ft.blob, len(ft.blob))
    lib = cref.lib()
    lib.cref_check(buf, len(ft.blob), ctypes.byref(bb), out.ctypes.data, threads)  # warm
    t0 = time.perf_counter()
    passes = 0
    while True:
        rc = lib.cref_check(buf, len(ft.blob), ctypes.byref(bb), out.ctypes.data, threads)
        assert rc == 0
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or passes >= 10000:
            break
    return passes * b.n * b.max_actions / dt, threads, passes, b.n, dt


def run_reference(args):
    """--impl reference: the CPU path alone, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = get_workload(args.workload)
    from cerbos_b200 import workloads as W
    _, ft, enc = W.build(w)
    # one "step" = a bounded sample: 2^20 requests (C2 full batch) on all host threads
    n = min(w.default_n, 1 << 20)
    from oracle import cref
    import ctypes
    threads = os.cpu_count() or 1
    b = w.columns(w.fields(n), enc)
    cols = [np.ascontiguousarray(c) for c in b.columns]
    ptrs = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    sizes = (ctypes.c_size_t * len(cols))(*[c.nbytes for c in cols])
    bb = cref._Batch(b.n, b.max_actions, NOW_NS, 0, ptrs, sizes, len(cols))
    out = np.zeros((b.n, b.max_actions), dtype=np.uint8)
    buf = ctypes.create_string_buffer(ft.blob, len(ft.blob))
    lib = cref.lib()
    for _ in range(args.warmup):
        lib.cref_check(buf, len(ft.blob), ctypes.byref(bb), out.ctypes.data, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = lib.cref_check(buf, len(ft.blob), ctypes.byref(bb), out.ctypes.data, threads)
        assert rc == 0
    dt = time.perf_counter() - t0
    value = args.steps * b.n * b.max_actions / dt
    sample = f"{b.n} requests x {b.max_actions} actions per step ({w.name} stream prefix), {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": w.name, "requests_per_step": b.n, "actions": b.max_actions,
                   "note": "CPU port of the reference algorithm (oracle/c/check_ref.c); the reference's Go engine "
                           "cannot be built in this image (no Go toolchain)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def ncu_traffic(workload: str, n: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed
    `ncu --set full` capture of this very command (profiles/; None if there is none for the workload / size)."""
    if workload != "C2" or n != (1 << 20):
        return None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_final_check_kernel_ncu_full.json")) as f:
            return float(json.load(f)["_traffic_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--requests", type=int, default=0, help="override requests per step per GPU (profiling only)")
    ap.add_argument("--streams", type=int, default=2, help="CUDA streams the steps are issued on round-robin (independent batches)")
    ap.add_argument("--no-verify", action="store_true", help="skip the post-run comparison of the result images with the oracle")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from cerbos_b200 import capi, workloads as W
    from cerbos_b200.device import DeviceBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cerbos_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    w = get_workload(args.workload)
    # rank 0 flattens the policies; the blob is broadcast over NCCL
    if rank == 0:
        _, ft, enc = W.build(w)
        blob = ft.blob
    from cerbos_b200.dist import all_gather_bitmaps, broadcast_blob
    blob = broadcast_blob(blob if rank == 0 else None, dev)
    from cerbos_b200.encode import Encoder, manifest_from_blob
    if rank != 0:
        enc = Encoder(manifest_from_blob(blob))
    ctx = capi.Context(local_rank)
    table = ctx.load_table(blob)
    spec_ready, spec_note = table.wait_ready()   # table-specialised kernels (NVRTC, background thread) are in place

    n = args.requests or w.default_n
    K = len(w.actions)
    n_buf = 4 if n >= (1 << 18) else 1
    # distinct batches: this rank's shard of requests, n_buf consecutive windows of the workload stream
    batches, host_batches = [], []
    for j in range(n_buf):
        hb = w.columns(shard_fields(w, n, rank * n_buf + j), enc)
        host_batches.append(hb)
        batches.append(DeviceBatch(hb, dev))
    footprint = sum(b.nbytes() for b in batches)
    kbytes = batches[0].kbytes

    calls = [b.prepare(table, NOW_NS) for b in batches]
    views = [b.bitmap[: n * kbytes] for b in batches]
    torch.cuda.synchronize()
    # Steps are independent batches: they are issued round-robin on `--streams` explicit streams (the tail of one step's
    # kernels and its peer-store round trips overlap the next step's kernel).  Everything timed is launched on these
    # streams through the C ABI and bracketed by CUDA events recorded on them.
    n_streams = max(1, min(args.streams, n_buf))
    while n_buf % n_streams:
        n_streams -= 1
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    stream = streams[0]
    torch.cuda.set_stream(stream)
    stream_h = stream.cuda_stream
    stream_hs = [st.cuda_stream for st in streams]
    assert all(h != 0 for h in stream_hs) and torch.cuda.current_stream().cuda_stream == stream_h

    # Multi-GPU result exchange.  Default: FUSED all-gather -- the check kernels store every result byte straight into
    # this rank's slice of every rank's gather buffer over NVLink peer memory (cerbos_b200.dist.PeerGather), a flag
    # release follows, no collective kernel runs.  Fallback (CERBOS_B200_NCCL_GATHER=1, or CUDA IPC unavailable):
    # asynchronous NCCL all_gather_into_tensor overlapping the next batch's kernel.
    gather_mode = "none"
    pg, gcalls = None, None
    if world > 1:
        ok = 0
        no_exchange = os.environ.get("CERBOS_B200_NO_GATHER") == "1"   # diagnosis only: ranks run independently, nothing is exchanged
        if os.environ.get("CERBOS_B200_NCCL_GATHER") != "1" and kbytes <= 8 and not no_exchange:
            try:
                from cerbos_b200.dist import PeerGather
                pg = PeerGather(ctx, n * kbytes, n_buf)
                gcalls = [table.prepared_gather_call(b.ptrs, b.sizes, b.n, b.max_actions, pg.bufs[j], pg.lane_flags(j % n_streams), rank, n * kbytes, NOW_NS)
                          for j, b in enumerate(batches)]
                ok = 1
            except Exception as e:  # noqa: BLE001 -- any failure here just selects the NCCL path on every rank
                sys.stderr.write(f"[rank {rank}] peer gather unavailable ({e}); using NCCL\n")
        t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        gather_mode = "fused-peer-stores" if int(t_ok.item()) == 1 else "nccl-all-gather"
        if gather_mode != "fused-peer-stores":
            pg, gcalls = None, None
        if no_exchange:
            gather_mode = "none (diagnosis)"
        elif pg is None:       # NCCL orders its collective after torch's current stream only: one issuing stream
            n_streams, streams, stream_hs = 1, streams[:1], stream_hs[:1]
    gathered = [torch.empty(world * n * kbytes, dtype=torch.uint8, device=dev) for _ in range(n_buf)] if (world > 1 and pg is None) else None
    pending = []
    it = [0]          # steps issued so far (warm-up included): numbers the gather steps

    def step(i):
        g = it[0]
        it[0] += 1
        j = g % n_buf
        lane = g % n_streams            # == j % n_streams: a buffer always travels on the same stream
        sh = stream_hs[lane]
        if pg is not None:
            # buffers rotate: step g-(n_buf-1) must have landed on this rank before its stream moves on (the wait rides in
            # the same launch: the kernel that publishes this step's flag also holds the stream for the older one).
            # Steps are numbered per stream ("lane"), so that every flag array only ever counts up.
            k = g - (n_buf - 1)
            if k >= 0:
                gcalls[j](g // n_streams + 1, sh, k // n_streams + 1, pg.local_flags(k % n_streams))
            else:
                gcalls[j](g // n_streams + 1, sh, 0, None)
            return
        calls[j](sh)
        if world > 1 and gather_mode != "none (diagnosis)":
            if len(pending) >= n_buf - 1:          # buffers are reused after n_buf steps: retire the oldest gather
                pending.pop(0).wait()
            _, work = all_gather_bitmaps(views[j], gathered[j], async_op=True)
            pending.append(work)

    def drain():
        if pg is not None:
            for lane in range(n_streams):          # the last step issued on every stream (a stream finishes its steps in order)
                last = [g for g in range(max(0, it[0] - n_streams), it[0]) if g % n_streams == lane]
                if last:
                    pg.wait(last[-1] // n_streams + 1, stream_hs[lane], lane)
            return
        while pending:
            pending.pop(0).wait()

    def join_streams():
        """stream 0 waits for the work issued so far on the other streams"""
        for st in streams[1:]:
            e = torch.cuda.Event()
            e.record(st)
            stream.wait_event(e)

    if world > 1 and pg is None:
        assert n_streams == 1 or gather_mode == "none (diagnosis)"
    for i in range(args.warmup):
        step(i)
    drain()
    for h in stream_hs:
        ctx.sync(h)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        ev0.record(stream)
        for st in streams[1:]:
            st.wait_event(ev0)               # no stream starts before the start event
        for i in range(args.steps):
            step(i)
        drain()                              # the last exchanges are part of the timed work
        join_streams()
        ev1.record(stream)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    for h in stream_hs:
        ctx.sync(h)
    if world > 1:
        tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
    per_step_ms = total_ms / args.steps
    value = world * n * K / (per_step_ms * 1e-3)

    # kernel-only duration: events around each single launch (no collective), measured after the timed region
    # (a) the whole device step of one call (clustering kernels + check kernel), (b) the check kernel alone, from
    # the library's own CUDA events recorded on the launching stream around that kernel (cgpu_profile)
    step_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.profile(True)
    for i in range(min(args.steps, 50)):
        e0.record()
        calls[i % n_buf](stream_h)
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    k_sum, k_n = ctx.profile(False)
    kern_ms_mean = k_sum / max(k_n, 1)
    kern_ms = step_ms
    peak, peak_src = load_peaks()
    algo_bytes = w.bytes_per_request() * n
    achieved = algo_bytes / (kern_ms_mean * 1e-3) / 1e9

    # correctness check of what was timed (the overlapped back-to-back launches included): the result images the timed
    # loop left behind, every rotating batch, bit for bit against the C port of the reference algorithm (oracle/)
    verified = None
    if not args.no_verify and n <= (1 << 21):
        from oracle import cref as _cref
        bad = 0
        for j, hb in enumerate(host_batches):
            want = _cref.check(blob, hb.columns, hb.n, hb.max_actions, NOW_NS, 0, n_threads=os.cpu_count() or 1)
            want_bits = np.packbits((want == 1).astype(np.uint8), axis=1, bitorder="little")[:, :kbytes].reshape(-1)
            if pg is not None:
                img = pg.read(j)
                got = img[rank * n * kbytes:(rank + 1) * n * kbytes]
            else:
                got = batches[j].bitmap[: n * kbytes].cpu().numpy()
            bad += int((got != want_bits).sum())
        if world > 1:
            tb = torch.tensor([bad], dtype=torch.int64, device=dev)
            dist.all_reduce(tb)
            bad = int(tb.item())
            if pg is not None and bad == 0:
                # every rank also holds every other rank's slice: compare the whole gathered image across ranks
                h = torch.tensor([int(np.frombuffer(pg.read(0).tobytes(), dtype=np.uint64).sum() & ((1 << 62) - 1))], dtype=torch.int64, device=dev)
                hs = [torch.zeros_like(h) for _ in range(world)]
                dist.all_gather(hs, h)
                bad = 0 if len({int(x.item()) for x in hs}) == 1 else -1
        verified = bad == 0
        if not verified:
            sys.stderr.write(f"[rank {rank}] VERIFY FAILED: {bad} result bytes differ from the oracle\n")

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{w.name}: {W.C2.__doc__.splitlines()[0] if w.name == 'C2' else w.name}",
                   "requests_per_step_per_gpu": n, "actions_per_request": K, "global_requests_per_step": world * n,
                   "parallelism": f"dp{world} (requests sharded by index; NCCL table broadcast; result exchange: {gather_mode})",
                   "streams": n_streams,
                   "l2": f"rotating {n_buf} distinct batches, {footprint / 1e6:.0f} MB of columns > 126 MB L2",
                   "kernel": ctx.last_kernel_config()},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(w.name, n), "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                     "kernel": "check_kernel", "kernel_ms_mean": kern_ms_mean,
                     "device_step_ms_mean": statistics.mean(step_ms), "device_step_ms_min": min(step_ms),
                     "step_achieved": algo_bytes / (statistics.mean(step_ms) * 1e-3) / 1e9,
                     "kernel_share_of_step": kern_ms_mean / statistics.mean(step_ms),
                     "note": "kernel_ms_mean: the check kernel alone, CUDA events recorded by the library on the launching stream around "
                             "that kernel, one call at a time after the timed region; device_step_ms_*: events around one whole "
                             "call issued alone (includes the launch gaps between its kernels, which back-to-back calls hide: "
                             "ms_per_step is lower because consecutive calls overlap at their tails)"},
        "clocks": clocks.summary(),
        "verified_vs_oracle": verified,
    }

    if rank == 0 and not args.no_e2e:
        # end to end through the host-buffer C ABI: pinned host columns, H2D + kernel + D2H + decode inside the timing
        hb = host_batches[0]
        pinned = []
        for c in hb.columns:
            a = np.ascontiguousarray(c)
            t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
            t.numpy()[:] = a.view(np.uint8).reshape(-1)
            pinned.append(t)
        ptrs = [t.data_ptr() for t in pinned]
        sizes = [t.numel() for t in pinned]
        out = torch.empty(hb.n * hb.max_actions, dtype=torch.uint8).pin_memory()
        for _ in range(3):
            table.check_into(ptrs, sizes, hb.n, hb.max_actions, out.data_ptr(), NOW_NS, 0)
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            table.check_into(ptrs, sizes, hb.n, hb.max_actions, out.data_ptr(), NOW_NS, 0)
        dt = (time.perf_counter() - t0) / args.e2e_steps
        result["e2e"] = {"value": hb.n * K / dt, "unit": UNIT, "h2d_bytes_per_step": int(sum(sizes)),
                         "d2h_bytes_per_step": int(hb.n * kbytes), "ms_per_step": dt * 1e3, "n_gpus": 1,
                         "note": "cgpu_check: pinned host columns -> H2D -> kernel -> D2H bitmap -> effect bytes"}
    if rank == 0 and world == 1 and not args.no_cpu:   # the CPU baseline is reported at N = 1 only
        v, threads, passes, ns, dt = cpu_port_rate(w, ft, enc, seconds=args.cpu_seconds)
        result["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                  "sample": f"{passes} passes over {ns} requests x {K} actions ({dt:.1f} s) of the "
                                            f"same {w.name} stream, oracle/c/check_ref.c"}
    if rank == 0:
        print(json.dumps(result))
    table.release()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
