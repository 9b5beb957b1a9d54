#!/usr/bin/env python3
"""bench.py - headline benchmark of the zfp hot path on B200 (see BASELINE.json / DESIGN.md).

One "step" = zfp_compress followed by zfp_decompress of one device-resident slab of a synthetic
smooth 3-D fp64 field at fixed rate 8 bits/value (BASELINE.json configs[1]: 1024^3 fp64, rate 8).
With N GPUs the global field is (1024*N) x 1024 x 1024 split into N block-aligned slabs along the
slowest dimension, one per rank, no data-path collective (fixed-rate slabs sit at deterministic bit
offsets) -> weak scaling.

metric = (uncompressed bytes compressed + uncompressed bytes decompressed) / time, GB = 1e9 B.
(`roundtrip_gbs` = uncompressed bytes of ONE array / step time is printed beside it.)

Extra keys of the line (all measured in this run):
  variable_rate   second timed leg on every N: the same slab at fixed accuracy 1e-6 - encode, exchange of the
                  slab bit lengths (all_gather, stream ordered: no host round trip), device prefix, decode -
                  with the slab bit lengths and a check that the slabs tile the global stream
  multi_gpu_parity (N >= 2) a small field compressed as N slabs and assembled over NCCL equals the stream ONE
                  GPU produces for the whole field (fixed rate and fixed accuracy)
  other_configs   (N == 1) the other BASELINE.json configurations that fit one GPU, compress / decompress ms and
                  fraction of the measured HBM peak
  ref_cuda        (N == 1) the backend this project replaces (reference src/cuda_zfp, unmodified, built for
                  sm_100 by oracle/Makefile) timed on the same device-resident 1024^3 field, streams compared

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0.
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

METRIC = "compress+decompress GB/s (uncompressed), 3D fp64 fixed-rate"
UNIT = "GB/s"
RATE = 8
SIDE = 1024          # per-GPU slab is SIDE^3 values (8 GiB fp64)
CPU_SIDE = 512       # bounded CPU sample of the in-run cpu_baseline leg: CPU_SIDE^3 sub-box of the same field (1 GiB)
VAR_MODE = {"accuracy": 1e-6}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(md_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the kernel on this exact workload, from
    the `ncu --set full` capture summarised under profiles/ (a profiler run, not this run); None if absent."""
    path = os.path.join(ROOT, "profiles", md_name)
    try:
        with open(path) as f:
            for line in f:
                if line.startswith("| traffic (dram read+write)"):
                    cells = [c.strip() for c in line.strip().strip("|").split("|")]
                    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(cells[2], None)
                    return float(cells[1]) * scale if scale else None
    except OSError:
        pass
    return None


def field_slab(torch, rank, nz, ny, nx, device):
    """SURVEY 8(d) S1 analytic smooth field, slab `rank` of the global array, generated on device."""
    gz = torch.arange(rank * nz, (rank + 1) * nz, device=device, dtype=torch.float64) / (nz * max(1, int(os.environ.get("WORLD_SIZE", "1"))) - 1)
    gy = torch.arange(ny, device=device, dtype=torch.float64) / (ny - 1)
    gx = torch.arange(nx, device=device, dtype=torch.float64) / (nx - 1)
    out = torch.empty((nz, ny, nx), dtype=torch.float64, device=device)
    chunk = 64
    for z0 in range(0, nz, chunk):
        z = gz[z0:z0 + chunk, None, None]
        y = gy[None, :, None]
        x = gx[None, None, :]
        out[z0:z0 + chunk] = torch.sin(2 * np.pi * (3 * x + 0.5 * y)) * torch.cos(4 * np.pi * z) + 0.25 * torch.sin(14 * np.pi * x * y * z)
    return out


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML
    polled every ~5 ms in a thread (nvidia-smi -lms as the fallback); only samples taken between
    mark_begin() and mark_end() count."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None
        self.t0 = self.t1 = None
        self.max_mhz = None
        self.how = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.how = "nvml"
            while not self.stop_flag:
                t = time.perf_counter()
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((t, mhz, mask))
                time.sleep(0.005)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        bits = [0x8, 0x40, 0x20, 0x4]
        try:
            self.how = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                r = [v.strip() for v in line.split(",")]
                if len(r) >= 6 and r[0].replace(".", "").isdigit():
                    self.max_mhz = float(r[1])
                    mask = sum(b for b, v in zip(bits, r[2:6]) if v.lower().startswith("active"))
                    self.rows.append((time.perf_counter(), float(r[0]), mask))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        inside = [r for r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        use = inside if inside else self.rows[-3:]
        mask = 0
        for r in use:
            mask |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in use])) if use else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "samples": len(inside), "source": self.how}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref), all host threads it can use
# ---------------------------------------------------------------------------------------------------
def cpu_sample_field(side):
    ax = np.linspace(0.0, 1.0, SIDE)[:side]
    z, y, x = ax[:, None, None], ax[None, :, None], ax[None, None, :]
    return np.ascontiguousarray(np.sin(2 * np.pi * (3 * x + 0.5 * y)) * np.cos(4 * np.pi * z) + 0.25 * np.sin(14 * np.pi * x * y * z))


def time_reference(a, steps, warmup, threads):
    """compress with zfp_exec_omp (all threads) + decompress serial (the reference has no parallel
    decompress, src/zfp.c:1137-1138); returns per-step seconds (compress, decompress), stream words."""
    from oracle.oracle import Reference
    R = Reference()
    n = list(reversed(a.shape)) + [0]
    flat = a.reshape(-1)
    out = np.empty_like(a)
    buf = np.zeros(R.maximum_size({"rate": RATE}, a.dtype, n) // 8 + 4, dtype=np.uint64)
    tc, td, words = [], [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        words, _ = R.compress_raw(flat, 0, a.dtype, n, None, {"rate": RATE}, policy=R.OMP if threads > 1 else R.SERIAL,
                                  threads=threads, out=buf)
        t1 = time.perf_counter()
        R.decompress_raw_noalloc(words, out.reshape(-1), a.dtype, n, {"rate": RATE})
        t2 = time.perf_counter()
        if i >= warmup:
            tc.append(t1 - t0)
            td.append(t2 - t1)
    return tc, td, words, out


def host_field(side):
    """The bench field (rank 0's slab of field_slab) on the host, generated in chunks with torch CPU kernels."""
    import torch
    out = torch.empty((side, SIDE, SIDE), dtype=torch.float64)
    g = torch.arange(SIDE, dtype=torch.float64) / (SIDE - 1)
    y, x = g[None, :, None], g[None, None, :]
    for z0 in range(0, side, 32):
        z = g[z0:z0 + 32, None, None]
        out[z0:z0 + 32] = torch.sin(2 * np.pi * (3 * x + 0.5 * y)) * torch.cos(4 * np.pi * z) + 0.25 * torch.sin(14 * np.pi * x * y * z)
    return out.numpy()


def time_parallel_decompress(words, a, threads):
    """Row f4 (oracle/ref_parallel_decompress.c): chunk-parallel fixed-rate decompress around the reference's block
    API; GB/s of uncompressed data, or None when the helper library is not built."""
    try:
        from oracle.oracle import parallel_decompress
        out = np.empty_like(a)
        w = np.concatenate([words, np.zeros(2, dtype=np.uint64)])
        parallel_decompress(w, a.shape, a.dtype, RATE * 64, threads, out=out)
        t0 = time.perf_counter()
        parallel_decompress(w, a.shape, a.dtype, RATE * 64, threads, out=out)
        return a.nbytes / (time.perf_counter() - t0) / 1e9
    except Exception:
        return None


def run_reference(args):
    """The reference's own CPU implementation on the box's host cores, SAME configuration as our arm: the full
    1024^3 fp64 slab at rate 8 per step (zfp_exec_omp compress with all threads + serial decompress - upstream
    has no parallel decompress, src/zfp.c:1137-1138).  Falls back to a 512^3 sub-box, and says so, only when
    the host cannot hold the 17 GiB of buffers."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    side = SIDE
    try:
        import psutil
        if psutil.virtual_memory().available < 22 * 2 ** 30:
            side = CPU_SIDE
    except Exception:
        pass
    if args.quick:
        side = 256
    a = host_field(side) if side == SIDE else cpu_sample_field(side)
    tc, td, ref_words, _ = time_reference(a, args.steps, args.warmup, threads)
    pdec = time_parallel_decompress(ref_words, a, threads)
    step = float(np.sum(tc) + np.sum(td)) / args.steps
    value = 2 * a.nbytes / step / 1e9
    same = side == SIDE
    workload = "3D fp64 %d^3 per GPU, fixed-rate %d, compress+decompress" % (SIDE, RATE)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "roundtrip_gbs": a.nbytes / step / 1e9,
        "config": {"workload": workload,
                   "sample": "the full %d^3 slab per step (same configuration as the GPU arm)" % SIDE if same else
                             "%d^3 sub-box of the same analytic field per step (host memory too small for the full slab)" % side},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": "%d x %d x %d fp64 (%.2f GiB) per step; zfp_exec_omp compress with %d threads (%.3f GB/s) + serial decompress (%.3f GB/s; reference has no parallel decompress)"
                                   % (a.shape[0], a.shape[1], a.shape[2], a.nbytes / 2 ** 30, threads, a.nbytes / np.mean(tc) / 1e9, a.nbytes / np.mean(td) / 1e9),
                         "chunk_parallel_decompress_gbs": pdec,
                         "chunk_parallel_note": "not part of `value`: an OpenMP driver of ours around the reference's block API (oracle/ref_parallel_decompress.c, scope row f4), what upstream announces in docs/source/execution.rst:302-304"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# extra legs of our arm
# ---------------------------------------------------------------------------------------------------
def variable_rate_leg(torch, dist, zb, x, y, rank, world, steps, warmup, barrier):
    """Fixed accuracy 1e-6 on the same slab: encode -> all_gather of the slab bit lengths -> device prefix ->
    decode, everything enqueued on one stream (zfp_b200/distributed.py DeviceSlab); the host synchronises only
    at the ends of the timed region."""
    from zfp_b200.distributed import DeviceSlab
    dev = x.device
    slab = DeviceSlab(tuple(x.shape), x.dtype, VAR_MODE, rank, world)
    for _ in range(max(warmup, 3)):
        slab.compress(x)
        slab.decompress(y)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    launches0 = zb.launch_count()
    ev[0].record()
    for i in range(steps):
        slab.compress(x)
        ev[2 * i + 1].record()
        slab.decompress(y)
        ev[2 * i + 2].record()
    barrier()
    launches = zb.launch_count() - launches0
    total = ev[0].elapsed_time(ev[-1])
    enc = float(np.mean([ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps)]))
    dec = float(np.mean([ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(steps)]))
    t = torch.tensor([total, enc, dec], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total, enc, dec = [float(v) for v in t.tolist()]
    # the slabs tile the global stream: every rank's base is the sum of the lower ranks' lengths
    bases = torch.zeros(world, dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(bases, slab.base)
    else:
        bases.copy_(slab.base)
    lengths = [int(v) for v in slab.lengths.tolist()]
    bases = [int(v) for v in bases.tolist()]
    tiles = all(bases[r] == sum(lengths[:r]) for r in range(world)) and all(v > 0 for v in lengths)
    # the decoded slab honours the tolerance
    err = float((y - x).abs().max().item())
    index_ok = int(slab.status.item()) == 0
    raw = x.numel() * x.element_size()
    out = {"mode": "fixed accuracy 1e-6", "ms_per_step": total / steps, "compress_ms": enc, "decompress_ms": dec,
           "value": 2 * raw * world / (total / steps * 1e-3) / 1e9, "unit": UNIT, "slab_bits": lengths, "slab_base_bits": bases,
           "slabs_tile_the_stream": bool(tiles), "max_abs_error": err, "within_tolerance": bool(err <= VAR_MODE["accuracy"]), "index_check_clean": bool(index_ok),
           "compression_ratio": raw * 8.0 / max(1, lengths[rank]), "gpu_launches": launches,
           "exchange": "all_gather_into_tensor of one int64 per rank on the compute stream, prefix + placement on the device (zfp_b200_bitcopy_ranked); no host synchronisation inside a step" if world > 1 else "single rank: no exchange"}
    slab.close()
    return out


def multi_gpu_parity(torch, dist, zb, rank, world, dev):
    """A small field compressed as `world` slabs, assembled into one stream over NCCL, against the stream ONE
    GPU produces for the whole field.  Fixed rate (deterministic offsets) and fixed accuracy (exchanged lengths)."""
    from zfp_b200.distributed import DeviceSlab, plan_slabs
    shape = (16 * world + 8, 52, 60)
    g = [torch.arange(n, device=dev, dtype=torch.float64) / (n - 1) for n in shape]
    whole = torch.sin(5 * g[0][:, None, None] + 3 * g[1][None, :, None]) * torch.cos(4 * g[2][None, None, :]) + 0.1 * g[0][:, None, None] ** 2
    plan = plan_slabs(shape, world)[rank]
    mine = whole[plan.z0:plan.z1].contiguous()
    ok = True
    for mode in ({"rate": 8}, {"accuracy": 1e-5}):
        ref = zb.compress(whole, **mode)
        nwords = ref.nbytes // 8
        out = torch.zeros(nwords + 2, dtype=torch.int64, device=dev)
        slab = DeviceSlab(tuple(mine.shape), mine.dtype, mode, rank, world)
        slab.compress(mine, global_words=out)
        dist.all_reduce(out, op=dist.ReduceOp.SUM)  # slabs are disjoint bit ranges of a zeroed buffer: sum == bitwise or
        ok = ok and bool(torch.equal(out[:nwords], ref.words[:nwords].view(torch.int64)))
        back = torch.empty_like(mine)
        slab.decompress(back)
        ok = ok and bool(torch.equal(back, zb.decompress(ref)[plan.z0:plan.z1]))
        slab.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def time_pair(torch, zb, x, mode, reps=5):
    c = zb.compress(x, **mode)
    y = torch.empty_like(x)
    zb.decompress(c, out=y)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tc = td = 0.0
    for _ in range(reps):
        e[0].record()
        c = zb.compress(x, reuse=c, **mode)
        e[1].record()
        zb.decompress(c, out=y)
        e[2].record()
        torch.cuda.synchronize()
        tc += e[0].elapsed_time(e[1]) / reps
        td += e[1].elapsed_time(e[2]) / reps
    return tc, td, c.nbytes


def other_configs(torch, zb, dev, peak):
    """BASELINE.json configurations besides the headline that fit one GPU (device-resident, same analytic
    field family): compress / decompress time and fraction of the measured HBM peak (uncompressed + compressed
    bytes over the time of the call)."""
    def field(shape, dtype):
        g = [torch.arange(n, device=dev, dtype=torch.float64) / (n - 1) for n in shape]
        if len(shape) == 3:
            z, y, x = g[0][:, None, None], g[1][None, :, None], g[2][None, None, :]
            f = torch.empty(shape, dtype=torch.float64, device=dev)
            for z0 in range(0, shape[0], 64):
                zz = z[z0:z0 + 64]
                f[z0:z0 + 64] = torch.sin(2 * np.pi * (3 * x + 0.5 * y)) * torch.cos(4 * np.pi * zz) + 0.25 * torch.sin(14 * np.pi * x * y * zz)
        elif len(shape) == 2:
            y, x = g[0][:, None], g[1][None, :]
            f = torch.sin(2 * np.pi * (3 * x + 0.5 * y)) + 0.25 * torch.sin(14 * np.pi * x * y)
        elif len(shape) == 1:
            f = torch.sin(40 * np.pi * g[0]) + 0.25 * torch.sin(300 * np.pi * g[0] ** 2)
        else:
            w, z, y, x = g[0][:, None, None, None], g[1][None, :, None, None], g[2][None, None, :, None], g[3][None, None, None, :]
            f = torch.sin(2 * np.pi * (x + 0.5 * y)) * torch.cos(3 * np.pi * z) + 0.25 * torch.sin(5 * np.pi * w * x)
        return (torch.round(f * 2 ** 20) if dtype in (torch.int32, torch.int64) else f).to(dtype)

    cases = [("3D fp64 1024^3 rate 4", (SIDE, SIDE, SIDE), torch.float64, {"rate": 4}),
             ("3D fp64 1024^3 rate 16", (SIDE, SIDE, SIDE), torch.float64, {"rate": 16}),
             ("3D fp64 1024^3 precision 32", (SIDE, SIDE, SIDE), torch.float64, {"precision": 32}),
             ("3D fp32 1024^3 rate 8", (SIDE, SIDE, SIDE), torch.float32, {"rate": 8}),
             ("2D fp32 16384^2 rate 8", (16384, 16384), torch.float32, {"rate": 8}),
             ("4D fp64 64^4 rate 8", (64, 64, 64, 64), torch.float64, {"rate": 8}),
             ("1D fp64 2^28 rate 8", (1 << 28,), torch.float64, {"rate": 8}),
             ("3D int32 1024^3 reversible", (SIDE, SIDE, SIDE), torch.int32, {"reversible": True})]
    out, cached = {}, (None, None, None)
    for name, shape, dtype, mode in cases:
        try:
            if cached[0] != (shape, dtype):
                cached = (None, None, None)
                torch.cuda.empty_cache()
                cached = ((shape, dtype), field(shape, dtype), None)
            x = cached[1]
            tc, td, nbytes = time_pair(torch, zb, x, mode, reps=3)
            raw = x.numel() * x.element_size()
            out[name] = {"compress_ms": round(tc, 3), "decompress_ms": round(td, 3), "ratio": round(raw / nbytes, 2),
                         "compress_hbm_frac": round((raw + nbytes) / (tc * 1e-3) / 1e9 / peak, 3),
                         "decompress_hbm_frac": round((raw + nbytes) / (td * 1e-3) / 1e9 / peak, 3)}
        except Exception as ex:  # report, never fake
            out[name] = {"error": repr(ex)[:160]}
    cached = None
    torch.cuda.empty_cache()
    return out


def ref_cuda_leg(torch, zb, x, ours_stream_words, ours_nbytes):
    """The backend this project replaces - reference src/cuda_zfp, unmodified, compiled for sm_100 into
    oracle/_ref/libzfp_ref_cudaorig.so - on the same device-resident slab at rate 8 through the reference's
    own zfp_compress / zfp_decompress under zfp_exec_cuda; its stream must equal ours."""
    so = os.path.join(ROOT, "oracle", "_ref", "libzfp_ref_cudaorig.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/libzfp_ref_cudaorig.so not built"}
    from oracle.oracle import Reference, ZFP_TYPE
    R = Reference(so)
    L = R.L
    n = tuple(reversed(x.shape)) + (0,)
    words = torch.zeros(ours_nbytes // 8 + 16, dtype=torch.int64, device=x.device)
    y = torch.empty_like(x)
    f, dims = R._field(x.data_ptr(), np.float64, n, None)
    z = L.zfp_stream_open(None)
    L.zfp_stream_set_rate(z, float(RATE), ZFP_TYPE[np.dtype(np.float64)], dims, 0)
    bs = L.stream_open(words.data_ptr(), words.numel() * 8)
    L.zfp_stream_set_bit_stream(z, bs)
    if not L.zfp_stream_set_execution(z, 2):
        return {"unavailable": "reference library built without CUDA"}

    def run(fn, ptr, reps=3):
        L.zfp_field_set_pointer(f, ptr)
        L.zfp_stream_rewind(z)
        nb = fn(z, f)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            L.zfp_stream_rewind(z)
            fn(z, f)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, nb

    tc, nbytes = run(L.zfp_compress, x.data_ptr())
    same = bool(nbytes == ours_nbytes and torch.equal(words[: nbytes // 8], ours_stream_words[: nbytes // 8].view(torch.int64)))
    td, _ = run(L.zfp_decompress, y.data_ptr())
    raw = x.numel() * 8
    L.zfp_field_free(f); L.zfp_stream_close(z); L.stream_close(bs)
    return {"compress_ms": tc, "decompress_ms": td, "value": 2 * raw / ((tc + td) * 1e-3) / 1e9, "unit": UNIT,
            "stream_identical_to_ours": same, "what": "reference src/cuda_zfp (unmodified) built for sm_100, zfp_exec_cuda, same device-resident slab"}


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import zfp_b200 as zb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    zb.load_library(build_if_missing=False)

    nz = ny = nx = SIDE
    x = field_slab(torch, rank, nz, ny, nx, dev)
    raw_bytes = x.numel() * 8
    mode = {"rate": RATE}
    words = torch.empty(zb.max_stream_words(x.shape, x.dtype, mode), dtype=torch.int64, device=dev)
    y = torch.empty_like(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prev = [None]

    def step(record=None):
        if record: record[0].record()
        # the first call creates the zfp_stream over `words`; later calls rewind and recycle it
        c = zb.compress(x, reuse=prev[0], **mode) if prev[0] is not None else zb.compress(x, out=words, async_fixed_rate=True, **mode)
        prev[0] = c
        if record: record[1].record()
        zb.decompress(c, out=y)
        if record: record[2].record()
        return c

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        c = step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = zb.launch_count()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    t_begin.record()
    for i in range(args.steps):
        c = step(evs[i])
    t_end.record()
    barrier()
    sampler.mark_end()
    launches = zb.launch_count() - launches0
    clocks = sampler.finish()
    total_ms = t_begin.elapsed_time(t_end)
    enc_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    dec_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    tmax = torch.tensor([total_ms, enc_ms, dec_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms = [float(v) for v in tmax.tolist()]
    ms_per_step = total_ms / args.steps
    value = 2 * raw_bytes * world / (ms_per_step * 1e-3) / 1e9
    comp_bytes = c.nbytes

    # ---- end to end through the public C API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    try:
        # pinned host buffers: field in, field out, stream.  All ranks of the node pin at once, so the
        # slab used for this leg is cut down (whole block layers) if the node's free memory is short.
        ez = nz
        try:
            import psutil
            budget = 0.5 * psutil.virtual_memory().available / max(1, world)
            per_layer = (2 * 8 + RATE / 8.0) * ny * nx
            ez = int(min(nz, max(64, (budget / per_layer) // 4 * 4)))
        except Exception:
            pass
        e_raw = ez * ny * nx * 8
        e_comp = comp_bytes * ez // nz
        hx = torch.empty((ez, ny, nx), dtype=torch.float64, pin_memory=True)
        hx.copy_(x[:ez])
        hy = torch.empty((ez, ny, nx), dtype=torch.float64, pin_memory=True)
        hw = torch.zeros(words.numel() * ez // nz + 64, dtype=torch.int64, pin_memory=True)
        L = zb.load_library()
        from zfp_b200.api import Stream, _make_field
        s = Stream(hw.data_ptr(), hw.numel() * 8, mode, 4, 3)
        fin = _make_field(L, hx.data_ptr(), 4, (ez, ny, nx), None)
        fout = _make_field(L, hy.data_ptr(), 4, (ez, ny, nx), None)

        def e2e_step():
            L.zfp_stream_rewind(s.z)
            nb = L.zfp_compress(s.z, fin)       # H2D field, kernels, D2H stream
            L.zfp_stream_rewind(s.z)
            nb2 = L.zfp_decompress(s.z, fout)   # H2D stream, kernels, D2H field
            assert nb == nb2 == e_comp, (nb, nb2, e_comp)

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": 2 * e_raw * world / (float(dt.item()) / e2e_steps) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": e_raw + e_comp, "d2h_bytes_per_step": e_raw + e_comp,
               "steps": e2e_steps, "note": "zfp_compress/zfp_decompress on pinned host field + host stream buffer"
                                            + ("" if ez == nz else "; first %d of %d layers per rank (host memory)" % (ez, nz))}
        same = bool(torch.equal(hy, y[:ez].cpu()))
        e2e["matches_device_path"] = same
        L.zfp_field_free(fin)
        L.zfp_field_free(fout)
        s.close()
        del hx, hy, hw
    except Exception as ex:  # report, never fake
        e2e = {"value": None, "unit": UNIT, "error": repr(ex)[:200]}

    # ---- second timed leg: variable rate with the stream-ordered slab-length exchange (every rank)
    try:
        var = variable_rate_leg(torch, dist, zb, x, y, rank, world, args.steps, args.warmup, barrier)
    except Exception as ex:  # report, never fake
        var = {"error": repr(ex)[:300]}
    parity_multi = None
    if world > 1:
        try:
            parity_multi = multi_gpu_parity(torch, dist, zb, rank, world, dev)
        except Exception as ex:
            parity_multi = "error: " + repr(ex)[:200]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    values = x.numel()
    enc_alg = values * (8 + RATE / 8.0)   # bytes: read fp64 + write RATE bits per value
    dec_alg = enc_alg
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch on this exact workload, from the ncu
    # --set full captures summarised in profiles/ (a profiler run, not this run)
    roof_enc = {"kernel": "encode_staged_kernel<double,3>", "bound": "hbm", "achieved": enc_alg / (enc_ms * 1e-3) / 1e9,
                "peak": peak, "unit": "GB/s", "traffic": profiled_traffic("r2_encode_staged_fp64_r8_1024cubed.md"),
                "traffic_source": "profiles/r2_encode_staged_fp64_r8_1024cubed.md",
                "algorithmic_bytes": enc_alg, "peak_source": peak_src, "ms": enc_ms}
    roof_enc["frac"] = roof_enc["achieved"] / peak
    roof_dec = {"kernel": "decode_staged_kernel<double,3>", "bound": "hbm", "achieved": dec_alg / (dec_ms * 1e-3) / 1e9,
                "peak": peak, "unit": "GB/s", "traffic": profiled_traffic("r2_decode_staged_fp64_r8_1024cubed.md"),
                "traffic_source": "profiles/r2_decode_staged_fp64_r8_1024cubed.md",
                "algorithmic_bytes": dec_alg, "peak_source": peak_src, "ms": dec_ms}
    roof_dec["frac"] = roof_dec["achieved"] / peak
    dominant = roof_dec if dec_ms >= enc_ms else roof_enc

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "3D fp64 %dx%dx%d per GPU (global %dx%dx%d), fixed-rate %d bits/value, zfp_compress + zfp_decompress, device resident"
                               % (nz, ny, nx, nz * world, ny, nx, RATE),
                   "parallelism": "slab-per-GPU x%d, no collective" % world,
                   "l2": "inputs (8 GiB) and outputs exceed the 126 MB L2; no explicit flush"},
        "roundtrip_gbs": raw_bytes * world / (ms_per_step * 1e-3) / 1e9,
        "compress_gbs": raw_bytes * world / (enc_ms * 1e-3) / 1e9, "decompress_gbs": raw_bytes * world / (dec_ms * 1e-3) / 1e9,
        "compressed_bytes_per_gpu": comp_bytes,
        "roofline": dominant, "roofline_encode": roof_enc, "roofline_decode": roof_dec,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "variable_rate": var,
    }
    if parity_multi is not None:
        line["multi_gpu_parity"] = parity_multi
    if world == 1 and not args.quick:
        try:
            line["ref_cuda"] = ref_cuda_leg(torch, zb, x, c.words, comp_bytes)
            if line["ref_cuda"].get("value"):
                line["ref_cuda"]["ours_over_ref_cuda"] = value / line["ref_cuda"]["value"]
        except Exception as ex:
            line["ref_cuda"] = {"error": repr(ex)[:200]}

    # ---- CPU baseline on this box's host cores (N=1 only), bounded sample, plus a parity check on it
    if world == 1 and not args.no_cpu:
        try:
            threads = os.cpu_count() or 1
            side = CPU_SIDE if not args.quick else 256
            a = x[:side, :side, :side].contiguous().cpu().numpy()
            tc, td, ref_words, ref_out = time_reference(a, 2, 1, threads)
            tcs, tds, _, _ = time_reference(a[: side // 2], 1, 0, 1)
            pdec = time_parallel_decompress(ref_words, a, threads)
            cs = zb.compress(torch.from_numpy(a).to(dev), **mode)
            parity = cs.to_numpy().tobytes() == ref_words.tobytes() and zb.decompress(cs).cpu().numpy().tobytes() == ref_out.tobytes()
            val = 2 * a.nbytes / (np.mean(tc) + np.mean(td)) / 1e9
            line["cpu_baseline"] = {
                "value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": "%d^3 sub-box of the bench field; zfp_exec_omp compress x%d threads %.3f GB/s + serial decompress %.3f GB/s (no parallel decompress upstream); serial compress %.3f GB/s on half the sample"
                          % (side, threads, a.nbytes / np.mean(tc) / 1e9, a.nbytes / np.mean(td) / 1e9, a.nbytes / 2 / np.mean(tcs) / 1e9),
                "chunk_parallel_decompress_gbs": pdec,
                "stream_and_array_bit_identical_to_gpu": bool(parity)}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "error": repr(ex)[:200]}
    if world == 1 and not args.quick and not args.no_others:
        prev[0] = c = None
        del x, y, words
        torch.cuda.empty_cache()
        try:
            line["other_configs"] = other_configs(torch, zb, dev, peak)
        except Exception as ex:
            line["other_configs"] = {"error": repr(ex)[:200]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="smaller CPU sample, no ref_cuda / other_configs legs")
    ap.add_argument("--no-others", action="store_true", help="skip the other_configs leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
