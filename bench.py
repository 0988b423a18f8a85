#!/usr/bin/env python
"""bench.py -- benchmark of the B200 resize path, one JSON line.

Metric (BASELINE.json): output Mpixel/s of an 8K (7680x4320) -> 4K (3840x2160)
lanczos3 RGBA float resize (workload c0).  One "step" = one batch of
independent frames per GPU through the hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c0|c1|c2|c3|c5]
  python bench.py --impl reference ...     # the CPU algorithm, host cores

value      device-resident: src and dst live in HBM, CUDA events around the K
           steps on the launching stream, max over ranks (burst clocks: the
           timed region is tens of ms).
sustained  the same measurement repeated after --sustain-ms (default 300) of
           untimed load, when the B200 sits in its 1 kW power cap.
e2e        the same resize through the public API with HOST (pinned) buffers:
           H2D of the source and D2H of the result inside the timed region.
roofline   algorithmic bytes (every source pixel read once + every output
           pixel written once) / mean step time, against MEASURED_PEAKS.json.
cpu_baseline  the oracle (the reference's single-pass algorithm, row-band
           threaded like parallel_image) on a bounded row band of the same job.
configs    the other BASELINE configs, each with its own value / roofline
           fraction / e2e: c1, c2, c3, c4 (16K^2 MIP chain, box and lanczos3)
           and the real c5 batch (256 frames, sharded over the ranks).
banded     (N > 1) the 16K^2 MIP chain cut into row bands over the N GPUs of
           the box from ONE process (b2r_mipchain_create_banded: peer access,
           halo rows pushed over NVLink, per-band upload), next to the
           one-GPU chain including its upload.
Multi-GPU: frames are independent -> one batch per rank per step, no
collective in the data path (weak scaling); NCCL only for the barrier and the
max-over-ranks reduction of the timing.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (src_w, src_h, nch, src dtype, dst_w, dst_h, dst dtype, filter)
    "c0": (7680, 4320, 4, "float32", 3840, 2160, "float32", "lanczos3"),
    "c1": (1920, 1080, 3, "float32", 960, 540, "float32", "lanczos3"),
    "c2": (7680, 4320, 4, "float16", 3840, 2160, "float16", "blackman-harris"),
    "c3": (1920, 1080, 3, "uint8", 7680, 4320, "uint8", "mitchell"),
    "c5": (3840, 2160, 3, "uint16", 1920, 1080, "uint16", "lanczos3"),
}
DESCR = {
    "c0": "8K(7680x4320)->4K(3840x2160) lanczos3 RGBA float32",
    "c1": "1920x1080->960x540 lanczos3 RGB float32",
    "c2": "8K->4K blackman-harris RGBA half",
    "c3": "1920x1080->7680x4320 mitchell RGB uint8",
    "c5": "4K->1080p lanczos3 RGB uint16 (one frame of the 256-frame batch)",
}
# frames in one step's batch, per GPU (each frame has its own buffers)
FRAMES_PER_STEP = {"c0": 8, "c1": 8, "c2": 8, "c3": 8, "c5": 8}
DT_SHORT = {"float32": "f32", "float16": "f16", "uint8": "u8", "uint16": "u16"}
C5_BATCH = 256
C4_SIZE = 16384


def config_of(wl):
    """The config dict BOTH arms print (the driver compares them)."""
    return {"workload": DESCR[wl], "name": wl}


def algorithmic_bytes(wl):
    sw, sh, c, sdt, dw, dh, ddt, _ = WORKLOADS[wl]
    return sw * sh * c * np.dtype(sdt).itemsize + dw * dh * c * np.dtype(ddt).itemsize


def chain_bytes(size, nch=4):
    """Algorithmic bytes of a MIP chain from a size x size float image:
    every level read once as a source, every level >= 1 written once."""
    rd = wr = 0
    w = h = size
    while w > 1 or h > 1:
        rd += w * h * nch * 4
        w, h = max(1, w // 2), max(1, h // 2)
        wr += w * h * nch * 4
    return rd + wr, wr // (nch * 4)


def measured_traffic(wl):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel,
    per launch, from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(wl)
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,"
         "clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "5"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")]
                             + [time.perf_counter()])

    def wait_first(self, timeout=2.0):
        """Block until nvidia-smi has delivered its first sample."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.005)

    def window(self, t0, t1):
        """Median SM clock over the samples taken in [t0, t1] (host
        perf_counter times); all samples so far if none fell inside."""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        rows = list(self.rows)
        inside = [r for r in rows if t0 <= r[-1] <= t1 + 0.004]
        for r in inside or rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "samples_total": len(rows)}

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


def sample_rows(wl, cores):
    """Rows of the workload one CPU sample covers: the whole frame on a box
    with >= 16 cores (about 10-15 core-seconds for C0), proportionally fewer
    rows on smaller hosts so a sample stays around a second of wall time."""
    dh = WORKLOADS[wl][5]
    return max(8, min(dh, int(dh * min(1.0, cores / 16.0))))


def cpu_sample(wl, nthreads=0, rows=None, repeat=5):
    """Time the oracle (reference algorithm restated) on a row band of the
    workload.  Returns (Mpixel/s, cores, description)."""
    import oracle_py
    import synth
    sw, sh, c, sdt, dw, dh, ddt, filt = WORKLOADS[wl]
    cores = nthreads or (os.cpu_count() or 1)
    if rows is None:
        rows = sample_rows(wl, cores)
    src = synth.image(sw, sh, c, np.dtype(sdt), seed=20261019)
    dst = np.zeros((dh, dw, c), np.dtype(ddt))
    y0 = (dh - rows) // 2
    S, D = oracle_py.Img(src), oracle_py.Img(dst)
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        oracle_py.resize(D, S, filt, 0.0, roi=(0, dw, y0, y0 + rows),
                         nthreads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    mpix = dw * rows / best / 1e6
    return mpix, cores, ("%d of %d dst rows (full width) of %s, oracle port, "
                         "%d threads, %.2f s" % (rows, dh, wl, cores, best))


def cpu_c4_sample(size=4096, filt="lanczos3"):
    """The CPU side of the same maketx stage on a bounded sample: the oracle's
    write_mipmap level loop on a size x size RGBA float image, then every
    level cut into 64x64 half tiles and deflated (zlib level 1) on all host
    threads.  Mpixel/s over all levels, so it compares with tiles.mpixel_s."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    import oracle_py
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(20261023)
    lvl0 = rng.random((size, size, 4), dtype=np.float32)
    t0 = time.perf_counter()
    levels = [lvl0] + oracle_py.mip_chain(lvl0, filt)
    t_chain = time.perf_counter() - t0

    def encode(level):
        h, w, c = level.shape
        H2, W2 = -(-h // 64) * 64, -(-w // 64) * 64
        pad = np.zeros((H2, W2, c), np.float16)
        pad[:h, :w] = level
        t = pad.reshape(H2 // 64, 64, W2 // 64, 64, c).transpose(0, 2, 1, 3, 4)
        t = np.ascontiguousarray(t).reshape(-1, 64 * 64 * c)
        return sum(len(zlib.compress(row.tobytes(), 1)) for row in t)

    def encode_rows(job):
        level, a, b = job
        return encode(level[a:b])
    jobs = []
    for lv in levels:
        for a in range(0, lv.shape[0], 256):
            jobs.append((lv, a, min(lv.shape[0], a + 256)))
    t1 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        stored = sum(ex.map(encode_rows, jobs))
    t_enc = time.perf_counter() - t1
    npx = sum(l.shape[0] * l.shape[1] for l in levels)
    total = t_chain + t_enc
    return {"value": npx / 1e6 / total, "unit": "Mpixel/s (all levels)", "cores": cores,
            "kind": "port", "chain_ms": t_chain * 1e3, "encode_ms": t_enc * 1e3,
            "stored_mb": stored / 1e6,
            "sample": "%dx%d RGBA float %s chain (oracle port, threaded) + 64x64 half tiles, "
                      "zlib level 1 on %d threads; level 0 already in host memory"
                      % (size, size, filt, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    sw, sh, c, sdt, dw, dh, ddt, filt = WORKLOADS[wl]
    import oracle_py
    import synth
    cores = os.cpu_count() or 1
    rows = sample_rows(wl, cores)
    src = synth.image(sw, sh, c, np.dtype(sdt), seed=20261019)
    dst = np.zeros((dh, dw, c), np.dtype(ddt))
    y0 = (dh - rows) // 2
    S, D = oracle_py.Img(src), oracle_py.Img(dst)

    def step():
        oracle_py.resize(D, S, filt, 0.0, roi=(0, dw, y0, y0 + rows),
                         nthreads=cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    mpix = dw * rows / dt / 1e6
    sample = ("each step = %d of %d dst rows (full width) of the workload; "
              "Mpixel/s is per output pixel so it extrapolates to the frame"
              % (rows, dh))
    line = {
        "impl": "reference", "metric": "output Mpixel/s", "value": mpix,
        "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DT_SHORT[sdt], "data": "synthetic",
        "config": config_of(wl),
        "details": {"note": "reference algorithm restated in oracle/ (the OIIO "
                            "library itself cannot be built here: needs Imath, "
                            "OpenEXR, libtiff, OpenColorIO)"},
        "cpu_baseline": {"value": mpix, "unit": "Mpixel/s", "cores": cores,
                         "kind": "port", "sample": sample},
        "e2e": {"value": mpix, "unit": "Mpixel/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------
class Harness:
    """Everything one rank needs to time resizes on its GPU."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import openimageio_b200 as oiio
        self.torch, self.dist, self.oiio = torch, dist, oiio
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            # host-side barrier (gloo): ranks that wait while rank 0 drives all
            # GPUs from one process must not park a spinning NCCL kernel on them
            self.cpu_group = dist.new_group(backend="gloo")
        self.tdt = {"float32": torch.float32, "float16": torch.float16,
                    "uint8": torch.uint8, "uint16": torch.uint16}
        self.peak, self.peak_src = measured_peak()
        self.launches = 0

    def to_torch(self, a):
        t = self.torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
        return t.view(self.torch.uint16) if a.dtype == np.uint16 else t

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        self.torch.cuda.synchronize()
        if self.cpu_group is not None:
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, step, steps):
        """K steps between two CUDA events on the launch stream, queued behind
        a ~2 ms device spin so the region measures the device and not the
        host's launch rate.  Returns (ms per step on this rank, t0, t1)."""
        torch = self.torch
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        self.barrier()
        t0 = time.perf_counter()
        torch.cuda._sleep(4_000_000)
        ev0.record()
        for _ in range(steps):
            step()
        ev1.record()
        self.barrier()
        return ev0.elapsed_time(ev1) / steps, t0, time.perf_counter()

    def frames(self, wl, nfr, seed=20261019):
        """nfr independent device-resident (src, dst) frames of workload wl,
        plus one pinned host pair."""
        import synth
        torch = self.torch
        sw, sh, c, sdt, dw, dh, ddt, _ = WORKLOADS[wl]
        src_np = synth.image(sw, sh, c, np.dtype(sdt), seed=seed + self.rank)
        src_host = self.to_torch(src_np).pin_memory()
        dst_host = torch.empty((dh, dw, c), dtype=self.tdt[ddt]).pin_memory()
        srcs, dsts = [], []
        base = src_host.to(self.dev, non_blocking=True)
        for f in range(nfr):
            t = base
            if f:  # a different picture per frame: roll the rows
                t = torch.roll(base.view(torch.uint8), shifts=137 * f, dims=0).view(base.dtype)
            srcs.append(t)
            dsts.append(torch.zeros((dh, dw, c), dtype=self.tdt[ddt], device=self.dev))
        torch.cuda.synchronize()
        return src_np, src_host, dst_host, srcs, dsts

    def resize_config(self, wl, steps, warmup, nfr, e2e_steps, sustain_ms=0.0,
                      sampler=None, pageable=False, total_frames=None, pcie=False):
        """Device-resident + end-to-end numbers of one resize workload."""
        oiio, torch = self.oiio, self.torch
        IBA = oiio.ImageBufAlgo
        sw, sh, c, sdt, dw, dh, ddt, filt = WORKLOADS[wl]
        src_np, src_host, dst_host, srcs, dsts = self.frames(wl, nfr)
        S_devs = [oiio.ImageBuf(t) for t in srcs]
        D_devs = [oiio.ImageBuf(t) for t in dsts]
        S_host, D_host = oiio.ImageBuf(src_host), oiio.ImageBuf(dst_host)

        def step_dev():
            # one step = the batch of nfr frames through ONE call (frames of one
            # shape run as a single persistent launch, b2r_resize_batch)
            if not IBA.resize_batch(D_devs, S_devs, filtername=filt):
                raise RuntimeError(D_devs[0].geterror())

        def step_host():
            # the batch through the host-memory API: every frame's source goes
            # up and every result comes down (the same pinned host frame stands
            # in for the nfr frames; a host->device copy cannot be cached)
            for _ in range(nfr):
                if not IBA.resize(D_host, S_host, filtername=filt, device=self.local_rank):
                    raise RuntimeError(D_host.geterror())

        for _ in range(warmup):
            step_dev()
        # kernels one step launches (the library's own report of the batch call)
        per_step = nfr
        try:
            if IBA.resize_batch(D_devs, S_devs, filtername=filt, report=True):
                # 1 for a batch launch, the sum over the frames otherwise
                per_step = max(1, int(IBA.last_report.kernels_launched))
        except Exception:
            pass
        ms, t0, t1 = self.time_steps(step_dev, steps)
        ms = self.max_over_ranks(ms)
        clocks = sampler.window(t0, t1) if sampler else None
        out = {"ms_per_step": ms, "clocks": clocks}
        if sustain_ms > 0:
            t_ramp = time.perf_counter()
            while (time.perf_counter() - t_ramp) * 1e3 < sustain_ms:
                step_dev()
                torch.cuda.synchronize()
            ms_s, t0, t1 = self.time_steps(step_dev, steps)
            out["sustained_ms_per_step"] = self.max_over_ranks(ms_s)
            out["sustained_clocks"] = sampler.window(t0, t1) if sampler else None
            # the same power state's copy roofline: MEASURED_PEAKS.json's recipe
            # (b.copy_(a), read + write bytes) timed right behind the load
            try:
                n = 1 << 29
                a = torch.empty(n, dtype=torch.bfloat16, device=self.dev)
                b = torch.empty_like(a)
                t_ramp = time.perf_counter()
                while (time.perf_counter() - t_ramp) * 1e3 < sustain_ms:
                    b.copy_(a)
                    torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    b.copy_(a)
                e1.record()
                torch.cuda.synchronize()
                out["sustained_copy_gbs"] = 2 * n * 2 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
                del a, b
            except Exception:
                out["sustained_copy_gbs"] = None
        self.launches += (warmup + steps) * per_step
        out["launches_timed"] = steps * per_step  # kernels inside the timed region

        # end to end (host buffers through the public API)
        for _ in range(2):
            step_host()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        self.barrier()
        e2e_ms = self.max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
        rep = IBA.last_report
        out["e2e_ms_per_step"] = e2e_ms
        if pcie:
            # what this host can feed the GPU(s): plain pinned -> device copies of
            # the same source frame, all ranks at once, no kernels
            flat = src_host.view(torch.uint8).reshape(-1)
            dflat = torch.empty_like(flat, device=self.dev)
            for _ in range(2):
                dflat.copy_(flat, non_blocking=True)
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(5):
                dflat.copy_(flat, non_blocking=True)
            self.barrier()
            cp_ms = self.max_over_ranks((time.perf_counter() - t0) * 1e3 / 5)
            out["h2d_ceiling_gbs"] = flat.numel() / (cp_ms * 1e-3) / 1e9
            del dflat
        out["h2d"] = int(rep.h2d_bytes) * nfr
        out["d2h"] = int(rep.d2h_bytes) * nfr
        out["kernel_path"] = ("stream" if getattr(rep, "kernels_launched", 2) == 1
                              and sdt in ("float32", "float16") else "fused")
        if pageable and self.world == 1:
            Sp = oiio.ImageBuf(src_np)
            Dp = oiio.ImageBuf(np.zeros((dh, dw, c), np.dtype(ddt)))
            for _ in range(2):
                IBA.resize(Dp, Sp, filtername=filt, device=self.local_rank)
            t0 = time.perf_counter()
            for _ in range(3):
                IBA.resize(Dp, Sp, filtername=filt, device=self.local_rank)
            out["pageable_ms"] = (time.perf_counter() - t0) * 1e3 / 3
        # parity spot check of what was just timed (device result == host result)
        out["same"] = bool(torch.equal(dsts[0].cpu().view(torch.uint8),
                                       dst_host.view(torch.uint8)))
        out["nfr"] = nfr
        del S_devs, D_devs, srcs, dsts
        torch.cuda.empty_cache()
        return out

    def summarise(self, wl, r, frames_total):
        """value / roofline / e2e record of a resize workload from
        resize_config's timings; frames_total = frames per step, all ranks."""
        sw, sh, c, sdt, dw, dh, ddt, _ = WORKLOADS[wl]
        mpix = dw * dh / 1e6
        ab = algorithmic_bytes(wl)
        per_gpu = frames_total / self.world
        ach = ab * per_gpu / (r["ms_per_step"] * 1e-3) / 1e9
        rec = {"workload": DESCR[wl],
               "value": mpix * frames_total / (r["ms_per_step"] * 1e-3),
               "unit": "Mpixel/s", "launch_us": r["ms_per_step"] * 1e3 / per_gpu,
               "achieved_gbs": ach, "frac": ach / self.peak,
               "e2e_value": mpix * frames_total / (r["e2e_ms_per_step"] * 1e-3),
               "device_equals_host_result": r["same"]}
        if "sustained_ms_per_step" in r:
            a2 = ab * per_gpu / (r["sustained_ms_per_step"] * 1e-3) / 1e9
            rec["sustained_value"] = mpix * frames_total / (r["sustained_ms_per_step"] * 1e-3)
            rec["sustained_frac"] = a2 / self.peak
        return rec

    # ---- C5: the real 256-frame batch, sharded over the ranks -----------
    def c5_batch(self, steps=3):
        oiio, torch = self.oiio, self.torch
        IBA = oiio.ImageBufAlgo
        wl = "c5"
        sw, sh, c, sdt, dw, dh, ddt, filt = WORKLOADS[wl]
        mine = oiio.shard_frames(C5_BATCH, self.world, self.rank)
        nf = len(mine)
        _, src_host, dst_host, srcs, dsts = self.frames(wl, nf, seed=20261024)
        S = [oiio.ImageBuf(t) for t in srcs]
        D = [oiio.ImageBuf(t) for t in dsts]
        S_host, D_host = oiio.ImageBuf(src_host), oiio.ImageBuf(dst_host)

        def step_dev():
            if not IBA.resize_batch(D, S, filtername=filt):
                raise RuntimeError(D[0].geterror())

        def step_host():
            for _ in range(nf):
                if not IBA.resize(D_host, S_host, filtername=filt, device=self.local_rank):
                    raise RuntimeError(D_host.geterror())
        step_dev()
        ms, _, _ = self.time_steps(step_dev, steps)
        ms = self.max_over_ranks(ms)
        self.launches += (1 + steps)
        step_host()
        self.barrier()
        t0 = time.perf_counter()
        step_host()
        self.barrier()
        e2e_ms = self.max_over_ranks((time.perf_counter() - t0) * 1e3)
        same = bool(torch.equal(dsts[0].cpu().view(torch.uint8), dst_host.view(torch.uint8)))
        del S, D, srcs, dsts
        torch.cuda.empty_cache()
        mpix = dw * dh / 1e6 * C5_BATCH
        ach = algorithmic_bytes(wl) * nf / (ms * 1e-3) / 1e9  # per GPU
        return {"workload": "batch of %d 4K->1080p lanczos3 RGB uint16 frames, "
                            "frames sharded over %d GPU(s)" % (C5_BATCH, self.world),
                "frames": C5_BATCH, "frames_per_gpu": nf, "scaling": "strong",
                "value": mpix / (ms * 1e-3), "unit": "Mpixel/s", "ms_per_batch": ms,
                "launch_us": ms * 1e3 / nf, "achieved_gbs": ach, "frac": ach / self.peak,
                "e2e_value": mpix / (e2e_ms * 1e-3), "e2e_ms_per_batch": e2e_ms,
                "device_equals_host_result": same}

    # ---- C4: 16K^2 RGBA float MIP chain on one GPU -----------------------
    def c4_level0(self, size):
        """A synthetic size x size RGBA float level 0 in HBM (counter hash of
        the element index, computed on the device) and its pinned host copy."""
        torch = self.torch
        n = size * size * 4
        t = torch.empty(n, dtype=torch.float32, device=self.dev)
        chunk = 1 << 26
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            i = torch.arange(a, b, device=self.dev, dtype=torch.int64)
            v = (i * 2654435761 + 20261023) & 0xFFFFFFFF
            v = ((v ^ (v >> 15)) * 2246822519) & 0xFFFFFFFF
            v = v ^ (v >> 13)
            t[a:b] = v.to(torch.float32) * (1.0 / 4294967296.0)
            del i, v
        return t.view(size, size, 4)

    def c4_chain(self, size, host_pinned, lvl0, filt, repeats=2, with_tiles=False):
        oiio, torch = self.oiio, self.torch
        ab, out_px = chain_bytes(size)
        best = None
        nlev = 0
        for _ in range(repeats):
            ch = oiio.MipChain(oiio.ImageBuf(lvl0), filtername=filt, device=self.local_rank)
            nlev = ch.nlevels
            best = ch.kernel_ms if best is None else min(best, ch.kernel_ms)
            del ch
        self.launches += repeats * (nlev - 1)
        # end to end: pinned host level 0 up, chain, every level >= 1 back
        t0 = time.perf_counter()
        ch = oiio.MipChain(oiio.ImageBuf(host_pinned), filtername=filt, device=self.local_rank)
        t_up = time.perf_counter()
        total = 0
        outs = []
        for lv in range(1, ch.nlevels):
            outs.append(ch.download(lv))
            total += outs[-1].nbytes
        e2e_ms = (time.perf_counter() - t0) * 1e3
        up_ms = (t_up - t0) * 1e3
        del ch
        torch.cuda.empty_cache()
        # the same as ONE pipeline (b2r_mipchain_create_streamed): level 1 comes down
        # band by band while level 0 is still going up; into the same (now touched) arrays
        streamed_ms = None
        try:
            best_s = None
            for _ in range(2):
                t0 = time.perf_counter()
                ch, _lv = oiio.MipChain.streamed(oiio.ImageBuf(host_pinned), filt,
                                                 device=self.local_rank, outputs=outs)
                dt = (time.perf_counter() - t0) * 1e3
                best_s = dt if best_s is None else min(best_s, dt)
                del ch
            streamed_ms = best_s
        except Exception as e:  # a side measurement
            streamed_ms = "failed: %s" % e
        del outs
        torch.cuda.empty_cache()
        # maketx's output stage (maketexture.cpp:957-974): level 0 up, chain,
        # EVERY level re-laid into 64x64 tiles (float -> half, maketx -d half) on
        # the device, brought down on a copy stream and deflated by host threads
        tiles = None
        if with_tiles:
            t0 = time.perf_counter()
            ch = oiio.MipChain(oiio.ImageBuf(host_pinned), filtername=filt, device=self.local_rank)
            ts = oiio.TileSet(ch, tile=64, half=True, zlevel=1, nthreads=0)
            t_ms = (time.perf_counter() - t0) * 1e3
            npx = sum(ts.level_info(l)[0] * ts.level_info(l)[1] for l in range(ts.nlevels))
            tiles = {"what": "pinned level 0 up, chain, all %d levels -> 64x64 half tiles, "
                             "zlib level 1 on host threads; wall clock" % ts.nlevels,
                     "e2e_ms": t_ms, "mpixel_s_all_levels": npx / 1e6 / (t_ms * 1e-3),
                     "encode_ms": ts.stat("total_ms"), "device_ms": ts.stat("device_ms"),
                     "deflate_busy_ms": ts.stat("deflate_ms"),
                     "setup_ms": ts.stat("setup_ms"), "download_wait_ms": ts.stat("wait_ms"),
                     "raw_mb": ts.stat("raw_bytes") / 1e6, "stored_mb": ts.stat("stored_bytes") / 1e6,
                     "host_threads": os.cpu_count()}
            del ts, ch
            torch.cuda.empty_cache()
        ach = ab / (best * 1e-3) / 1e9
        return {"tiles": tiles,
                "workload": "maketx MIP chain of a %dx%d RGBA float32 texture, %s, "
                            "%d levels, device-resident" % (size, size, filt, nlev),
                "value": out_px / 1e6 / (best * 1e-3), "unit": "Mpixel/s (levels >= 1)",
                "chain_ms": best, "achieved_gbs": ach, "frac": ach / self.peak,
                "algorithmic_bytes": ab,
                "frac_note": ("algorithmic bytes = every level read once and written once; "
                              "peak = the measured COPY bandwidth (half reads, half writes). "
                              "The chain is 3/4 reads and its levels <= 67 MB turn around "
                              "inside the 126 MB L2, so frac can pass 1.0"),
                "e2e_value": out_px / 1e6 / (e2e_ms * 1e-3), "e2e_ms": e2e_ms,
                "e2e_upload_and_chain_ms": up_ms,
                "e2e_streamed_ms": streamed_ms,
                "e2e_streamed_value": (out_px / 1e6 / (streamed_ms * 1e-3)
                                       if isinstance(streamed_ms, float) else None),
                "e2e_note": "level 0 (%.1f GB, pinned host) up, chain, levels >= 1 "
                            "(%.1f GB) down to pageable host memory"
                            % (size * size * 16 / 1e9, total / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c0", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true",
                    help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true",
                    help="headline workload only (skip c1..c5 / c4 / banded)")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--frames-per-step", type=int, default=0,
                    help="frames in one step's batch per GPU (0 = workload default)")
    ap.add_argument("--ramp-ms", type=float, default=0.0,
                    help="extra untimed load BEFORE the headline measurement")
    ap.add_argument("--sustain-ms", type=float, default=300.0,
                    help="untimed load between the burst measurement and its "
                         "sustained twin (drives the GPU into its 1 kW power "
                         "cap); 0 = skip the twin")
    ap.add_argument("--c4-size", type=int, default=C4_SIZE)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    H = Harness(args)
    torch = H.torch
    wl = args.workload
    sw, sh, c, sdt, dw, dh, ddt, filt = WORKLOADS[wl]
    nfr = args.frames_per_step or FRAMES_PER_STEP[wl]
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 10))

    sampler = ClockSampler(H.local_rank) if H.rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first()
    if args.ramp_ms > 0:
        x = torch.empty(1 << 28, dtype=torch.float32, device=H.dev)
        t_ramp = time.perf_counter()
        while (time.perf_counter() - t_ramp) * 1e3 < args.ramp_ms:
            x.add_(1.0)
            torch.cuda.synchronize()
        del x
    r = H.resize_config(wl, args.steps, args.warmup, nfr, e2e_steps,
                        sustain_ms=args.sustain_ms, sampler=sampler, pageable=True,
                        pcie=True)
    head = H.summarise(wl, r, nfr * H.world)

    configs, banded = {}, None
    if not args.no_configs:
        for w2 in ("c1", "c2", "c3"):
            if w2 == wl or H.world > 1:
                continue  # single-GPU shapes: measured at N = 1
            rr = H.resize_config(w2, 10, 3, FRAMES_PER_STEP[w2], 3)
            configs[w2] = H.summarise(w2, rr, FRAMES_PER_STEP[w2])
        configs["c5"] = H.c5_batch()
        if H.rank == 0 and H.world == 1:
            try:
                size = args.c4_size
                lvl0 = H.c4_level0(size)
                host = torch.empty((size, size, 4), dtype=torch.float32).pin_memory()
                host.copy_(lvl0)
                torch.cuda.synchronize()
                for f in ("box", "lanczos3"):
                    configs["c4_" + f] = H.c4_chain(size, host, lvl0, f,
                                                    with_tiles=(f == "lanczos3"))
                if not args.no_cpu:
                    configs["c4_lanczos3"]["cpu_baseline"] = cpu_c4_sample()
                del lvl0, host
                torch.cuda.empty_cache()
            except Exception as e:  # a side config must not sink the headline
                configs["c4_error"] = str(e)
        if H.world > 1:
            H.barrier()
            H.cpu_barrier()
            if H.rank == 0:
                try:
                    import banded_bench
                    banded = banded_bench.run(H, args.c4_size)
                except Exception as e:
                    banded = {"error": str(e)}
            H.cpu_barrier()
            H.barrier()
    if sampler:
        sampler.stop()

    if H.rank == 0:
        mpix_frame = dw * dh / 1e6
        abytes = algorithmic_bytes(wl)
        src_mb = sw * sh * c * np.dtype(sdt).itemsize / 1e6
        line = {
            "metric": "output Mpixel/s", "value": head["value"], "unit": "Mpixel/s",
            "n_gpus": H.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DT_SHORT[sdt],
            "data": "synthetic",
            "config": config_of(wl),
            "details": {
                "frames_per_step": nfr * H.world, "frames_per_step_per_gpu": nfr,
                "l2": ("source frame (%.0f MB) larger than L2; no flush" % src_mb)
                      if src_mb > 126 else
                      ("%d independent frames per step, %.0f MB of sources per "
                       "step (> L2), run back to back" % (nfr, src_mb * nfr)),
                "parallelism": "frame-per-gpu x%d" % H.world,
                "timing": "CUDA events on the launch stream; launches queued "
                          "behind a 2 ms device spin; burst = right after the W "
                          "warm-ups, sustained = after %.0f ms of untimed load"
                          % args.sustain_ms},
            "e2e": {"value": head["e2e_value"], "unit": "Mpixel/s",
                    "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "ms_per_step": r["e2e_ms_per_step"], "steps": e2e_steps,
                    "host_memory": "pinned",
                    # per-GPU upload rate of the e2e step against a plain
                    # cudaMemcpyAsync loop of the same frame run by all ranks at
                    # once (the host's ceiling at this N)
                    "h2d_gbs_per_gpu": r["h2d"] / (r["e2e_ms_per_step"] * 1e-3) / 1e9,
                    "h2d_ceiling_gbs_per_gpu": r.get("h2d_ceiling_gbs"),
                    "pcie_frac": (r["h2d"] / (r["e2e_ms_per_step"] * 1e-3) / 1e9
                                  / r["h2d_ceiling_gbs"]) if r.get("h2d_ceiling_gbs") else None,
                    "pageable_value": (mpix_frame / (r["pageable_ms"] * 1e-3)
                                       if r.get("pageable_ms") else None),
                    "pageable_ms_per_frame": r.get("pageable_ms")},
            # kernels of this library launched inside the headline's timed region
            # (one persistent launch per step's batch of frames)
            "gpu_launches": r.get("launches_timed", args.steps),
            "gpu_launches_whole_run": H.launches,
            "roofline": {"bound": "hbm", "achieved": head["achieved_gbs"],
                         "peak": H.peak, "unit": "GB/s", "frac": head["frac"],
                         "traffic": measured_traffic(wl),
                         "traffic_source": "static: dram__bytes_read+write of the "
                                           "committed ncu capture (profiles/traffic.json), "
                                           "not measured in this run",
                         "peak_source": H.peak_src, "algorithmic_bytes": abytes,
                         "launch_us": head["launch_us"],
                         "frac_of_nominal_8tbs": head["achieved_gbs"] / 8000.0,
                         "kernel": "resize_stream_kernel (TMA, warp-specialised)"},
            "clocks": r["clocks"],
            "device_equals_host_result": r["same"],
        }
        if "sustained_value" in head:
            s_ms = r["sustained_ms_per_step"]
            line["sustained"] = {
                "value": head["sustained_value"], "unit": "Mpixel/s",
                "ms_per_step": s_ms, "launch_us": s_ms * 1e3 / nfr,
                "frac": head["sustained_frac"],
                "frac_of_nominal_8tbs": head["sustained_frac"] * H.peak / 8000.0,
                "after_ms_of_load": args.sustain_ms,
                # a plain device copy timed in the same power state (after the same
                # load): the sustained twin of MEASURED_PEAKS.json's burst figure
                "copy_gbs_same_state": r.get("sustained_copy_gbs"),
                "frac_of_copy_same_state": (head["sustained_frac"] * H.peak / r["sustained_copy_gbs"]
                                            if r.get("sustained_copy_gbs") else None),
                "clocks": r.get("sustained_clocks")}
        if configs:
            line["configs"] = configs
        if banded is not None:
            line["banded"] = banded
        if not args.no_cpu and H.world == 1:
            try:
                mp, cores, sample = cpu_sample(wl)
                line["cpu_baseline"] = {"value": mp, "unit": "Mpixel/s",
                                        "cores": cores, "kind": "port",
                                        "sample": sample}
            except Exception as e:  # the baseline must not sink the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "Mpixel/s",
                                        "cores": 0, "kind": "port",
                                        "sample": "failed: %s" % e}
        print(json.dumps(line))
    if H.world > 1:
        H.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
