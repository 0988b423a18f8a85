#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 resize path.

Metric (BASELINE.json): output Mpixel/s of an 8K (7680x4320) -> 4K (3840x2160)
lanczos3 RGBA float resize.  One "step" = one resize of one frame per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c0|c1|c2|c3|c5]
  python bench.py --impl reference ...     # the CPU algorithm, host cores

value : device-resident: src and dst live in HBM, CUDA events around the K
        steps on the launching stream, max over ranks.
e2e   : the same resize through the public API with HOST (pinned) buffers:
        H2D of the source and D2H of the result inside the timed region.
roofline : algorithmic bytes (every source pixel read once + every output
        pixel written once) / mean step time, against MEASURED_PEAKS.json.
cpu_baseline : the oracle (the reference's single-pass algorithm, row-band
        threaded like parallel_image) on a bounded row band of the same job.
Multi-GPU: frames are independent -> one frame per rank per step, no
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


def algorithmic_bytes(wl):
    sw, sh, c, sdt, dw, dh, ddt, _ = WORKLOADS[wl]
    return sw * sh * c * np.dtype(sdt).itemsize + dw * dh * c * np.dtype(ddt).itemsize


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

    def stop(self, t0=None, t1=None):
        """Median SM clock over the samples taken in [t0, t1] (host
        perf_counter times of the timed region); all samples if none fell
        inside."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        rows = self.rows
        if t0 is not None:
            inside = [r for r in rows if t0 <= r[-1] <= t1 + 0.004]
            rows = inside or rows
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "samples_total": len(self.rows)}


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
        "config": {"workload": DESCR[wl], "name": wl,
                   "note": "reference algorithm restated in oracle/ (the OIIO "
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c0", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true",
                    help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--frames-per-step", type=int, default=0,
                    help="frames in one step's batch per GPU (0 = workload default)")
    ap.add_argument("--ramp-ms", type=float, default=0.0,
                    help="extra untimed load after the warm-up steps (a long "
                         "ramp drives the GPU into its 1 kW power cap: "
                         "sustained rather than burst clocks)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import openimageio_b200 as oiio
    import synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = args.workload
    sw, sh, c, sdt, dw, dh, ddt, filt = WORKLOADS[wl]
    tdt = {"float32": torch.float32, "float16": torch.float16,
           "uint8": torch.uint8, "uint16": torch.uint16}

    # One step = one batch of `nfr` independent frames per GPU, each with its
    # own source and destination in HBM (nothing is shared or reused between
    # the frames of a batch); every rank has its own frames.
    nfr = args.frames_per_step or FRAMES_PER_STEP[wl]

    def to_torch(a):
        t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
        return t.view(torch.uint16) if a.dtype == np.uint16 else t

    src_np = synth.image(sw, sh, c, np.dtype(sdt), seed=20261019 + rank)
    src_host = to_torch(src_np).pin_memory()
    dst_host = torch.empty((dh, dw, c), dtype=tdt[ddt]).pin_memory()
    src_devs, dst_devs = [], []
    for f in range(nfr):
        t = src_host.to(dev, non_blocking=True)
        if f:  # a different picture per frame: roll the rows
            t = torch.roll(t.view(torch.uint8), shifts=137 * f, dims=0).view(t.dtype)
        src_devs.append(t)
        dst_devs.append(torch.zeros((dh, dw, c), dtype=tdt[ddt], device=dev))
    torch.cuda.synchronize()

    S_devs = [oiio.ImageBuf(t) for t in src_devs]
    D_devs = [oiio.ImageBuf(t) for t in dst_devs]
    S_host, D_host = oiio.ImageBuf(src_host), oiio.ImageBuf(dst_host)
    IBA = oiio.ImageBufAlgo
    dst_dev = dst_devs[0]

    def step_dev():
        for S, D in zip(S_devs, D_devs):
            if not IBA.resize(D, S, filtername=filt):
                raise RuntimeError(D.geterror())

    def step_host():
        # the batch through the host-memory API: every frame's source goes up
        # and every result comes down (the same pinned host frame stands in
        # for the nfr frames; a host->device copy cannot be cached)
        for _ in range(nfr):
            if not IBA.resize(D_host, S_host, filtername=filt, device=local_rank):
                raise RuntimeError(D_host.geterror())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----------------------------------------
    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    # Optional extra untimed load (--ramp-ms).  The clock sampler runs from
    # here to the end of the timed region ("under load").
    t_ramp = time.perf_counter()
    while (time.perf_counter() - t_ramp) * 1e3 < args.ramp_ms:
        step_dev()
        torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    # a ~2 ms device-side spin in front of the first event: the K launches
    # queue up behind it, so the timed region measures the device, not the
    # host's launch rate (which matters when K is small)
    t_host0 = time.perf_counter()
    torch.cuda._sleep(4_000_000)
    ev0.record()
    for _ in range(args.steps):
        step_dev()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop(t_host0, time.perf_counter()) if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- end-to-end timing (host buffers through the public API) --------
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 10))
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    rep = IBA.last_report
    h2d, d2h = int(rep.h2d_bytes) * nfr, int(rep.d2h_bytes) * nfr
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_max = float(t.item())

    # the same from PAGEABLE host memory (what an ImageBuf normally owns): the
    # library stages it through pinned bounce rings with a few host threads
    page_ms = None
    if world == 1:
        Sp = oiio.ImageBuf(src_np)
        Dp = oiio.ImageBuf(np.zeros((dh, dw, c), np.dtype(ddt)))
        for _ in range(2):
            IBA.resize(Dp, Sp, filtername=filt, device=local_rank)
        t0 = time.perf_counter()
        for _ in range(3):
            IBA.resize(Dp, Sp, filtername=filt, device=local_rank)
        page_ms = (time.perf_counter() - t0) * 1e3 / 3

    # parity spot check of what was just timed (device result == host result)
    same = bool(torch.equal(dst_dev.cpu().view(torch.uint8),
                            dst_host.view(torch.uint8))) if rank == 0 else True

    if rank == 0:
        mpix_frame = dw * dh / 1e6
        value = mpix_frame * nfr * world / (ms_max * 1e-3)
        e2e_value = mpix_frame * nfr * world / (e2e_ms_max * 1e-3)
        abytes = algorithmic_bytes(wl)
        peak, which = measured_peak()
        achieved = abytes * nfr / (ms_max * 1e-3) / 1e9  # per launch = per frame
        src_is_fp = sdt in ("float32", "float16")
        line = {
            "metric": "output Mpixel/s", "value": value, "unit": "Mpixel/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DT_SHORT[sdt], "data": "synthetic",
            "config": {"workload": DESCR[wl], "name": wl,
                       "frames_per_step": nfr * world,
                       "frames_per_step_per_gpu": nfr,
                       "l2": "source frame (%.0f MB) larger than L2; no flush"
                             % (sw * sh * c * np.dtype(sdt).itemsize / 1e6)
                       if sw * sh * c * np.dtype(sdt).itemsize > 126e6 else
                       "working set fits L2; steps run back to back",
                       "parallelism": "frame-per-gpu x%d" % world,
                       "timing": "CUDA events on the launch stream; launches "
                                 "queued behind a 2 ms device spin; %.0f ms "
                                 "untimed clock ramp after the W warm-ups"
                                 % args.ramp_ms},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_max, "steps": e2e_steps,
                    "host_memory": "pinned",
                    "pageable_value": (mpix_frame / (page_ms * 1e-3)
                                       if page_ms else None),
                    "pageable_ms_per_frame": page_ms},
            "gpu_launches": args.steps * nfr * (2 if src_is_fp else 1),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(wl), "peak_source": which,
                         "algorithmic_bytes": abytes,
                         "launch_us": ms_max * 1e3 / nfr,
                         "kernel": "resize_fused_kernel"},
            "clocks": clocks,
            "device_equals_host_result": same,
        }
        if not args.no_cpu and world == 1:
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
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
