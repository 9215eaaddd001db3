#!/usr/bin/env python
"""bench.py -- throughput of the PyGOFPID foreground-extraction hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Headline (BASELINE.json configs[1]): one 1080p camera stream per GPU through the whole hot path --
Gaussian blur -> MOG2 -> open/close -> connected-component blobs.  A "step" is PASSES passes over a
72-frame clip (864 frames), so the default 20 steps time ~0.5 s of device work.  N > 1 is launched
by torchrun, one rank per GPU, each rank owning its own independent stream (weak scaling, no
data-path collective; torch.distributed carries the barrier, the max-over-ranks and the final host
gather of per-stream detections only).

Output: ONE JSON line (see the task contract):
  value        frames/s with the clip resident in HBM (gofpid_pipeline_process_device_sequence)
  e2e          the same through the C-ABI host entry point with PINNED HOST frames: every frame is copied
               host->device and every blob list device->host inside the timed region
  roofline     the dominant kernel (fused blur+MOG2): algorithmic bytes / CUDA-event time / measured peak
  cpu_baseline the reference's own OpenCV calls (oracle/cv_path.py) timed on this box's host cores
  parity_checked  before anything is timed, the SAME entry points are fed a short clip and every frame's
               blob list (and the last mask) is compared with the reference's cv2 calls; a mismatch aborts
  dropin       frames/s through the drop-in class pygofpid_b200.GOFPID (detect / detect_clip)
  configs      the other four BASELINE configs + the 256-stream 720x576 config sharded over the ranks
               (streams_for_rank), each with value / e2e / roofline / cpu_baseline
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLIP_FRAMES = 72  # distinct frames cycled per stream: 72 x 2.07 MB = 149 MB > 126 MB of L2
PASSES = 12       # clip passes per step (headline): 864 frames
TOTAL_STREAMS = 256  # BASELINE configs[3]


def rect(n):
    return np.ones((n, n), np.uint8)


def ellipse15():
    # cv.getStructuringElement(MORPH_ELLIPSE, (15, 15)) of OpenCV 4.13 (row widths, centred)
    widths = [1, 9, 11, 13, 13, 15, 15, 15, 15, 15, 13, 13, 11, 9, 1]
    k = np.zeros((15, 15), np.uint8)
    for r, w in enumerate(widths):
        k[r, (15 - w) // 2:(15 - w) // 2 + w] = 1
    return k


# clip: distinct frames per stream (chosen so that one pass of a rank's streams exceeds the 126 MB L2)
WORKLOADS = {
    # BASELINE.json configs[1] -- the headline
    "1080p_mog2_open_close_blobs": dict(H=1080, W=1920, C=1, detector="MOG2", n_obj=12, seed=2, clip=72, passes=PASSES,
                                        morph=[("erode", 5), ("dilate", 5), ("dilate", 5), ("erode", 5)]),
    "1080p_bgr_mog2_open_close_blobs": dict(H=1080, W=1920, C=3, detector="MOG2", n_obj=12, seed=2, clip=24, passes=32,
                                            morph=[("erode", 5), ("dilate", 5), ("dilate", 5), ("erode", 5)]),
    # configs[0]
    "320x240_fd_open_blobs": dict(H=240, W=320, C=1, detector="FD", n_obj=6, seed=1, clip=1728, passes=2,
                                  morph=[("erode", 3), ("dilate", 3)]),
    # configs[2]
    "4k_mog2_ellipse15_open_blobs": dict(H=2160, W=3840, C=1, detector="MOG2", n_obj=16, seed=3, clip=18, passes=24,
                                         morph=[("erode", "e15"), ("dilate", "e15")]),
    # configs[3]: 256 independent streams in total, sharded over the ranks (or --streams per GPU)
    "ilids_720x576_bgr_defaults": dict(H=576, W=720, C=3, detector="MOG2", n_obj=6, seed=1000, clip=8, passes=4,
                                       morph=[("erode", 5), ("dilate", 5)], total_streams=TOTAL_STREAMS),
    # configs[4]
    "1080p_knn_many_blobs": dict(H=1080, W=1920, C=1, detector="KNN", n_obj=400, seed=5, clip=72, passes=6,
                                 obj_size=(12, 90, 8, 80), morph=[("erode", 5), ("dilate", 5)]),
    # (not a BASELINE configuration: KNN on the headline scene -- a dozen objects over a static background)
    "1080p_knn_few_blobs": dict(H=1080, W=1920, C=1, detector="KNN", n_obj=12, seed=2, clip=72, passes=6,
                                morph=[("erode", 5), ("dilate", 5)]),
    # (not a BASELINE configuration: the colour KNN kernel for profiles/)
    "1080p_bgr_knn": dict(H=1080, W=1920, C=3, detector="KNN", n_obj=12, seed=6, clip=24, passes=4,
                          morph=[("erode", 5), ("dilate", 5)]),
}
DEFAULT_WORKLOAD = "1080p_mog2_open_close_blobs"
EXTRA_CONFIGS = ["320x240_fd_open_blobs", "1080p_bgr_mog2_open_close_blobs", "4k_mog2_ellipse15_open_blobs",
                 "1080p_knn_many_blobs", "ilids_720x576_bgr_defaults"]
BASE_STREAMS = 16  # distinct synthetic streams generated for the 256-stream config (the rest are time-shifted copies)


def morph_list(spec):
    return [(op, ellipse15() if k == "e15" else rect(k)) for op, k in spec]


def make_clip(wl, n_frames, stream=0):
    """(n_frames, H, W[, C]) of camera `stream` of the workload (its own seed)."""
    from pygofpid_b200.synth import SyntheticClip
    clip = SyntheticClip(wl["H"], wl["W"], wl["C"], seed=wl["seed"] + stream, n_obj=wl["n_obj"],
                         noise_sigma=4.0, obj_size=wl.get("obj_size"), rng="pcg")
    with ThreadPoolExecutor(min(8, os.cpu_count() or 1)) as ex:
        return np.stack(list(ex.map(clip.frame_at, range(n_frames))))


def make_streams(wl, n_frames, streams):
    """(n_frames, len(streams), H, W[, C]): the clips of the given GLOBAL stream indices.  At most
    BASE_STREAMS distinct cameras are synthesised; stream i beyond that replays camera i % BASE_STREAMS
    shifted by i // BASE_STREAMS frames (independent models all the same: no two streams see the same
    frame at the same time)."""
    base = {}
    out = []
    for i in streams:
        b = i % BASE_STREAMS
        if b not in base:
            base[b] = make_clip(wl, n_frames, stream=b)
        out.append(np.roll(base[b], -(i // BASE_STREAMS), axis=0))
    return np.stack(out, axis=1)


def config_of(name, wl, S, world):
    """The `config` object: identical for both arms of the same workload."""
    F = wl["clip"] * wl["passes"]
    clip_mb = wl["clip"] * S * wl["H"] * wl["W"] * wl["C"] / 1e6
    l2 = ("inputs larger than L2: every pass streams a %.0f MB clip per GPU (L2 is 126 MB), no flush needed" % clip_mb
          if clip_mb > 126 else
          "the %.1f MB clip of this frame size stays in the 126 MB L2 (no flush: frames this small are L2-resident)" % clip_mb)
    return {"workload": name, "l2": l2, "shape": [wl["H"], wl["W"], wl["C"]], "detector": wl["detector"],
            "blur": "gaussian3x3_sigma4", "morph": ["%s%s" % (op, k) for op, k in wl["morph"]], "area_min": 100,
            "streams_per_gpu": S, "clip_frames": wl["clip"], "passes_per_step": wl["passes"], "frames_per_step": F * S}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        return len(self.rows)

    def stop(self, lo=0, hi=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[lo:hi] or self.rows
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm
def cv_hot_path(wl):
    """The reference's own OpenCV CPU path (oracle/cv_path.py issues the cv2 calls of GOFPID.detect
    pipeline.py:461-479 + _find_blob :555-588) for a workload, on all host cores."""
    import cv2 as cv
    from oracle.cv_path import CvHotPath
    cv.setNumThreads(os.cpu_count() or 1)
    cv.setRNGSeed(12345)
    se = {5: cv.getStructuringElement(cv.MORPH_RECT, (5, 5)), 3: cv.getStructuringElement(cv.MORPH_RECT, (3, 3)),
          "e15": cv.getStructuringElement(cv.MORPH_ELLIPSE, (15, 15))}
    mm = [dict(fun=cv.erode if op == "erode" else cv.dilate, kernel=se[k]) for op, k in wl["morph"]]
    return CvHotPath(blur="default", frg_detect=wl["detector"], mat_morph=mm), cv


def cpu_reference_run(wl, frames, n_frames=None, seconds=None, warm=4):
    """Times path.step over `n_frames` frames (or for about `seconds`) cycling through `frames` (T, H, W[, C])."""
    path, cv = cv_hot_path(wl)
    T = len(frames)
    i = 0
    for _ in range(warm):
        path.step(frames[i % T]); i += 1
    t0 = time.perf_counter()
    n = nb = 0
    while True:
        _, blobs = path.step(frames[i % T]); i += 1
        n += 1
        nb += len(blobs)
        if n_frames is not None and n >= n_frames:
            break
        if n_frames is None and n >= 8 and time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    return dict(fps=n / dt, seconds=dt, frames=n, threads=cv.getNumThreads(), cores=os.cpu_count(),
                cv=cv.__version__, blobs_per_frame=nb / max(n, 1))


def cpu_baseline_obj(r, what):
    return {"value": r["fps"], "unit": "frames/s", "cores": r["threads"], "kind": "port",
            "sample": "%d frames (%.1f s) of %s through the reference's cv2 %s calls (oracle/cv_path.py: blur, apply, "
                      "erode/dilate, findContours + contourArea + moments), %d cv2 threads on %d host cores"
                      % (r["frames"], r["seconds"], what, r["cv"], r["threads"], r["cores"])}


def run_reference_arm(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S = args.streams
    cfg = config_of(name, wl, S, 1)
    frames = make_clip(wl, wl["clip"])
    F = cfg["frames_per_step"]
    # the CPU path runs the streams one after the other (one GOFPID object per camera, dataset_ilids.py:98-118):
    # stream-frames per second is what one stream's loop delivers
    r0 = cpu_reference_run(wl, frames, n_frames=max(1, min(args.warmup, 3)) * 8, warm=0)
    r = cpu_reference_run(wl, frames, n_frames=args.steps * F, warm=0)
    mpix = wl["H"] * wl["W"] / 1e6
    line = {
        "impl": "reference", "metric": "frames_per_sec", "value": r["fps"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
        "mpix_per_s": r["fps"] * mpix, "config": cfg,
        "details": {"cv2": r["cv"], "warmup_frames": r0["frames"],
                    "note": "reference's OpenCV CPU hot path (blur, apply, erode/dilate, findContours+area+moments), "
                            "same synthetic clip, same frames per step as the GPU arm"},
        "cpu_baseline": cpu_baseline_obj(r, "the same synthetic clip"),
        "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_per_launch(kernel, workload, streams, frames_per_launch):
    """dram bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json: {workload: {kernel, streams, frames_per_launch, bytes}}); None otherwise."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p)).get(workload)
        if d and d.get("kernel") == kernel and d.get("streams") == streams and d.get("frames_per_launch") == frames_per_launch:
            return d.get("bytes")
    except Exception:
        pass
    return None


def new_hot_path(wl, S, device, **kw):
    from pygofpid_b200 import HotPath
    H, W, C = wl["H"], wl["W"], wl["C"]
    return HotPath((H, W) if C == 1 else (H, W, C), blur=("gaussian", (3, 3), 4.0), detector=wl["detector"],
                   morph=morph_list(wl["morph"]), area_min=100.0, n_streams=S, device=device, knn_seed=12345, **kw)


def parity_check(name, wl, device):
    """The timed entry points against the reference's cv2 calls, on a fresh handle, before anything is
    timed: a clip long enough to reach the grouped MOG2 launches (depth eager frames + two groups) goes
    through gofpid_pipeline_process_device_sequence AND gofpid_pipeline_process_host_sequence; every
    frame's blob list (contour points, order, centre, bottom) and the last frame's mask must be identical
    to cv2's.  Multi-stream workloads are checked with two streams per handle.  Raises on mismatch."""
    import torch
    from pygofpid_b200.hotpath import blobs_from_records
    S = 2 if wl.get("total_streams") else 1
    report = {}
    for via in ("device", "host"):
        hp = new_hot_path(wl, S, device)
        n = hp.depth + 2 * hp.group
        clips = [make_clip(wl, n, stream=s) for s in range(S)]
        want = []
        for s in range(S):
            path, _ = cv_hot_path(wl)
            want.append([path.step(f) for f in clips[s]])
        arr = np.stack(clips, axis=1)
        src = torch.from_numpy(arr).cuda(device) if via == "device" else torch.from_numpy(arr).pin_memory()
        counts, recs, pts = hp.process_sequence(src)
        ob = op = nblobs = 0
        for t in range(n):
            for s in range(S):
                nb, npt = int(counts[t, s, 0]), int(counts[t, s, 1])
                got = blobs_from_records(recs[ob:ob + nb], pts[op:op + npt])
                ob += nb; op += npt
                wb = want[s][t][1]
                ok = len(got) == len(wb) and all(
                    np.array_equal(g["contour"], w["contour"]) and np.array_equal(g["center"], w["center"]) and
                    np.array_equal(g["bottom"], w["bottom"]) for g, w in zip(got, wb))
                if not ok:
                    raise SystemExit("bench.py: PARITY FAILURE on %s via %s clip entry point: frame %d stream %d: blob list "
                                     "differs from cv2 (%d vs %d blobs)" % (name, via, t, s, len(got), len(wb)))
                nblobs += nb
        for s in range(S):
            if not np.array_equal(hp.mask(s), want[s][n - 1][0]):
                raise SystemExit("bench.py: PARITY FAILURE on %s via %s clip entry point: last mask differs from cv2" % (name, via))
        report[via] = {"frames": n * S, "blobs": nblobs, "frames_per_detector_launch": hp.group}
        hp.close()
        del src
    return report


def measure(name, wl, S, streams, steps, warmup, local, dist, world, e2e_per_frame=False, sampler=None):
    """One workload on this rank: `S` camera streams (global indices `streams`) on GPU `local`.
    Returns a dict of raw measurements (times already max-reduced over the ranks)."""
    import torch
    H, W, C = wl["H"], wl["W"], wl["C"]
    n_clip, passes = wl["clip"], wl["passes"]
    hp = new_hot_path(wl, S, local)
    host = make_streams(wl, n_clip, streams)  # (T, S, H, W[, C])
    pinned = torch.from_numpy(host).pin_memory()
    dev = pinned.cuda(local, non_blocking=False)
    row = W * C
    frame_bytes = row * H
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def device_step():
        for _ in range(passes):
            hp.enqueue_device_sequence(dev.data_ptr(), n_clip, S * frame_bytes, row, frame_bytes, sh)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        device_step()
    hp.frame_info()  # first result fetch outside the timed region
    barrier()
    mark0 = sampler.mark() if sampler else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = hp.launch_total()
    e0.record(stream)
    for _ in range(steps):
        device_step()
    hp.join(sh)  # morphology / blob kernels of the last frames run on the handle's own streams
    e1.record(stream)
    barrier()
    mark1 = sampler.mark() if sampler else 0
    dev_ms = e0.elapsed_time(e1)
    launches = hp.launch_total() - l0
    info = hp.frame_info()
    F = n_clip * passes  # time steps per step

    # ---- end to end through the clip entry point: pinned host frames in, blob lists out ----------
    # One library call per host clip.  A call fills and drains the frame pipeline once, so the clip handed over
    # is the 72-frame clip several times over (what a caller decoding a file in chunks would pass), as many
    # passes per call as fit in ~1.2 GB of pinned memory.
    e2e_steps = max(1, steps // 4)
    per_call = max(1, min(passes, int(1.2e9 // max(pinned.numel(), 1))))
    while passes % per_call:
        per_call -= 1
    long_clip = pinned if per_call == 1 else torch.cat([pinned] * per_call).pin_memory()
    hp.process_sequence(pinned[:min(n_clip, hp.depth + hp.group)])  # untimed: staging buffers, output sizing
    barrier()
    w0 = time.perf_counter()
    seq_d2h = 0
    nblobs = 0
    for _ in range(e2e_steps * passes // per_call):
        counts, recs, pts = hp.process_sequence(long_clip)
        seq_d2h += 32 * n_clip * per_call * S + len(recs) * 64 + pts.shape[0] * 8
        nblobs += len(recs)
    torch.cuda.synchronize()
    seq_s = time.perf_counter() - w0
    seq_frames = e2e_steps * F
    del long_clip

    # ---- the host link alone, same buffers, all ranks at once: what end to end cannot beat ----------
    # (copies of one detector group of frames, as the clip entry point issues them)
    gfr = max(1, min(hp.group, n_clip))
    stage = torch.empty((gfr,) + tuple(pinned.shape[1:]), dtype=torch.uint8, device="cuda")
    reps = max(4, min(64, int(2e8 // max(stage.numel(), 1))))
    for i in range(2):
        stage.copy_(pinned[:gfr], non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    for i in range(reps):
        k = (i * gfr) % (n_clip - gfr + 1)
        stage.copy_(pinned[k:k + gfr], non_blocking=True)
    c1.record(stream)
    torch.cuda.synchronize()
    link_ms = c0.elapsed_time(c1)
    link_frames = reps * gfr
    del stage

    # ---- the same, one library call per frame (submit_host / select / blobs) ---------------------
    pf_s, pf_frames = None, 0
    if e2e_per_frame:
        pf_frames = F * max(1, e2e_steps // 2)
        LAG = hp.depth - 1
        barrier()
        w0 = time.perf_counter()
        for i in range(pf_frames + LAG):
            if i < pf_frames:
                hp.submit_pinned(pinned[i % n_clip].data_ptr(), row, frame_bytes)
            if i >= LAG:
                hp.select(min(i, pf_frames - 1) - (i - LAG))
                for z in range(S):
                    hp.raw_blobs(z, copy=False)
        torch.cuda.synchronize()
        pf_s = time.perf_counter() - w0

    # ---- the dominant kernel alone (CUDA events around it on the launch stream) ------------------
    hp.set_profiling(True)
    G = hp.group
    det_ms, morph_ms, blob_ms, step_ms = [], [], [], []
    for i in range(max(8, 64 // G)):
        t0 = (i * G) % (n_clip - G + 1)
        hp.enqueue_device_sequence(dev[t0].data_ptr(), G, S * frame_bytes, row, frame_bytes, sh)
        st = hp.stage_ms()
        det_ms.append(st["detect"]); morph_ms.append(st["morph"]); blob_ms.append(st["blobs"]); step_ms.append(st["step"])
    hp.set_profiling(False)
    nbar = hp.mog2_mean_modes() if wl["detector"] == "MOG2" else None
    depth = hp.depth
    hp.close()
    del dev, pinned, host

    if dist is not None:
        tt = torch.tensor([dev_ms, seq_s, pf_s or 0.0, link_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, seq_s, link_ms = float(tt[0]), float(tt[1]), float(tt[3])
        if pf_s:
            pf_s = float(tt[2])
        cnt = torch.tensor([launches, S], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        launches, S_total = int(cnt[0]), int(cnt[1])
    else:
        S_total = S
    return dict(S=S, S_total=S_total, F=F, G=G, depth=depth, dev_ms=dev_ms, launches=launches, info=info,
                seq_s=seq_s, seq_frames=seq_frames, seq_d2h=seq_d2h, e2e_steps=e2e_steps, nblobs=nblobs,
                pf_s=pf_s, pf_frames=pf_frames, det_ms=float(np.mean(det_ms)), n_det=len(det_ms),
                morph_ms=float(np.mean(morph_ms)), blob_ms=float(np.mean(blob_ms)), group_ms=float(np.mean(step_ms)),
                nbar=nbar, frame_bytes=frame_bytes, marks=(mark0, mark1), steps=steps,
                link_ms=link_ms, link_frames=link_frames, link_copy_frames=gfr, frames_per_call=n_clip * per_call)


def roofline_obj(name, wl, m):
    """Algorithmic bytes of ONE launch of the dominant kernel (SURVEY 8d per-pixel figure x the pixels the
    launch processes) over its CUDA-event time, against the measured copy peak."""
    peak, peak_src = peaks()
    H, W, C = wl["H"], wl["W"], wl["C"]
    if wl["detector"] == "MOG2":
        bpp = C + 2.0 * (1 + m["nbar"] * 4 * (2 + C)) + 1
        kname = "mog2_lean_kernel"  # (default MOG2 parameters are "ordinary": launch_mog2 takes the lean kernel)
    elif wl["detector"] == "KNN":
        bpp = C + 21 * (C + 1) + 6 + 1
        kname = "knn_gray_kernel" if C == 1 else "knn_kernel"
    else:
        bpp = 3 * C + 1
        kname = "fd_kernel"
    alg = bpp * H * W * m["S"] * m["G"]  # one launch: G frames of S streams
    achieved = alg / (m["det_ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic_per_launch(kname, name, m["S"], m["G"]),
            "peak_source": peak_src, "kernel_ms": m["det_ms"], "algorithmic_bytes": alg,
            "bytes_per_pixel": bpp, "mean_modes": m["nbar"], "frames_per_launch": m["G"], "streams_per_launch": m["S"],
            "kernel_us_per_frame": m["det_ms"] * 1e3 / m["G"],
            "how": "CUDA events around the kernel on the launch stream, mean of %d launches" % m["n_det"],
            "frac_of_nominal_8TBs": achieved / 8000.0}


def e2e_obj(wl, m):
    mpix = wl["H"] * wl["W"] / 1e6
    clip_fps = m["seq_frames"] * m["S_total"] / m["seq_s"]
    clip_api = ("gofpid_pipeline_process_host_sequence: one call per %d-frame host clip (the %d distinct frames repeated), "
                "pinned host frames in, blob lists with contours of every frame out" % (m["frames_per_call"], wl["clip"]))
    # The end-to-end entry point of a workload is fixed here, not chosen by the result: the clip call, except for
    # the 256-stream workload -- its clip holds 8 time steps only (2.5 GB of frames), less than the frame slots of
    # the pipeline, so a clip call fills and drains the pipeline once per call; a camera feed of that many streams
    # is submitted time step by time step, which keeps the pipeline full across clip repeats.
    per_frame = bool(wl.get("total_streams")) and bool(m["pf_s"])
    fps = m["pf_frames"] * m["S_total"] / m["pf_s"] if per_frame else clip_fps
    out = {"value": fps, "unit": "frames/s",
           "h2d_bytes_per_step": m["frame_bytes"] * m["S_total"] * m["F"],
           "d2h_bytes_per_step": int(m["seq_d2h"] / max(m["e2e_steps"], 1)) * (m["S_total"] // max(m["S"], 1)),
           "frames": (m["pf_frames"] if per_frame else m["seq_frames"]) * m["S_total"],
           "seconds": m["pf_s"] if per_frame else m["seq_s"], "mpix_per_s": fps * mpix,
           "blobs_per_frame": m["nblobs"] / max(m["seq_frames"] * m["S"], 1),
           "api": ("gofpid_pipeline_submit_host + gofpid_pipeline_select + gofpid_pipeline_blobs, one call per time step of all "
                   "streams (%d time steps in flight), pinned host frames in, blob lists with contours out" % m["depth"])
                  if per_frame else clip_api + " (the headline end-to-end entry point)"}
    if per_frame:
        out["clip_call"] = {"value": clip_fps, "unit": "frames/s", "api": clip_api}
    floor = m["link_frames"] * m["S_total"] / (m["link_ms"] * 1e-3)
    out["host_link_floor"] = {
        "value": floor, "unit": "frames/s", "gb_per_s_per_gpu": m["link_frames"] * m["S"] * m["frame_bytes"] / (m["link_ms"] * 1e-3) / 1e9,
        "e2e_over_floor": fps / floor,
        "how": "the same pinned clip copied host->device alone, %d frames per copy, all ranks at once, CUDA events, max over ranks: "
               "the rate the upload of the frames allows at this number of GPUs" % m["link_copy_frames"]}
    if m["pf_s"] and not per_frame:
        out["per_frame_call"] = {
            "value": m["pf_frames"] * m["S_total"] / m["pf_s"], "unit": "frames/s", "frames": m["pf_frames"],
            "api": "gofpid_pipeline_submit_host + gofpid_pipeline_select + gofpid_pipeline_blobs, one call per frame "
                   "(%d frames in flight); reported beside the headline, not instead of it" % m["depth"]}
    return out


def dropin_rates(wl, device):
    """frames/s through the drop-in class (pygofpid_b200.GOFPID, the reference's API): detect() per frame
    (synchronous: H2D, GPU steps 1-4, blob lists D2H, host tracking / post-filter / alarm) and detect_clip()."""
    import cv2 as cv
    from pygofpid_b200 import GOFPID
    se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
    mm = lambda: [dict(fun=cv.erode if op == "erode" else cv.dilate, kernel=se) for op, _ in wl["morph"]]
    clip = make_clip(wl, 48)
    out = {}
    g = GOFPID(frg_detect=wl["detector"], mat_morph=mm(), device=device).initialize()
    for f in clip[:16]:
        g.detect(f)
    t0 = time.perf_counter()
    n = 0
    for _ in range(4):
        for f in clip:
            g.detect(f); n += 1
    out["detect_fps"] = n / (time.perf_counter() - t0)
    out["tracks"] = len(g.tracked_blobs_)
    g = GOFPID(frg_detect=wl["detector"], mat_morph=mm(), device=device).initialize()
    g.detect_clip(clip)
    t0 = time.perf_counter()
    n = 0
    for _ in range(6):
        g.detect_clip(clip); n += len(clip)
    out["detect_clip_fps"] = n / (time.perf_counter() - t0)
    out["api"] = "pygofpid_b200.GOFPID(frg_detect=%r).detect(frame) / .detect_clip(frames): pageable host frames in, " \
                 "alarm decision out, host tracking tail included" % wl["detector"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=1, help="camera streams per GPU (the headline: 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-measurements of the other BASELINE configs")
    ap.add_argument("--no-parity", action="store_true", help="(profiling runs only) skip the self-check against cv2")
    ap.add_argument("--no-dropin", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args, wl, args.workload)
        return

    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pygofpid_b200 import sharding

    # ---- parity of the timed entry points against cv2, before anything is timed -------------------
    parity = None
    if not args.no_parity and rank == 0:
        parity = {args.workload: parity_check(args.workload, wl, local)}

    S = args.streams
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    streams = list(range(rank * S, rank * S + S))
    m = measure(args.workload, wl, S, streams, args.steps, args.warmup, local, dist, world, e2e_per_frame=True, sampler=sampler)
    clocks = sampler.stop(*m["marks"]) if rank == 0 else None
    frames_total = args.steps * m["F"] * m["S_total"]
    fps = frames_total / (m["dev_ms"] * 1e-3)
    mpix = wl["H"] * wl["W"] / 1e6

    line = None
    if rank == 0:
        cfg = config_of(args.workload, wl, S, world)
        line = {
            "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["dev_ms"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+u8", "data": "synthetic",
            "mpix_per_s": fps * mpix, "frames_per_sec_per_gpu": fps / world,
            "config": cfg,
            "details": {"l2": "inputs cycle through a %.0f MB clip (> 126 MB L2); the background model is resident in HBM"
                              % (wl["clip"] * S * m["frame_bytes"] / 1e6),
                        "api": "gofpid_pipeline_process_device_sequence (clip in HBM, %d frames per detector launch, "
                               "%d frame slots) + gofpid_pipeline_join before the closing event" % (m["G"], m["depth"]),
                        "timed_region_s": m["dev_ms"] * 1e-3,
                        "blobs_last_frame": m["info"]["n_kept"], "components_last_frame": m["info"]["n_components"]},
            "e2e": e2e_obj(wl, m),
            "gpu_launches": m["launches"],
            "roofline": roofline_obj(args.workload, wl, m),
            "stage_ms": {"detect_launch": m["det_ms"], "morph_first_frame": m["morph_ms"],
                         "rest_of_group": m["blob_ms"], "group": m["group_ms"], "frames_per_group": m["G"]},
            "clocks": clocks,
            "parity_checked": bool(parity),
        }

    # ---- the other BASELINE configs (N = 1), and the 256-stream config sharded over the ranks -------
    configs = {}
    if not args.no_configs and args.workload == DEFAULT_WORKLOAD and S == 1:
        for name in EXTRA_CONFIGS:
            w2 = WORKLOADS[name]
            total = w2.get("total_streams")
            if world > 1 and not total:
                continue  # single-stream configs are reported at N = 1
            if total:
                lo, hi = sharding.streams_for_rank(total, rank, world)
                strs = list(range(lo, hi))
            else:
                strs = [0]
            if rank == 0 and not args.no_parity:
                parity[name] = parity_check(name, w2, local)
            steps2 = 6 if not total else 4 * world  # (strong scaling: the same time on the clock at every N)
            mm = measure(name, w2, len(strs), strs, steps2, 3, local, dist, world, e2e_per_frame=bool(total))
            # the final host gather of per-stream detections (the only cross-rank step of the path)
            got = sharding.gather_detections({s: mm["info"]["n_kept"] for s in strs[:1]}, dist)
            if rank != 0:
                continue
            f2 = steps2 * mm["F"] * mm["S_total"] / (mm["dev_ms"] * 1e-3)
            entry = {"value": f2, "unit": "frames/s", "mpix_per_s": f2 * w2["H"] * w2["W"] / 1e6,
                     "config": config_of(name, w2, mm["S"], world), "timed_region_s": mm["dev_ms"] * 1e-3,
                     "scaling": "strong" if total else "n/a", "total_streams": mm["S_total"],
                     "e2e": e2e_obj(w2, mm), "roofline": roofline_obj(name, w2, mm), "gpu_launches": mm["launches"],
                     "stage_ms": {"detect_launch": mm["det_ms"], "group": mm["group_ms"], "frames_per_group": mm["G"]},
                     "gathered_ranks": len(got) if got is not None else None}
            if not args.no_cpu_baseline and world == 1:
                r = cpu_reference_run(w2, make_clip(w2, min(w2["clip"], 12)), seconds=3.0)
                entry["cpu_baseline"] = cpu_baseline_obj(r, "one stream of this workload")
            configs[name] = entry

    if rank == 0:
        if configs:
            line["configs"] = configs
        if parity:
            line["parity"] = parity
        if not args.no_dropin and world == 1:
            line["dropin"] = dropin_rates(wl, local)
        if not args.no_cpu_baseline and world == 1:  # (reported at N = 1 only: it takes the host cores ~12 s)
            r = cpu_reference_run(wl, make_clip(wl, 24), seconds=12.0)
            line["cpu_baseline"] = cpu_baseline_obj(r, "the same synthetic workload")
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
