#!/usr/bin/env python
"""bench.py -- throughput of the PyGOFPID foreground-extraction hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass over a 72-frame clip of one camera stream (per GPU) through the whole hot path: Gaussian blur -> MOG2 -> open/close -> connected-component blobs
(BASELINE.json configs[1], the configuration the metric is quoted on).  N > 1 is launched by
torchrun, one rank per GPU, each rank owning its own independent stream(s) (weak scaling, no
data-path collective; torch.distributed is used for the barrier and the max-over-ranks only).

Output: ONE JSON line (see the task contract): frames/s with inputs resident in HBM (`value`),
the same through the C-ABI host entry point with pinned host frames (`e2e`), the roofline of the
dominant kernel (fused blur+MOG2) and the CPU baseline (the reference's own OpenCV calls,
oracle/cv_path.py) timed on this box's host cores.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLIP_FRAMES = 72  # distinct frames cycled per stream: 72 x 2.07 MB = 149 MB > 126 MB of L2
# a step = one pass over the device-resident clip (72 time steps of every stream at 1 stream / GPU)
# through the clip entry point gofpid_pipeline_process_device_sequence


def rect(n):
    return np.ones((n, n), np.uint8)


def ellipse15():
    # cv.getStructuringElement(MORPH_ELLIPSE, (15, 15)) of OpenCV 4.13 (row widths, centred)
    widths = [1, 9, 11, 13, 13, 15, 15, 15, 15, 15, 13, 13, 11, 9, 1]
    k = np.zeros((15, 15), np.uint8)
    for r, w in enumerate(widths):
        k[r, (15 - w) // 2:(15 - w) // 2 + w] = 1
    return k


WORKLOADS = {
    # BASELINE.json configs[1] -- the headline
    "1080p_mog2_open_close_blobs": dict(H=1080, W=1920, C=1, detector="MOG2", n_obj=12, seed=2,
                                        morph=[("erode", 5), ("dilate", 5), ("dilate", 5), ("erode", 5)]),
    "1080p_bgr_mog2_open_close_blobs": dict(H=1080, W=1920, C=3, detector="MOG2", n_obj=12, seed=2,
                                            morph=[("erode", 5), ("dilate", 5), ("dilate", 5), ("erode", 5)]),
    # configs[0]
    "320x240_fd_open_blobs": dict(H=240, W=320, C=1, detector="FD", n_obj=6, seed=1,
                                  morph=[("erode", 3), ("dilate", 3)]),
    # configs[2]
    "4k_mog2_ellipse15_open_blobs": dict(H=2160, W=3840, C=1, detector="MOG2", n_obj=16, seed=3,
                                         morph=[("erode", "e15"), ("dilate", "e15")]),
    # configs[3] (per-GPU share of the 256 streams is set with --streams)
    "ilids_720x576_bgr_defaults": dict(H=576, W=720, C=3, detector="MOG2", n_obj=6, seed=1000,
                                       morph=[("erode", 5), ("dilate", 5)]),
    # configs[4]
    "1080p_knn_many_blobs": dict(H=1080, W=1920, C=1, detector="KNN", n_obj=400, seed=5,
                                 obj_size=(12, 90, 8, 80), morph=[("erode", 5), ("dilate", 5)]),
}
DEFAULT_WORKLOAD = "1080p_mog2_open_close_blobs"


def morph_list(spec):
    return [(op, ellipse15() if k == "e15" else rect(k)) for op, k in spec]


def make_clip(wl, n_frames, stream=0):
    from pygofpid_b200.synth import SyntheticClip
    clip = SyntheticClip(wl["H"], wl["W"], wl["C"], seed=wl["seed"] + stream, n_obj=wl["n_obj"],
                         noise_sigma=4.0, obj_size=wl.get("obj_size"))
    return np.stack(clip.frames(n_frames))


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

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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
def cpu_reference_run(wl, steps, warmup, frames_per_step):
    """The reference's own OpenCV CPU path (oracle/cv_path.py issues the cv2 calls of
    GOFPID.detect pipeline.py:461-479 + _find_blob :555-588) on all host cores."""
    import cv2 as cv
    from oracle.cv_path import CvHotPath
    cv.setNumThreads(os.cpu_count() or 1)
    cv.setRNGSeed(12345)
    se = {5: cv.getStructuringElement(cv.MORPH_RECT, (5, 5)), 3: cv.getStructuringElement(cv.MORPH_RECT, (3, 3)),
          "e15": cv.getStructuringElement(cv.MORPH_ELLIPSE, (15, 15))}
    mm = [dict(fun=cv.erode if op == "erode" else cv.dilate, kernel=se[k]) for op, k in wl["morph"]]
    path = CvHotPath(blur="default", frg_detect=wl["detector"], mat_morph=mm)
    n_clip = min(CLIP_FRAMES, 24)
    frames = make_clip(wl, n_clip)
    i = 0
    for _ in range(warmup * frames_per_step):
        path.step(frames[i % n_clip]); i += 1
    t0 = time.perf_counter()
    nb = 0
    for _ in range(steps * frames_per_step):
        _, blobs = path.step(frames[i % n_clip]); i += 1
        nb += len(blobs)
    dt = time.perf_counter() - t0
    n = steps * frames_per_step
    return dict(fps=n / dt, seconds=dt, frames=n, threads=cv.getNumThreads(), cores=os.cpu_count(),
                cv=cv.__version__, blobs_per_frame=nb / max(n, 1))


def run_reference_arm(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps_frames = 8  # bounded sample per step
    r = cpu_reference_run(wl, args.steps, min(args.warmup, 3), fps_frames)
    mpix = wl["H"] * wl["W"] / 1e6
    line = {
        "impl": "reference", "metric": "frames_per_sec", "value": r["fps"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
        "mpix_per_s": r["fps"] * mpix,
        "config": {"workload": name, "frames_per_step": fps_frames, "cv2": r["cv"],
                   "note": "reference's OpenCV CPU hot path (blur, apply, erode/dilate, findContours+area+moments)"},
        "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["threads"], "kind": "port",
                         "sample": "%d frames of the same synthetic clip, cv2 %s with %d threads on %d host cores"
                                   % (r["frames"], r["cv"], r["threads"], r["cores"])},
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


def traffic_per_launch(name, workload, streams, frames_per_launch):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (taken on the
    headline workload, one stream, profiles/traffic.json's frames per launch); None otherwise."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p) and workload == DEFAULT_WORKLOAD and streams == 1:
        try:
            d = json.load(open(p))
            return d.get(name) if d.get("_frames_per_launch", 1) == frames_per_launch else None
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=1, help="camera streams per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args, wl, args.workload)
        return

    import torch
    from pygofpid_b200 import HotPath

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

    H, W, C, S = wl["H"], wl["W"], wl["C"], args.streams
    hp = HotPath((H, W) if C == 1 else (H, W, C), blur=("gaussian", (3, 3), 4.0), detector=wl["detector"],
                 morph=morph_list(wl["morph"]), area_min=100.0, n_streams=S, device=local, knn_seed=12345)
    # synthetic clip: every rank / stream gets its own seed (independent cameras)
    n_clip = CLIP_FRAMES if S == 1 else max(8, CLIP_FRAMES // S)
    F = n_clip  # frames (time steps) per step
    host = np.stack([make_clip(wl, n_clip, stream=rank * S + s) for s in range(S)], axis=1)  # (T, S, H, W[,C])
    pinned = torch.from_numpy(host).pin_memory()
    dev = pinned.cuda(non_blocking=False)
    row = W * C
    frame_bytes = row * H
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def device_step(t0):
        # the whole clip, already in HBM, through the clip entry point (MOG2: hp.group frames per
        # detector launch); per-frame results land in pinned host memory as part of every step
        hp.enqueue_device_sequence(dev.data_ptr(), n_clip, S * frame_bytes, row, frame_bytes, sh)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also burns in the background model: >= 20 frames) ---------------------
    t_idx = 0
    for _ in range(args.warmup):
        device_step(t_idx); t_idx += F
    hp.frame_info()  # first result fetch outside the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- timed: device-resident frames ----------------------------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = hp.launch_total()
    e0.record(stream)
    for _ in range(args.steps):
        device_step(t_idx); t_idx += F
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = (hp.launch_total() - l0) * world  # every rank launches the same kernels on its own streams
    info = hp.frame_info()

    # ---- timed: end to end through the host entry points (pinned frames in, blob lists out) ----
    # Every frame is copied host->device inside the timed region and its blob list is read back
    # to the host; frames are submitted ahead of the one whose results are collected (hp.depth frames in
    # flight, the library's pipelined host API: submit_host / select / blobs).
    e2e_frames = F * max(1, args.steps // 4)
    for i in range(4):  # untimed: first use of the host path allocates its staging buffers
        hp.process_pinned(pinned[(t_idx + i) % n_clip].data_ptr(), row, frame_bytes)
    t_idx += 4
    barrier()
    w0 = time.perf_counter()
    d2h = 0
    nblobs = 0
    LAG = hp.depth - 1  # frames submitted ahead of the one whose results are collected
    for i in range(e2e_frames + LAG):
        if i < e2e_frames:
            t = (t_idx + i) % n_clip
            hp.submit_pinned(pinned[t].data_ptr(), row, frame_bytes)
        if i >= LAG:
            hp.select(min(i, e2e_frames - 1) - (i - LAG))
            for z in range(S):
                recs, pts = hp.raw_blobs(z, copy=False)
                nblobs += len(recs)
                d2h += 32 + len(recs) * 56 + pts.shape[0] * 8
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    t_idx += e2e_frames

    # ---- the same end-to-end work through the clip entry point (one library call for a whole
    # clip in pinned host memory: gofpid_pipeline_process_host_sequence) -------------------------
    seq_calls = max(1, e2e_frames // n_clip)
    seq_frames = seq_calls * n_clip
    seq_fps, seq_s, seq_d2h = None, None, 0
    if n_clip >= 8:
        hp.process_sequence(pinned[:8])  # untimed: sizes the output buffers
        barrier()
        w0 = time.perf_counter()
        for _ in range(seq_calls):
            counts, recs, pts = hp.process_sequence(pinned)
            seq_d2h += 32 * n_clip * S + len(recs) * 56 + pts.shape[0] * 8
        torch.cuda.synchronize()
        seq_s = time.perf_counter() - w0
        seq_fps = seq_frames * S / seq_s
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel time of the dominant kernel (CUDA events around it, same steps) --------
    # one detector launch = G consecutive frames of every stream (G = hp.group; 1 unless MOG2); the
    # stages run on one stream here, so the events bracket the launch alone
    hp.set_profiling(True)
    G = hp.group
    det_ms, morph_ms, blob_ms, step_ms = [], [], [], []
    for i in range(max(8, 64 // G)):
        t0 = (i * G) % (n_clip - G + 1)
        hp.enqueue_device_sequence(dev[t0].data_ptr(), G, S * frame_bytes, row, frame_bytes, sh)
        st = hp.stage_ms()
        det_ms.append(st["detect"]); morph_ms.append(st["morph"]); blob_ms.append(st["blobs"])
        step_ms.append(st["step"])
    hp.set_profiling(False)
    det = float(np.mean(det_ms))

    # ---- aggregate over ranks ---------------------------------------------------------------
    if dist is not None:
        tt = torch.tensor([dev_ms, e2e_s, seq_s or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(tt[0]), float(tt[1])
        if seq_s:
            seq_fps = seq_frames * S / float(tt[2])
    frames_total = args.steps * F * S * world
    fps = frames_total / (dev_ms * 1e-3)
    e2e_fps = e2e_frames * S * world / e2e_s
    mpix = H * W / 1e6

    if rank == 0:
        peak, peak_src = peaks()
        if wl["detector"] == "MOG2":
            nbar = hp.mog2_mean_modes()
            bpp = C + 2.0 * (1 + nbar * 4 * (2 + C)) + 1
            kname = "mog2_kernel"
        elif wl["detector"] == "KNN":
            nbar = None
            bpp = C + 21 * (C + 1) + 6 + 1
            kname = "knn_kernel"
        else:
            nbar = None
            bpp = 3 * C + 1
            kname = "fd_kernel"
        per_frame = {"value": e2e_fps, "unit": "frames/s", "frames": e2e_frames,
                     "api": "gofpid_pipeline_submit_host + gofpid_pipeline_select + gofpid_pipeline_blobs, one call "
                            "per frame (pinned host frames in, blob lists with contours out, %d frames in flight)" % hp.depth}
        clip = None if seq_fps is None else {
            "value": seq_fps * world, "unit": "frames/s", "frames": seq_frames,
            "api": "gofpid_pipeline_process_host_sequence, one call per 72-frame clip (pinned host frames in, "
                   "blob lists with contours of every frame out)"}
        # headline: the faster of the two public host entry points (both copy every frame host->device
        # and every blob list device->host inside the timed region); the other one is reported beside it
        use_clip = clip is not None and clip["value"] >= e2e_fps
        e2e_obj = {"value": clip["value"] if use_clip else e2e_fps, "unit": "frames/s",
                   "h2d_bytes_per_step": frame_bytes * S * F,
                   "d2h_bytes_per_step": int((seq_d2h / max(seq_frames, 1) if use_clip else d2h / max(e2e_frames, 1)) * F),
                   "frames": seq_frames if use_clip else e2e_frames,
                   "mpix_per_s": (clip["value"] if use_clip else e2e_fps) * mpix,
                   "blobs_per_frame": nblobs / max(e2e_frames * S, 1),
                   "api": (clip if use_clip else per_frame)["api"],
                   "clip_call": clip, "per_frame_call": per_frame}
        alg_bytes = bpp * H * W * S * G  # one launch: G frames of S streams
        achieved = alg_bytes / (det * 1e-3) / 1e9
        line = {
            "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+u8", "data": "synthetic",
            "mpix_per_s": fps * mpix, "frames_per_sec_per_gpu": fps / world,
            "config": {"workload": args.workload, "frames_per_step": F, "streams_per_gpu": S,
                       "shape": [H, W, C], "clip_frames": n_clip,
                       "l2": "inputs cycle through a %.0f MB clip (> 126 MB L2); the background model is resident in HBM"
                             % (n_clip * S * frame_bytes / 1e6),
                       "api": "gofpid_pipeline_process_device_sequence (clip in HBM, %d frames per detector launch, "
                              "%d frame slots)" % (hp.group, hp.depth),
                       "blobs_last_frame": info["n_kept"], "components_last_frame": info["n_components"]},
            "e2e": e2e_obj,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic_per_launch(kname, args.workload, S, G),
                         "peak_source": peak_src, "kernel_ms": det, "algorithmic_bytes": alg_bytes,
                         "bytes_per_pixel": bpp, "mean_modes": nbar, "frames_per_launch": G,
                         "kernel_us_per_frame": det * 1e3 / G,
                         "how": "CUDA events around the kernel on the launch stream, mean of %d launches" % len(det_ms),
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "stage_ms": {"detect_launch": det, "morph_first_frame": float(np.mean(morph_ms)),
                         "rest_of_group": float(np.mean(blob_ms)), "group": float(np.mean(step_ms)),
                         "frames_per_group": G},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:  # (reported at N = 1 only: it takes the host cores ~20 s)
            r = cpu_reference_run(wl, 10, 2, 8)
            line["cpu_baseline"] = {
                "value": r["fps"], "unit": "frames/s", "cores": r["threads"], "kind": "port",
                "sample": "%d frames of the same synthetic workload through the reference's cv2 %s calls "
                          "(oracle/cv_path.py), %d cv2 threads on %d host cores"
                          % (r["frames"], r["cv"], r["threads"], r["cores"])}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
