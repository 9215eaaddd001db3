"""Host-side handle of the CUDA hot path: blur -> foreground detection -> morphology -> blobs.

`HotPath` wraps one `gofpid_pipeline` of the C ABI (include/gofpid_b200.h): `n_streams`
independent camera streams of identical shape whose background models stay resident in HBM.
It is what `GOFPID.detect` (pipeline.py of this package) drives per frame, and what `bench.py`
times.  PyTorch is used only to own device buffers (`torch.Tensor.data_ptr()`), never for
arithmetic.

Reference being replaced: pygofpid/pipeline.py:461-479 (the cv2 calls of GOFPID.detect) and
pygofpid/pipeline.py:555-588 / helpers.py:125-168 (blob creation).
"""
import ctypes as ct

import numpy as np

from . import _native as nat


def _as_u8_image(X):
    X = np.asarray(X)
    if X.dtype != np.uint8:
        raise TypeError("frames must be uint8, got %s" % X.dtype)
    if X.ndim == 2:
        return X, 1
    if X.ndim == 3:
        return X, X.shape[2]
    raise ValueError("frame must have shape (H, W) or (H, W, C)")


def gaussian_taps_q8(n, sigma):
    """8.8 fixed-point taps of cv.GaussianBlur's u8 path (host helper of the C ABI)."""
    q = (ct.c_uint16 * max(int(n), 1))()
    nat.check(nat.lib().gofpid_gaussian_taps_q8(int(n), float(sigma), q))
    return np.array(q[:int(n)], dtype=np.uint16)


class PinnedFrames:
    """Page-locked (optionally write-combined) host memory for frames, as a numpy uint8 array
    (gofpid_host_alloc): `PinnedFrames(shape).array[...] = clip`, then hand `.array` to
    process_sequence / submit_pinned.  Freed with the object."""

    def __init__(self, shape, write_combined=False):
        n = int(np.prod(shape))
        self._p = ct.c_void_p()
        nat.check(nat.lib().gofpid_host_alloc(n, 1 if write_combined else 0, ct.byref(self._p)))
        self.array = np.ctypeslib.as_array(ct.cast(self._p, ct.POINTER(ct.c_uint8)), shape=(n,)).reshape(shape)

    def close(self):
        p, self._p = getattr(self, "_p", None), None
        if p:
            self.array = None
            nat.lib().gofpid_host_free(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# numpy view of gofpid_blob (include/gofpid_b200.h)
BLOB_DTYPE = np.dtype([(n, '<i4') for n in ('first', 'x0', 'y0', 'x1', 'y1', 'area_px', 'n_points', 'point_offset',
                                              'nested', 'reserved')] + [(n, '<i8') for n in ('a00', 'a10', 'a01')])


class HotPath:
    """One CUDA pipeline for `n_streams` camera streams of shape (H, W[, C]).

    Parameters
    ----------
    shape : (H, W) or (H, W, C) with C in {1, 3}
    blur : None | ('gaussian', (kw, kh), sigma) | ('box', (kw, kh))
        `sigma` is what the reference passes positionally as `borderType` (pipeline.py:463-467).
    detector : 'MOG2' | 'KNN' | 'FD'
    morph : sequence of ('erode' | 'dilate', kernel ndarray)
    area_min : contours with cv.contourArea < area_min are dropped (pipeline.py:555).
    find_blobs : run labelling / tracing (False: stop after the mask).
    fd_threshold : FrameDifferencing threshold (segmentation.py:36).
    knn_seed : state of the cv::RNG replica KNN draws its update ticks from
        (the reference uses the calling thread's cv.theRNG(); cv.setRNGSeed(s) <-> knn_seed=s).
    """

    def __init__(self, shape, blur=("gaussian", (3, 3), 4.0), detector="MOG2", morph=(),
                 area_min=100.0, n_streams=1, device=0, find_blobs=True, fd_threshold=5,
                 knn_seed=0, max_blobs=1 << 16, max_kept=4096, max_points=0,
                 detect_shadows=True, convert=None, depth=0, mog2_group=0):
        """convert: None | 'bgr2gray' | 'rgb2gray' -- frames are (H, W, 3) and the path works on their
        gray value (cv.cvtColor of pipeline.py:458-459, computed while the frame is loaded).
        max_blobs / max_kept / max_points: per-frame capacities of the pipelined path (components,
        kept blobs, contour points; 0 points: sized from the frame); a frame beyond them is re-run with
        larger buffers when its blobs are fetched, never refused.
        depth: frame slots (frames in flight; 0: default), mog2_group: frames per MOG2 launch in the
        clip entry points (0: default)."""
        L = nat.lib()
        if len(shape) == 2:
            H, W, C = int(shape[0]), int(shape[1]), 1
        elif len(shape) == 3:
            H, W, C = int(shape[0]), int(shape[1]), int(shape[2])
        else:
            raise ValueError("shape must be (H, W) or (H, W, C)")
        cfg = nat.Config()
        L.gofpid_config_default(ct.byref(cfg))
        if blur is None:
            cfg.blur_kind = nat.BLUR_NONE
        else:
            kind = blur[0]
            if kind == "gaussian":
                cfg.blur_kind = nat.BLUR_GAUSSIAN
                cfg.blur_sigma = float(blur[2]) if len(blur) > 2 else 0.0
            elif kind == "box":
                cfg.blur_kind = nat.BLUR_BOX
            else:
                raise ValueError("unknown blur kind %r" % (kind,))
            cfg.blur_kw, cfg.blur_kh = int(blur[1][0]), int(blur[1][1])
        try:
            cfg.detector = {"FD": nat.DETECT_FD, "MOG2": nat.DETECT_MOG2, "KNN": nat.DETECT_KNN}[detector]
        except KeyError:
            raise ValueError("Unknown method for foreground detection")
        cfg.fd_threshold = float(fd_threshold)
        cfg.knn_rng_seed = int(knn_seed)
        cfg.detect_shadows = 1 if detect_shadows else 0
        morph = list(morph or ())
        if len(morph) > nat.MAX_MORPH_OPS:
            raise NotImplementedError("at most %d morphology steps are supported" % nat.MAX_MORPH_OPS)
        self._se = []  # keep the kernel arrays alive while the config is in use
        cfg.n_morph = len(morph)
        for i, (op, kernel) in enumerate(morph):
            k = np.ascontiguousarray(np.asarray(kernel) != 0, dtype=np.uint8)
            if k.ndim != 2:
                raise ValueError("structuring element must be 2-D")
            self._se.append(k)
            cfg.morph[i].op = {"erode": nat.MORPH_ERODE, "dilate": nat.MORPH_DILATE}[op]
            cfg.morph[i].kh, cfg.morph[i].kw = k.shape
            cfg.morph[i].kernel = k.ctypes.data_as(ct.POINTER(ct.c_uint8))
        cfg.find_blobs = 1 if find_blobs else 0
        cfg.blob_area_min = float(area_min)
        cfg.max_blobs, cfg.max_kept, cfg.max_points = int(max_blobs), int(max_kept), int(max_points)
        cfg.convert = {None: 0, "bgr2gray": 1, "rgb2gray": 2}[convert]
        cfg.depth, cfg.mog2_group = int(depth), int(mog2_group)
        # self.C: channels of the frames the caller passes (3 with a gray conversion)
        self.H, self.W, self.C, self.n_streams, self.device = H, W, C, int(n_streams), int(device)
        self.Cm = 1 if convert else C  # channels of the model / state
        self.detector = detector
        self.find_blobs = bool(find_blobs)
        self._h = ct.c_void_p()
        nat.check(L.gofpid_pipeline_create(ct.byref(cfg), H, W, C, self.n_streams, self.device,
                                           ct.byref(self._h)))
        self._L = L

    # ------------------------------------------------------------------ lifetime
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.gofpid_pipeline_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        nat.check(self._L.gofpid_pipeline_reset(self._h))

    # ------------------------------------------------------------------ stepping
    def process(self, frames):
        """One time step from HOST memory.  frames: (H, W[, C]) for one stream or
        (n_streams, H, W[, C]); u8, C-contiguous rows.  Synchronous: copies host->device,
        runs the step, brings the blob lists back."""
        X = np.asarray(frames)
        if X.dtype != np.uint8:
            raise TypeError("frames must be uint8, got %s" % X.dtype)
        want = (self.H, self.W) if self.C == 1 else (self.H, self.W, self.C)
        if X.shape == want or (self.C == 1 and X.shape == (self.H, self.W, 1)):
            if self.n_streams != 1:
                raise ValueError("expected %d frames, got one" % self.n_streams)
            X = X.reshape((1,) + X.shape)
        elif X.shape[1:] == want or (self.C == 1 and X.shape[1:] == (self.H, self.W, 1)):
            if X.shape[0] != self.n_streams:
                raise ValueError("expected %d frames, got %d" % (self.n_streams, X.shape[0]))
        else:
            raise ValueError("Input shape has changed.")
        if not X.flags.c_contiguous:
            X = np.ascontiguousarray(X)
        row = self.W * self.C
        nat.check(self._L.gofpid_pipeline_process_host(self._h, X.ctypes.data, row, row * self.H))

    def process_pinned(self, ptr, pitch, stream_stride):
        """process() for a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        nat.check(self._L.gofpid_pipeline_process_host(self._h, ptr, pitch, stream_stride))

    def submit_pinned(self, ptr, pitch, stream_stride):
        """Asynchronous process_pinned(): enqueue the H2D copy and the step, return at once.
        Up to `depth` frames are in flight; collect results with select(age) + the getters."""
        nat.check(self._L.gofpid_pipeline_submit_host(self._h, ptr, pitch, stream_stride))

    @property
    def depth(self):
        """Frame slots of the handle: frames that can be in flight at once (submit_pinned)."""
        return int(self._L.gofpid_pipeline_depth(self._h))

    def select(self, age=0):
        """Result getters refer to the newest frame (age 0) or the ones before it (age 1, 2)."""
        nat.check(self._L.gofpid_pipeline_select(self._h, age))

    def process_device(self, ptr, pitch, stream_stride, cuda_stream=0):
        """One time step from frames already resident in HBM; asynchronous on `cuda_stream`."""
        nat.check(self._L.gofpid_pipeline_process_device(self._h, ptr, pitch, stream_stride,
                                                         cuda_stream))

    # ------------------------------------------------------------------ results
    def frame_info(self, stream=0):
        fi = nat.FrameInfo()
        nat.check(self._L.gofpid_pipeline_frame_info(self._h, stream, ct.byref(fi)))
        return dict(n_components=fi.n_components, n_candidates=fi.n_candidates, n_kept=fi.n_kept,
                    n_points=fi.n_points, overflow=fi.overflow, frame_index=fi.frame_index)

    def mask(self, stream=0):
        """foreground_mask_ (pipeline.py:476) as a fresh (H, W) uint8 array."""
        out = np.empty((self.H, self.W), np.uint8)
        nat.check(self._L.gofpid_pipeline_mask(self._h, stream, out.ctypes.data, self.W, 0))
        return out

    def labels(self, stream=0):
        """Canonical 8-connected labels: 0 background, else 1 + raster index of first pixel."""
        out = np.empty((self.H, self.W), np.int32)
        nat.check(self._L.gofpid_pipeline_labels(self._h, stream, out.ctypes.data, self.W * 4, 0))
        return out

    def components(self, stream=0):
        """Per-component stats keyed by canonical label (first pixel index + 1)."""
        n = ct.c_int(0)
        nat.check(self._L.gofpid_pipeline_components(self._h, stream, None, 0, ct.byref(n)))
        buf = np.zeros((max(n.value, 1), 6), np.int32)
        nat.check(self._L.gofpid_pipeline_components(
            self._h, stream, buf.ctypes.data_as(ct.POINTER(ct.c_int32)), buf.shape[0], ct.byref(n)))
        return {int(r[0]) + 1: dict(area=int(r[5]), x0=int(r[1]), y0=int(r[2]), x1=int(r[3]), y1=int(r[4]))
                for r in buf[:n.value]}

    def raw_blobs(self, stream=0, copy=True):
        """(records, points): structured blob records in the reference's order and one
        (n_points, 2) int32 array of contour points.

        One library call into buffers the handle keeps (sized for a normal frame; a busier frame
        -- the call reports the sizes it needs -- grows them).  copy=False returns views of those
        buffers, valid until the next call."""
        nb, npt = ct.c_int(0), ct.c_int(0)
        while True:
            buf = getattr(self, "_blob_buf", None)
            if buf is None:
                buf = self._blob_buf = ((nat.Blob * 64)(), np.empty((8192, 2), np.int32))
            recs, pts = buf
            rc = self._L.gofpid_pipeline_blobs(self._h, stream, recs, len(recs),
                                               pts.ctypes.data_as(ct.POINTER(ct.c_int32)), pts.shape[0],
                                               ct.byref(nb), ct.byref(npt))
            if rc == nat.ERR_INVALID and (nb.value > len(recs) or npt.value > pts.shape[0]):
                self._blob_buf = ((nat.Blob * max(len(recs), 2 * nb.value))(),
                                  np.empty((max(pts.shape[0], 2 * npt.value), 2), np.int32))
                continue
            nat.check(rc)
            out = pts[:npt.value]
            if not copy:
                return recs[:nb.value], out
            # an own copy of the records too (a ctypes array of gofpid_blob) and of the points
            own = (nat.Blob * max(nb.value, 1))()
            ct.memmove(own, recs, nb.value * ct.sizeof(nat.Blob))
            own_t = (nat.Blob * nb.value).from_buffer(own) if nb.value else (nat.Blob * 0)()
            return own_t, out.copy()

    def process_sequence(self, frames, blob_capacity=None, point_capacity=None):
        """A whole clip from HOST memory -- or a CUDA uint8 tensor -- in one library call
        (gofpid_pipeline_process_host_sequence / gofpid_pipeline_process_device_sequence):
        frames (T, H, W[, C]) for one stream or (T, n_streams, H, W[, C]).  Returns (counts, records,
        points): counts[t, z] = (blobs, points) of frame t / stream z, records a numpy structured
        array (fields of gofpid_blob) and points (n, 2) int32, both in frame order."""
        per = (self.H, self.W) if self.C == 1 else (self.H, self.W, self.C)
        if hasattr(frames, "data_ptr"):  # a (pinned) torch uint8 tensor: used in place
            if not frames.is_contiguous() or frames.element_size() != 1:
                raise TypeError("frames must be a contiguous uint8 tensor")
            shape, ptr, keep = tuple(frames.shape), frames.data_ptr(), frames
        else:
            keep = np.ascontiguousarray(frames, dtype=np.uint8)
            shape, ptr = keep.shape, keep.ctypes.data
        if shape[1:] == per and self.n_streams == 1:
            shape = (shape[0], 1) + per
        if shape[1:] != (self.n_streams,) + per:
            raise ValueError("Input shape has changed.")
        T = shape[0]
        row = self.W * self.C
        # default capacities: 256 blobs and 8 (H + W) contour points per frame and stream; a clip that
        # needs more still runs to its end inside the library, which keeps the lists and reports the
        # sizes -- they are fetched below with gofpid_pipeline_sequence_results
        bc = int(blob_capacity or 256 * T * self.n_streams)
        pc = int(point_capacity or max(4096, 8 * (self.H + self.W)) * T * self.n_streams)
        counts = np.zeros((T, self.n_streams, 2), np.int32)
        recs = np.empty(bc, BLOB_DTYPE)
        pts = np.empty((pc, 2), np.int32)
        if getattr(frames, "is_cuda", False):  # a clip that already is in device memory
            import torch
            rc = self._L.gofpid_pipeline_process_device_sequence(
                self._h, ptr, T, row * self.H * self.n_streams, row, row * self.H,
                torch.cuda.current_stream(frames.device).cuda_stream,
                recs.ctypes.data_as(ct.POINTER(nat.Blob)), bc, pts.ctypes.data_as(ct.POINTER(ct.c_int32)), pc,
                counts.ctypes.data_as(ct.POINTER(ct.c_int32)))
        else:
            rc = self._L.gofpid_pipeline_process_host_sequence(
                self._h, ptr, T, row * self.H * self.n_streams, row, row * self.H,
                recs.ctypes.data_as(ct.POINTER(nat.Blob)), bc, pts.ctypes.data_as(ct.POINTER(ct.c_int32)), pc,
                counts.ctypes.data_as(ct.POINTER(ct.c_int32)))
        nb, npt = int(counts[..., 0].sum()), int(counts[..., 1].sum())
        if rc == nat.ERR_INVALID and "gofpid_pipeline_sequence_results" in nat.last_error():
            recs = np.empty(max(nb, 1), BLOB_DTYPE)
            pts = np.empty((max(npt, 1), 2), np.int32)
            rc = self._L.gofpid_pipeline_sequence_results(
                self._h, recs.ctypes.data_as(ct.POINTER(nat.Blob)), len(recs),
                pts.ctypes.data_as(ct.POINTER(ct.c_int32)), len(pts))
        nat.check(rc)
        return counts, recs[:nb], pts[:npt]

    def enqueue_device_sequence(self, ptr, n_frames, frame_stride, pitch, stream_stride, cuda_stream=0):
        """A clip in DEVICE memory, enqueued only (gofpid_pipeline_process_device_sequence without
        collection): the newest `depth` frames stay selectable afterwards."""
        nat.check(self._L.gofpid_pipeline_process_device_sequence(
            self._h, ptr, int(n_frames), int(frame_stride), int(pitch), int(stream_stride), cuda_stream,
            None, 0, None, 0, None))

    def join(self, cuda_stream=0):
        """Orders `cuda_stream` after everything enqueued so far, stage B on the internal streams
        included (gofpid_pipeline_join); does not block the host."""
        nat.check(self._L.gofpid_pipeline_join(self._h, cuda_stream))

    @property
    def group(self):
        """Frames per detector launch in the clip entry points (temporal blocking of MOG2)."""
        return int(self._L.gofpid_pipeline_group(self._h))

    def blobs(self, stream=0, anchor="bottom"):
        """self.blobs_ of the reference (pipeline.py:555-588): list of dicts with keys
        contour (int32 (N,1,2)), anchor / bottom / center (int16 (2,))."""
        recs, pts = self.raw_blobs(stream)
        return [_blob_dict(pts, b.point_offset, b.n_points, b.y1, b.a00, b.a10, b.a01, anchor) for b in recs]

    # ------------------------------------------------------------------ model state
    def mog2_model(self, stream=0):
        n = self.H * self.W
        w = np.empty((self.H, self.W, 5), np.float32)
        v = np.empty((self.H, self.W, 5), np.float32)
        m = np.empty((self.H, self.W, 5, self.Cm), np.float32)
        k = np.empty((self.H, self.W), np.uint8)
        f32p, u8p = ct.POINTER(ct.c_float), ct.POINTER(ct.c_uint8)
        nat.check(self._L.gofpid_pipeline_mog2_model(self._h, stream, w.ctypes.data_as(f32p),
                                                     v.ctypes.data_as(f32p), m.ctypes.data_as(f32p),
                                                     k.ctypes.data_as(u8p)))
        del n
        return w, v, m, k

    def knn_model(self, stream=0):
        s = np.empty((self.H, self.W, 21, self.Cm + 1), np.uint8)
        idx = np.empty((3, self.H, self.W), np.uint8)
        nxt = np.empty((3, self.H, self.W), np.uint8)
        cnt = np.zeros(3, np.int32)
        rng = ct.c_uint64(0)
        u8p = ct.POINTER(ct.c_uint8)
        nat.check(self._L.gofpid_pipeline_knn_model(
            self._h, stream, s.ctypes.data_as(u8p), idx.ctypes.data_as(u8p), nxt.ctypes.data_as(u8p),
            cnt.ctypes.data_as(ct.POINTER(ct.c_int32)), ct.byref(rng)))
        return s, idx, nxt, cnt, rng.value

    def fd_model(self, stream=0):
        prev = np.empty((self.H, self.W) if self.Cm == 1 else (self.H, self.W, self.Cm), np.uint8)
        valid = ct.c_int(0)
        nat.check(self._L.gofpid_pipeline_fd_model(self._h, stream,
                                                   prev.ctypes.data_as(ct.POINTER(ct.c_uint8)),
                                                   ct.byref(valid)))
        return prev if valid.value else None

    def mog2_mean_modes(self):
        v = ct.c_double(0)
        nat.check(self._L.gofpid_pipeline_mog2_mean_modes(self._h, ct.byref(v)))
        return v.value

    # ------------------------------------------------------------------ measurement aids
    def set_profiling(self, on=True):
        nat.check(self._L.gofpid_pipeline_set_profiling(self._h, 1 if on else 0))

    def stage_ms(self):
        ms = (ct.c_float * 5)()
        nat.check(self._L.gofpid_pipeline_stage_ms(self._h, ms))
        return dict(detect=ms[0], morph=ms[1], blobs=ms[2], tick=ms[3], step=ms[4])

    def launch_count(self):
        return int(self._L.gofpid_pipeline_launch_count(self._h))

    def launch_total(self):
        """Kernels launched since creation / the last reset."""
        n = ct.c_longlong(0)
        nat.check(self._L.gofpid_pipeline_launch_total(self._h, ct.byref(n)))
        return int(n.value)


def _blob_dict(pts, point_offset, n_points, y1, a00, a10, a01, anchor):
    """One entry of the reference's blobs_ (pipeline.py:572-588) from a gofpid_blob record."""
    contour = pts[point_offset:point_offset + n_points].reshape(-1, 1, 2).copy()
    # cv.moments of a contour (helpers.py:156,165): m00 = a00/2, m10 = a10/6, m01 = a01/6,
    # with the sign of a00 folded into the factors
    a00, a10, a01 = float(a00), float(a10), float(a01)
    if a00 > 0:
        h, s = 0.5, 0.16666666666666666666666666666667
    else:
        h, s = -0.5, -0.16666666666666666666666666666667
    m00, m10, m01 = a00 * h, a10 * s, a01 * s
    cx = int(m10 / m00)
    bottom = np.array([cx, y1], dtype=np.int16)
    center = np.array([cx, int(m01 / m00)], dtype=np.int16)
    return dict(contour=contour, anchor=bottom if anchor == "bottom" else center, bottom=bottom, center=center)


def blobs_from_records(recs, pts, anchor="bottom"):
    """blobs_ lists from one frame's slice of process_sequence's output (records: BLOB_DTYPE rows,
    pts: that frame's contour points; point_offset restarts at 0 for every frame)."""
    if not isinstance(recs, np.ndarray):  # a ctypes array of gofpid_blob (HotPath.raw_blobs)
        return [_blob_dict(pts, b.point_offset, b.n_points, b.y1, b.a00, b.a10, b.a01, anchor) for b in recs]
    return [_blob_dict(pts, int(b['point_offset']), int(b['n_points']), int(b['y1']), int(b['a00']), int(b['a10']),
                       int(b['a01']), anchor) for b in recs]


# ---------------------------------------------------------------------- stage-level wrappers
def _dev_u8(t):
    import torch
    assert t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()
    return t.data_ptr()


def blur_u8(src, kind, ksize, sigma=0.0):
    """cv.GaussianBlur / cv.blur on a CUDA uint8 tensor (H, W[, C]); returns a new tensor."""
    import torch
    H, W = src.shape[:2]
    C = 1 if src.dim() == 2 else src.shape[2]
    dst = torch.empty_like(src)
    k = {"none": nat.BLUR_NONE, "gaussian": nat.BLUR_GAUSSIAN, "box": nat.BLUR_BOX}[kind]
    nat.check(nat.lib().gofpid_blur_u8(_dev_u8(src), W * C, _dev_u8(dst), W * C, H, W, C, k,
                                       int(ksize[0]), int(ksize[1]), float(sigma),
                                       torch.cuda.current_stream().cuda_stream))
    return dst


def morph_u8(src, ops, levels=None):
    """Fused chain of ('erode' | 'dilate', kernel) steps on a CUDA uint8 (H, W) tensor.

    `levels` (1..3 ascending values): the mask only holds 0 and these values; runs the bit-plane
    kernel the pipeline uses instead of the byte kernel."""
    import torch
    H, W = src.shape
    dst = torch.empty_like(src)
    arr = (nat.MorphOp * max(len(ops), 1))()
    keep = []
    for i, (op, kernel) in enumerate(ops):
        k = np.ascontiguousarray(np.asarray(kernel) != 0, dtype=np.uint8)
        keep.append(k)
        arr[i].op = {"erode": nat.MORPH_ERODE, "dilate": nat.MORPH_DILATE}[op]
        arr[i].kh, arr[i].kw = k.shape
        arr[i].kernel = k.ctypes.data_as(ct.POINTER(ct.c_uint8))
    if levels is not None:
        lv = (ct.c_int * len(levels))(*[int(v) for v in levels])
        nat.check(nat.lib().gofpid_morph_levels_u8(_dev_u8(src), W, _dev_u8(dst), W, H, W, arr, len(ops),
                                                   lv, len(levels), torch.cuda.current_stream().cuda_stream))
        return dst
    nat.check(nat.lib().gofpid_morph_u8(_dev_u8(src), W, _dev_u8(dst), W, H, W, arr, len(ops),
                                        torch.cuda.current_stream().cuda_stream))
    return dst


def ccl_u8(mask):
    """Canonical 8-connected labels (int32) of a CUDA uint8 (H, W) mask."""
    import torch
    H, W = mask.shape
    out = torch.empty((H, W), dtype=torch.int32, device=mask.device)
    nat.check(nat.lib().gofpid_ccl_u8(_dev_u8(mask), W, out.data_ptr(), W * 4, H, W,
                                      torch.cuda.current_stream().cuda_stream))
    return out
