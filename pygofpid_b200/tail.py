"""Native host tail of `GOFPID.detect`: tracking, post-filtering and alarm (steps 5-7).

Thin ctypes handle of `gofpid_tracker` (include/gofpid_b200.h; pygofpid_b200/csrc/tracker.cu), which
replaces the reference's per-blob / per-track Python of pygofpid/pipeline.py:590-799.  It keeps one
camera stream's tracked blobs and alarm state on the host -- pure host code, no CUDA call -- and is fed
either the blob records a `HotPath` returns (no Python object per blob) or the reference's blob
dictionaries.  `tracks()` materialises the reference's `tracked_blobs_` list on demand.
"""
import ctypes as ct

import numpy as np

from . import _native as nat

_FILTERS = ((1, 'presence'), (2, 'distance'), (4, 'perimeter'), (8, 'perspective'))


class NativeTail:
    def __init__(self, H, persp_height, persp_area, perimeter, tracking, post_filter, alarm_def):
        self._L = nat.lib()
        ph = np.ascontiguousarray(persp_height, dtype=np.float64)
        pa = np.ascontiguousarray(persp_area, dtype=np.float64)
        per = np.ascontiguousarray(np.asarray(perimeter).reshape(-1, 2), dtype=np.int32)
        if len(ph) != H or len(pa) != H:
            raise ValueError("perspective tables must have one entry per frame row")
        cfg = nat.TrackerConfig()
        cfg.H = int(H)
        cfg.persp_height = ph.ctypes.data_as(ct.POINTER(ct.c_double))
        cfg.persp_area = pa.ctypes.data_as(ct.POINTER(ct.c_double))
        cfg.perimeter = per.ctypes.data_as(ct.POINTER(ct.c_int32))
        cfg.n_perimeter = len(per)
        cfg.tracking_factor = float(tracking['factor'])
        cfg.absence_max = int(tracking['absence_max'])
        cfg.anchor_center = 1 if post_filter['anchor'] == 'center' else 0
        cfg.perspective_coeff = float(post_filter['perspective_coeff'])
        cfg.presence_min = float(post_filter['presence_min'])
        cfg.distance_min = float(post_filter['distance_min'])
        cfg.keep_alarm = 1 if alarm_def['keep_alarm'] else 0
        cfg.duration_min = float(alarm_def['duration_min'])
        self._h = ct.c_void_p()
        nat.check(self._L.gofpid_tracker_create(ct.byref(cfg), ct.byref(self._h)))  # (the tables are copied)
        self._alarm = ct.c_int(0)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.gofpid_tracker_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        nat.check(self._L.gofpid_tracker_reset(self._h))

    # ------------------------------------------------------------------ stepping
    def step_records(self, recs, n, pts):
        """One frame from `n` gofpid_blob records (a ctypes array or a BLOB_DTYPE numpy array) whose
        point_offset index into `pts` ((N, 2) int32)."""
        rp = recs.ctypes.data if isinstance(recs, np.ndarray) else ct.addressof(recs)
        nat.check(self._L.gofpid_tracker_step(self._h, rp, int(n), pts.ctypes.data, ct.byref(self._alarm)))
        return bool(self._alarm.value)

    def step_dicts(self, blobs):
        """One frame from the reference's blob dictionaries (contour, anchor, bottom, center)."""
        n = len(blobs)
        anchor = np.zeros((max(n, 1), 2), np.int16)
        bottom = np.zeros((max(n, 1), 2), np.int16)
        center = np.zeros((max(n, 1), 2), np.int16)
        boxes = np.zeros((max(n, 1), 4), np.int32)
        offs = np.zeros(n + 1, np.int32)
        cs = []
        for i, b in enumerate(blobs):
            c = np.ascontiguousarray(np.asarray(b['contour']).reshape(-1, 2), dtype=np.int32)
            cs.append(c)
            anchor[i], bottom[i], center[i] = b['anchor'], b['bottom'], b['center']
            lo, hi = c.min(axis=0), c.max(axis=0)  # cv.boundingRect of the contour
            boxes[i] = (lo[0], lo[1], hi[0], hi[1])
            offs[i + 1] = offs[i] + len(c)
        pts = np.concatenate(cs) if cs else np.zeros((1, 2), np.int32)
        nat.check(self._L.gofpid_tracker_step_prepared(
            self._h, n, anchor.ctypes.data, bottom.ctypes.data, center.ctypes.data, boxes.ctypes.data,
            pts.ctypes.data, offs.ctypes.data, ct.byref(self._alarm)))
        return bool(self._alarm.value)

    def run(self, counts, recs, pts, stream=0):
        """A clip: the (counts, records, points) of HotPath.process_sequence -> list of bool, one per frame."""
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        T, S = counts.shape[0], counts.shape[1]
        alarms = np.zeros(max(T, 1), np.uint8)
        nat.check(self._L.gofpid_tracker_run(self._h, recs.ctypes.data if len(recs) else None,
                                             pts.ctypes.data if len(pts) else None, counts.ctypes.data, T, S,
                                             int(stream), alarms.ctypes.data))
        return [bool(a) for a in alarms[:T]]

    # ------------------------------------------------------------------ state
    def __len__(self):
        return int(self._L.gofpid_tracker_count(self._h))

    def tracks(self):
        """The reference's tracked_blobs_ (pipeline.py:630-645): a list of fresh dictionaries."""
        out = []
        tr = nat.Track()
        for i in range(len(self)):
            nat.check(self._L.gofpid_tracker_get(self._h, i, ct.byref(tr)))
            contour = np.empty((max(tr.n_points, 1), 2), np.int32)
            nat.check(self._L.gofpid_tracker_contour(self._h, i, contour.ctypes.data, len(contour)))
            out.append({'contour': contour[:tr.n_points].reshape(-1, 1, 2),
                        'anchor': np.array(tr.anchor[:], np.int16), 'bottom': np.array(tr.bottom[:], np.int16),
                        'center': np.array(tr.center[:], np.int16),
                        'center_first': np.array(tr.center_first[:], np.int16),
                        'presence': int(tr.presence), 'absence': int(tr.absence),
                        'filtered': set(name for bit, name in _FILTERS if tr.filtered & bit),
                        'already_in_alarm': bool(tr.already_in_alarm)})
        return out

    def alarm_state(self):
        a, d = ct.c_int(0), ct.c_double(0)
        nat.check(self._L.gofpid_tracker_alarm_state(self._h, ct.byref(a), ct.byref(d)))
        return {'already_in_alarm': bool(a.value), 'duration': d.value}
