"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize_case.py

Every kernel family of the library on tiny frames: MOG2 gray + BGR through the clip entry point (frame
groups: mbarriers, bulk copies, the shared-memory queue), KNN gray + BGR, FD, bit-plane and byte morphology,
run-based labelling + border tracing + nesting, the tiled union-find fallback, and the big re-run of an
overflowed frame.  Results are checked against the CPU oracle as in the tests."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import blobs as oblobs
from oracle import cport
from pygofpid_b200 import HotPath
from pygofpid_b200.synth import SyntheticClip

r3 = np.ones((3, 3), np.uint8)
ell = np.array([[0, 1, 1, 1, 0], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [0, 1, 1, 1, 0]], np.uint8)


def check_blobs(got, want, ctx):
    assert len(got) == len(want), (ctx, len(got), len(want))
    for g, w in zip(got, want):
        assert np.array_equal(g["contour"], w["contour"]), ctx


def mog2(C, H, W, G):
    hp = HotPath((H, W) if C == 1 else (H, W, C), detector="MOG2", morph=[("erode", r3), ("dilate", ell)], area_min=10.0,
                 depth=6, mog2_group=G)
    ora = cport.MOG2(H, W, C)
    clip = SyntheticClip(H, W, C, seed=3 + C, n_obj=3)
    frames = np.stack(clip.frames(6 + 3 * G))
    hp.process_sequence(frames)
    for f in frames:
        want = cport.dilate(cport.erode(ora.apply(cport.gaussian_blur(f, (3, 3), 4.0)), r3), ell)
    assert np.array_equal(hp.mask(), want), ("mog2", C, H, W)
    check_blobs(hp.blobs(), oblobs.blobs_from_mask(want, area_min=10.0), ("mog2", C))
    hp.close()


def knn(C, H, W):
    hp = HotPath((H, W) if C == 1 else (H, W, C), detector="KNN", morph=[("erode", r3), ("dilate", r3)], area_min=10.0,
                 knn_seed=99, depth=3)
    ora = cport.KNN(H, W, C, seed=99)
    clip = SyntheticClip(H, W, C, seed=8 + C, n_obj=3)
    for t in range(8):
        f = clip.next()
        hp.process(f)
        want = cport.dilate(cport.erode(ora.apply(cport.gaussian_blur(f, (3, 3), 4.0)), r3), r3)
    assert np.array_equal(hp.mask(), want), ("knn", C)
    hp.close()


def fd_and_blobs(H, W, **caps):
    rs = np.random.RandomState(1)
    hp = HotPath((H, W), blur=None, detector="FD", find_blobs=True, fd_threshold=0, area_min=4.0, depth=2, **caps)
    prev = np.zeros((H, W), np.uint8)
    hp.process(prev)
    for k in range(3):
        M = ((rs.rand(H, W) < (0.05, 0.45, 0.2)[k]) * 255).astype(np.uint8)
        M[10:30, 20:60] = 255
        M[15:25, 30:50] = 0
        M[18:22, 35:45] = 255  # a blob inside a hole
        frame = np.where(M > 0, 255 - prev, prev).astype(np.uint8)
        hp.process(frame)
        prev = frame
        assert np.array_equal(hp.mask(), M)
        assert np.array_equal(hp.labels(), cport.ccl8(M))
        check_blobs(hp.blobs(), oblobs.blobs_from_mask(M, area_min=4.0), ("fd", k, caps))
    hp.close()


if __name__ == "__main__":
    mog2(1, 40, 136, 3)    # border tiles
    mog2(1, 48, 256, 2)    # whole 64x8 tiles: the FULL instantiation, interior tiles blur in registers
    mog2(3, 40, 100, 3)
    knn(1, 40, 136)
    knn(3, 24, 72)
    fd_and_blobs(48, 96)
    fd_and_blobs(48, 96, max_blobs=32, max_kept=4, max_points=64)  # overflow -> re-run with larger buffers
    os.environ["GOFPID_RUNS_CAP"] = "16"                           # tiled union-find fallback
    fd_and_blobs(48, 96)
    print("sanitize_case ok")
