"""Pure-Python restatement of blob creation (TEST INFRASTRUCTURE ONLY, small cases).

Reference: GOFPID._find_blob / _create_blob (pipeline.py:555-588) and helpers.find_contours /
get_bottom / get_center (helpers.py:125-168), i.e. cv.findContours(RETR_EXTERNAL,
CHAIN_APPROX_NONE) + cv.contourArea + cv.moments.  Restated from the published
Suzuki-Abe border following algorithm that OpenCV implements.
"""
import numpy as np

from . import cport

# direction codes: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE  (y grows downwards)
_DX = (1, 1, 0, -1, -1, -1, 0, 1)
_DY = (0, -1, -1, -1, 0, 1, 1, 1)


def trace_outer_border(mask, x0, y0):
    """Outer border of the 8-connected component whose raster-first pixel is (x0, y0).

    Returns the list of (x, y) points in OpenCV's order (CHAIN_APPROX_NONE).
    """
    H, W = mask.shape

    def fg(x, y):
        return 0 <= x < W and 0 <= y < H and mask[y, x] != 0

    # look clockwise, starting just after "west", for the first foreground neighbour
    s = 4
    while True:
        s = (s - 1) & 7
        if fg(x0 + _DX[s], y0 + _DY[s]) or s == 4:
            break
    if s == 4 and not fg(x0 + _DX[4], y0 + _DY[4]):
        return [(x0, y0)]
    x1, y1 = x0 + _DX[s], y0 + _DY[s]
    pts = []
    cx, cy = x0, y0
    while True:
        # counter-clockwise search starting after the direction we came from
        while True:
            s = (s + 1) & 7
            nx, ny = cx + _DX[s], cy + _DY[s]
            if fg(nx, ny):
                break
        pts.append((cx, cy))
        if (nx, ny) == (x0, y0) and (cx, cy) == (x1, y1):
            break
        cx, cy = nx, ny
        s = (s + 4) & 7
    return pts


def contour_area(pts):
    """cv.contourArea for an integer polygon: |sum(x_prev*y - x*y_prev)| / 2 (pipeline.py:564)."""
    a = 0
    xp, yp = pts[-1]
    for x, y in pts:
        a += xp * y - x * yp
        xp, yp = x, y
    return abs(a) * 0.5


def contour_moments(pts):
    """m00, m10, m01 of cv.moments(contour) (helpers.py:156,165), same operation order."""
    a00 = a10 = a01 = 0.0
    xi_1, yi_1 = float(pts[-1][0]), float(pts[-1][1])
    for x, y in pts:
        xi, yi = float(x), float(y)
        dxy = xi_1 * yi - xi * yi_1
        a00 += dxy
        a10 += dxy * (xi_1 + xi)
        a01 += dxy * (yi_1 + yi)
        xi_1, yi_1 = xi, yi
    if abs(a00) > 1.1920928955078125e-07:  # FLT_EPSILON
        if a00 > 0:
            db1_2, db1_6 = 0.5, 0.16666666666666666666666666666667
        else:
            db1_2, db1_6 = -0.5, -0.16666666666666666666666666666667
        return a00 * db1_2, a10 * db1_6, a01 * db1_6
    return 0.0, 0.0, 0.0


def outside_background(mask):
    """Boolean map of the 4-connected background region reachable from outside the frame."""
    H, W = mask.shape
    pad = np.zeros((H + 2, W + 2), bool)
    pad[1:-1, 1:-1] = mask != 0
    out = np.zeros_like(pad)
    stack = [(0, 0)]
    out[0, 0] = True
    while stack:
        y, x = stack.pop()
        for dy, dx in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            yy, xx = y + dy, x + dx
            if 0 <= yy < H + 2 and 0 <= xx < W + 2 and not pad[yy, xx] and not out[yy, xx]:
                out[yy, xx] = True
                stack.append((yy, xx))
    return out[1:-1, 1:-1], out


def find_external_contours(mask):
    """cv.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_NONE)[0] restated (helpers.py:139-143).

    One outer border per 8-connected component that is not enclosed by another component,
    starting at the component's raster-first pixel, list in descending raster order of that
    pixel.
    """
    mask = np.asarray(mask)
    H, W = mask.shape
    labels = cport.ccl8(mask)
    _, outpad = outside_background(mask)
    firsts = np.unique(labels[labels > 0]) - 1
    contours = []
    for f in sorted(firsts.tolist(), reverse=True):
        y0, x0 = divmod(f, W)
        if not outpad[y0 + 1, x0]:  # west neighbour (padded coords) is not outside => nested
            continue
        pts = trace_outer_border(mask, x0, y0)
        contours.append(np.array(pts, np.int32).reshape(-1, 1, 2))
    return contours


def blobs_from_mask(mask, area_min=100, anchor="bottom"):
    """GOFPID._find_blob + _create_blob restated: list of dicts contour/anchor/bottom/center."""
    blobs = []
    for c in find_external_contours(mask):
        pts = [tuple(p) for p in c.reshape(-1, 2).tolist()]
        if contour_area(pts) < area_min:
            continue
        m00, m10, m01 = contour_moments(pts)
        cx = int(m10 / m00)
        bottom = np.array([cx, max(p[1] for p in pts)], dtype=np.int16)
        center = np.array([cx, int(m01 / m00)], dtype=np.int16)
        blobs.append(dict(contour=c, anchor=bottom if anchor == "bottom" else center,
                          bottom=bottom, center=center))
    return blobs
