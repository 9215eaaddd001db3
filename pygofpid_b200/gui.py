"""Interactive configuration of perimeter / perspective (reference pipeline.py:308-407).

These are mouse-driven OpenCV windows outside the per-frame hot path.  They need an OpenCV
build with GUI support; headless builds (as in this project's image) cannot open windows, in
which case a clear error is raised instead of silently skipping the configuration."""
import numpy as np

from .helpers import find_line, find_point, normalize_coords, plot_lines, plot_rectangles, unnormalize_coords


def _gui():
    try:
        import cv2 as cv
    except Exception as e:  # pragma: no cover
        raise NotImplementedError("display_config needs OpenCV with GUI support") from e
    if "GUI:                           NONE" in cv.getBuildInformation():
        raise NotImplementedError("display_config needs an OpenCV build with GUI support "
                                  "(this one is headless)")
    return cv


def configure_perimeter(gofpid, window_name):
    cv = _gui()
    img, clone, thickness = gofpid._get_config_img()
    points = unnormalize_coords(gofpid.post_filter['perimeter'], img.shape).tolist()
    state = {'point': -1, 'line': -1}

    def on_mouse(event, x, y, flags, params):
        if event == cv.EVENT_LBUTTONDOWN:
            state['point'] = find_point((x, y), points, thickness)
        elif event == cv.EVENT_LBUTTONUP and state['point'] >= 0:
            points[state['point']] = [x, y]
        elif event == cv.EVENT_RBUTTONDOWN:
            state['point'] = find_point((x, y), points, thickness)
            state['line'] = find_line((x, y), points, thickness)
        elif event == cv.EVENT_RBUTTONUP:
            if state['point'] >= 0:
                points.pop(state['point'])
            elif state['line'] >= 0:
                points.insert(state['line'] + 1, [x, y])

    cv.namedWindow(window_name)
    cv.setMouseCallback(window_name, on_mouse)
    while True:
        img = clone.copy()
        plot_lines(img, points, thickness)
        cv.imshow(window_name, img)
        key = cv.waitKey(1) & 0xFF
        if key == ord("r"):
            points[:] = unnormalize_coords(np.array([[0, 0], [0, 1], [1, 1], [1, 0]]), img.shape).tolist()
        elif key == ord("c"):
            cv.destroyWindow(window_name)
            break
    return normalize_coords(points, img.shape)


def configure_perspective(gofpid, window_name):
    cv = _gui()
    img, clone, thickness = gofpid._get_config_img()
    points = unnormalize_coords(gofpid.post_filter['perspective'], img.shape, dtype=np.int32)
    state = {'point': -1}

    def on_mouse(event, x, y, flags, params):
        if event == cv.EVENT_LBUTTONDOWN:
            state['point'] = find_point((x, y), points, thickness)
        elif event == cv.EVENT_LBUTTONUP and state['point'] >= 0:
            points[state['point']] = [x, y]

    cv.namedWindow(window_name)
    cv.setMouseCallback(window_name, on_mouse)
    while True:
        img = clone.copy()
        plot_rectangles(img, points, thickness)
        cv.imshow(window_name, img)
        if (cv.waitKey(1) & 0xFF) == ord("c"):
            cv.destroyWindow(window_name)
            break
    return normalize_coords(points, img.shape)
