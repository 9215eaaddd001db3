// tracker.cu -- the host tail of GOFPID.detect, native: blob tracking, post-filtering and the alarm.
//
// Replaces, per frame, steps 5-7 of the reference (pygofpid/pipeline.py):
//   :590-628  _track_blob            greedy nearest-track association in blob order
//   :630-733  _create_tracked_blob, _compute_tracking_distance, _compute_tracking_space,
//             _update_paired_tracked_blob, _compute_total_distance, _update_unpaired_tracked_blob
//   :734-762  _post_filter           perimeter (cv.pointPolygonTest) and perspective (cv.boundingRect) filters
//   :764-780  _detect_blob           a track without active filters raises the alarm
//   :782-799  _maintain_alarm        hold a raised alarm for duration_min frames
//
// Pure host code (no CUDA call): a few dozen scalars per blob and per track, which in Python cost
// 360 us per frame at 12 blobs -- an order of magnitude more than the whole pixel path of the frame.
// The arithmetic keeps the reference's data types, because they decide ties and wrap-arounds:
// centres / bottoms are int16, their differences and squares are int16 too (NumPy scalar arithmetic
// wraps), the radii of the search ellipse are int32, quotients and sums are float64.
#include "common.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

using namespace gofpid;

namespace {

enum { F_PRESENCE = 1, F_DISTANCE = 2, F_PERIMETER = 4, F_PERSPECTIVE = 8 };

struct BlobIn { // one instantaneous blob (an entry of the reference's blobs_)
    int16_t anchor[2], bottom[2], center[2];
    int x0, y0, x1, y1;   // cv.boundingRect of the contour (inclusive)
    const int32_t *pts;   // contour points (x, y)
    int n_points;
};

struct Track {
    int16_t anchor[2], bottom[2], center[2], center_first[2];
    int presence, absence;
    unsigned filtered;
    int already_in_alarm;
    int x0, y0, x1, y1;
    std::vector<int32_t> contour;
};

} // namespace

struct gofpid_tracker {
    int H = 0;
    std::vector<double> persp_height, persp_area;
    std::vector<double> perimeter; // (x, y) as doubles
    double factor = 1.5;
    int absence_max = 10;
    int anchor_center = 0;
    double perspective_coeff = 0.5, presence_min = 3, distance_min = 0.25;
    int keep_alarm = 1;
    double duration_min = 10;
    std::vector<Track> tracks;
    int sys_already_in_alarm = 0;
    double duration = 0;
    // scratch of a step: one row of the distance matrix and what the distance needs from every track
    std::vector<double> dist, a2, b2;
    std::vector<int16_t> tcx, tcy;
    std::vector<uint8_t> zero, col_neg;
    std::vector<BlobIn> blobs; // scratch
};

namespace {

inline int clip_row(const gofpid_tracker *t, int y) { return y < 0 ? 0 : (y > t->H - 1 ? t->H - 1 : y); }

// cv.pointPolygonTest(polygon, point, False) as pygofpid_b200/helpers.py:point_in_polygon states it
int point_in_polygon(const std::vector<double> &poly, double px, double py)
{
    const int n = (int)(poly.size() / 2);
    bool inside = false;
    for (int i = 0; i < n; i++) {
        const int h = i == 0 ? n - 1 : i - 1;
        const double x0 = poly[2 * h], y0 = poly[2 * h + 1], x1 = poly[2 * i], y1 = poly[2 * i + 1];
        const double cross = (px - x0) * (y1 - y0) - (py - y0) * (x1 - x0);
        if (cross == 0 && std::min(x0, x1) <= px && px <= std::max(x0, x1) && std::min(y0, y1) <= py && py <= std::max(y0, y1))
            return 0;
        if ((y0 > py) != (y1 > py)) {
            const double xi = x0 + (py - y0) * (x1 - x0) / (y1 - y0);
            if (xi > px) inside = !inside;
        }
    }
    return inside ? 1 : -1;
}

// (the distance of a (track, blob) pair, pipeline.py:659-689, is evaluated row by row in step())
Track new_track(const BlobIn &b)
{
    Track tr;
    memcpy(tr.anchor, b.anchor, 4); memcpy(tr.bottom, b.bottom, 4); memcpy(tr.center, b.center, 4);
    memcpy(tr.center_first, b.center, 4);
    tr.presence = 1; tr.absence = 0; tr.filtered = F_PRESENCE | F_DISTANCE; tr.already_in_alarm = 0;
    tr.x0 = b.x0; tr.y0 = b.y0; tr.x1 = b.x1; tr.y1 = b.y1;
    tr.contour.assign(b.pts, b.pts + (size_t)b.n_points * 2);
    return tr;
}

void pair(const gofpid_tracker *t, Track &tr, const BlobIn &b)
{
    memcpy(tr.anchor, b.anchor, 4); memcpy(tr.bottom, b.bottom, 4); memcpy(tr.center, b.center, 4);
    tr.x0 = b.x0; tr.y0 = b.y0; tr.x1 = b.x1; tr.y1 = b.y1;
    tr.contour.assign(b.pts, b.pts + (size_t)b.n_points * 2);
    tr.absence = 0;
    tr.presence += 1;
    if ((double)tr.presence >= t->presence_min) tr.filtered &= ~(unsigned)F_PRESENCE;
    // _compute_total_distance: |centre - first centre| over sqrt(expected person area at the bottom row)
    const int16_t dx = (int16_t)(tr.center[0] - tr.center_first[0]), dy = (int16_t)(tr.center[1] - tr.center_first[1]);
    const double d = sqrt((double)dx * (double)dx + (double)dy * (double)dy);
    const double area_min = t->persp_area[clip_row(t, tr.bottom[1])];
    const double rel = area_min <= 0 ? 0.0 : d / sqrt(area_min);
    if (rel >= t->distance_min) tr.filtered &= ~(unsigned)F_DISTANCE;
}

void miss(gofpid_tracker *t, int j)
{
    Track &tr = t->tracks[j];
    tr.presence = 0;
    tr.filtered |= F_PRESENCE;
    tr.absence += 1;
    if (tr.absence >= t->absence_max) t->tracks.erase(t->tracks.begin() + j);
}

// steps 5-7 for one frame
int step(gofpid_tracker *t, const BlobIn *blobs, int nb)
{
    std::vector<Track> &T = t->tracks;
    const int nt = (int)T.size();
    // ---- step 5: tracking ------------------------------------------------------------------
    if (nt == 0) {
        for (int i = 0; i < nb; i++) T.push_back(new_track(blobs[i]));
    } else if (nb == 0) {
        for (int j = nt - 1; j >= 0; j--) miss(t, j);
    } else {
        // The reference fills an (n_blobs x n_tracks) matrix up front, walks the blobs (argmin of the blob's
        // row; a pairing writes -1 into that one entry) and finally ages every track whose COLUMN has no
        // negative / NaN entry.  Rows never see each other's -1, so the matrix is not kept: a row at a time
        // over a structure-of-arrays snapshot of the tracks (pair() changes them while the blobs are walked)
        // -- a plain loop the compiler vectorises -- plus one "column has a negative entry" flag per track.
        t->dist.resize(nt); t->a2.resize(nt); t->b2.resize(nt); t->tcx.resize(nt); t->tcy.resize(nt);
        t->zero.resize(nt); t->col_neg.assign(nt, 0);
        double *row = t->dist.data(), *A2 = t->a2.data(), *B2 = t->b2.data();
        int16_t *tcx = t->tcx.data(), *tcy = t->tcy.data();
        uint8_t *zero = t->zero.data(), *col_neg = t->col_neg.data();
        for (int j = 0; j < nt; j++) {
            const double height_min = t->persp_height[clip_row(t, T[j].bottom[1])];
            const int32_t rmaj = (int32_t)(height_min * t->factor);   // np.int32(...): truncation
            const int32_t rmin = (int32_t)((double)rmaj / 2);
            zero[j] = rmaj <= 0 || rmin <= 0;                          // the reference's distance is 0 there
            A2[j] = zero[j] ? 1.0 : (double)(int32_t)((uint32_t)rmaj * (uint32_t)rmaj);
            B2[j] = zero[j] ? 1.0 : (double)(int32_t)((uint32_t)rmin * (uint32_t)rmin);
            tcx[j] = T[j].center[0]; tcy[j] = T[j].center[1];
        }
        for (int i = 0; i < nb; i++) {
            const int bcx = blobs[i].center[0], bcy = blobs[i].center[1];
            for (int j = 0; j < nt; j++) { // tracking_distance(), int16 differences and squares wrapping as NumPy's do
                const int16_t dx = (int16_t)(bcx - tcx[j]), dy = (int16_t)(bcy - tcy[j]);
                const int16_t dx2 = (int16_t)((int)dx * (int)dx), dy2 = (int16_t)((int)dy * (int)dy);
                const double d = (double)dx2 / A2[j] + (double)dy2 / B2[j];
                row[j] = zero[j] ? 0.0 : d;
            }
            for (int j = 0; j < nt; j++) col_neg[j] |= (uint8_t)!(row[j] >= 0); // (a wrapped negative or a NaN entry)
            // np.argmin: first minimum; a NaN counts as the minimum
            int jm = 0;
            if (!isnan(row[0]))
                for (int j = 1; j < nt; j++) {
                    if (isnan(row[j])) { jm = j; break; }
                    if (row[j] < row[jm]) jm = j;
                }
            if (row[jm] <= 1) {
                col_neg[jm] = 1; // (the -1 the reference writes)
                pair(t, T[jm], blobs[i]);
            } else {
                T.push_back(new_track(blobs[i])); // (T may reallocate: no reference into it is kept)
            }
        }
        for (int j = nt - 1; j >= 0; j--)
            if (!col_neg[j]) miss(t, j); // np.all(dist[:, j] >= 0)
    }
    // ---- step 6: post-filter ---------------------------------------------------------------
    for (Track &tr : T) {
        if (point_in_polygon(t->perimeter, (double)tr.anchor[0], (double)tr.anchor[1]) < 0) tr.filtered |= F_PERIMETER;
        else tr.filtered &= ~(unsigned)F_PERIMETER;
        const long area = (long)(tr.x1 - tr.x0 + 1) * (long)(tr.y1 - tr.y0 + 1);
        const double area_persp = t->persp_area[clip_row(t, tr.bottom[1])];
        if ((double)area <= area_persp * t->perspective_coeff) tr.filtered |= F_PERSPECTIVE;
        else tr.filtered &= ~(unsigned)F_PERSPECTIVE;
    }
    // ---- step 7: alarm -----------------------------------------------------------------------
    int alarm = 0;
    for (Track &tr : T) {
        if (tr.filtered == 0) {
            alarm = 1;
            if (t->keep_alarm) tr.already_in_alarm = 1;
        } else if (tr.already_in_alarm) {
            alarm = 1;
        }
    }
    t->duration -= 1;
    if (alarm && !t->sys_already_in_alarm) t->duration = t->duration_min;
    if (!alarm && t->duration >= 0) alarm = 1;
    t->sys_already_in_alarm = alarm;
    return alarm;
}

// the device-side half of GOFPID._create_blob lives in gofpid_blob (Green sums); this is the host half
// (pygofpid_b200/hotpath.py:_blob_dict): m00 = a00/2, m10 = a10/6, m01 = a01/6 with the sign of a00
// folded in, centre = (int(m10/m00), int(m01/m00)), bottom = (centre x, largest y)
int blob_from_record(const gofpid_tracker *t, const gofpid_blob &r, const int32_t *points, BlobIn *b)
{
    const double a00 = (double)r.a00, a10 = (double)r.a10, a01 = (double)r.a01;
    const double h = a00 > 0 ? 0.5 : -0.5, s = a00 > 0 ? 0.16666666666666666666666666666667 : -0.16666666666666666666666666666667;
    const double m00 = a00 * h, m10 = a10 * s, m01 = a01 * s;
    if (m00 == 0) { set_error("a contour of zero area has no centre (the reference divides by m00 there)"); return GOFPID_ERR_INVALID; }
    const double cx = trunc(m10 / m00), cy = trunc(m01 / m00);
    b->center[0] = (int16_t)cx; b->center[1] = (int16_t)cy;
    b->bottom[0] = (int16_t)cx; b->bottom[1] = (int16_t)r.y1;
    const int16_t *a = t->anchor_center ? b->center : b->bottom;
    b->anchor[0] = a[0]; b->anchor[1] = a[1];
    b->x0 = r.x0; b->y0 = r.y0; b->x1 = r.x1; b->y1 = r.y1;
    b->pts = points + (size_t)r.point_offset * 2;
    b->n_points = r.n_points;
    return GOFPID_OK;
}

} // namespace

extern "C" {

GOFPID_API int gofpid_tracker_create(const gofpid_tracker_config *cfg, gofpid_tracker **out)
{
    if (!cfg || !out) { set_error("null argument"); return GOFPID_ERR_INVALID; }
    *out = nullptr;
    if (cfg->H < 1 || !cfg->persp_height || !cfg->persp_area || cfg->n_perimeter < 0 || (cfg->n_perimeter && !cfg->perimeter)) {
        set_error("tracker needs the frame height, the two perspective tables and the perimeter");
        return GOFPID_ERR_INVALID;
    }
    gofpid_tracker *t = new (std::nothrow) gofpid_tracker();
    if (!t) { set_error("out of host memory"); return GOFPID_ERR_NOMEM; }
    t->H = cfg->H;
    t->persp_height.assign(cfg->persp_height, cfg->persp_height + cfg->H);
    t->persp_area.assign(cfg->persp_area, cfg->persp_area + cfg->H);
    t->perimeter.resize((size_t)cfg->n_perimeter * 2);
    for (int i = 0; i < cfg->n_perimeter * 2; i++) t->perimeter[i] = (double)cfg->perimeter[i];
    t->factor = cfg->tracking_factor; t->absence_max = cfg->absence_max; t->anchor_center = cfg->anchor_center;
    t->perspective_coeff = cfg->perspective_coeff; t->presence_min = cfg->presence_min; t->distance_min = cfg->distance_min;
    t->keep_alarm = cfg->keep_alarm; t->duration_min = cfg->duration_min;
    *out = t;
    return GOFPID_OK;
}

GOFPID_API int gofpid_tracker_destroy(gofpid_tracker *t)
{
    delete t;
    return GOFPID_OK;
}

GOFPID_API int gofpid_tracker_reset(gofpid_tracker *t)
{
    if (!t) { set_error("null tracker"); return GOFPID_ERR_INVALID; }
    t->tracks.clear(); t->sys_already_in_alarm = 0; t->duration = 0;
    return GOFPID_OK;
}

GOFPID_API int gofpid_tracker_step(gofpid_tracker *t, const gofpid_blob *h_blobs, int n_blobs, const int32_t *h_points,
                                   int *alarm)
{
    if (!t || n_blobs < 0 || (n_blobs && (!h_blobs || !h_points))) { set_error("bad argument"); return GOFPID_ERR_INVALID; }
    t->blobs.resize((size_t)n_blobs);
    for (int i = 0; i < n_blobs; i++) GOFPID_TRY(blob_from_record(t, h_blobs[i], h_points, &t->blobs[i]));
    const int a = step(t, t->blobs.data(), n_blobs);
    if (alarm) *alarm = a;
    return GOFPID_OK;
}

GOFPID_API int gofpid_tracker_step_prepared(gofpid_tracker *t, int n_blobs, const int16_t *h_anchor, const int16_t *h_bottom,
                                            const int16_t *h_center, const int32_t *h_boxes, const int32_t *h_points,
                                            const int32_t *h_offsets, int *alarm)
{
    if (!t || n_blobs < 0 || (n_blobs && (!h_anchor || !h_bottom || !h_center || !h_boxes || !h_points || !h_offsets))) {
        set_error("bad argument");
        return GOFPID_ERR_INVALID;
    }
    t->blobs.resize((size_t)n_blobs);
    for (int i = 0; i < n_blobs; i++) {
        BlobIn &b = t->blobs[i];
        for (int k = 0; k < 2; k++) { b.anchor[k] = h_anchor[2 * i + k]; b.bottom[k] = h_bottom[2 * i + k]; b.center[k] = h_center[2 * i + k]; }
        b.x0 = h_boxes[4 * i]; b.y0 = h_boxes[4 * i + 1]; b.x1 = h_boxes[4 * i + 2]; b.y1 = h_boxes[4 * i + 3];
        b.pts = h_points + (size_t)h_offsets[i] * 2;
        b.n_points = h_offsets[i + 1] - h_offsets[i];
    }
    const int a = step(t, t->blobs.data(), n_blobs);
    if (alarm) *alarm = a;
    return GOFPID_OK;
}

GOFPID_API int gofpid_tracker_run(gofpid_tracker *t, const gofpid_blob *h_blobs, const int32_t *h_points,
                                  const int32_t *h_counts, int n_frames, int n_streams, int stream_idx, uint8_t *h_alarms)
{
    if (!t || n_frames < 0 || n_streams < 1 || stream_idx < 0 || stream_idx >= n_streams || !h_counts || !h_alarms) {
        set_error("bad argument");
        return GOFPID_ERR_INVALID;
    }
    size_t ob = 0, op = 0;
    for (int f = 0; f < n_frames; f++)
        for (int z = 0; z < n_streams; z++) {
            const int nb = h_counts[((size_t)f * n_streams + z) * 2], np = h_counts[((size_t)f * n_streams + z) * 2 + 1];
            if (z == stream_idx) {
                int a = 0;
                GOFPID_TRY(gofpid_tracker_step(t, h_blobs ? h_blobs + ob : nullptr, nb, h_points ? h_points + op * 2 : nullptr, &a));
                h_alarms[f] = (uint8_t)a;
            }
            ob += (size_t)nb; op += (size_t)np;
        }
    return GOFPID_OK;
}

GOFPID_API int gofpid_tracker_count(gofpid_tracker *t)
{
    return t ? (int)t->tracks.size() : 0;
}

GOFPID_API int gofpid_tracker_get(gofpid_tracker *t, int i, gofpid_track *out)
{
    if (!t || !out || i < 0 || i >= (int)t->tracks.size()) { set_error("no such track"); return GOFPID_ERR_INVALID; }
    const Track &tr = t->tracks[i];
    memcpy(out->anchor, tr.anchor, 4); memcpy(out->bottom, tr.bottom, 4); memcpy(out->center, tr.center, 4);
    memcpy(out->center_first, tr.center_first, 4);
    out->presence = tr.presence; out->absence = tr.absence; out->filtered = (int32_t)tr.filtered;
    out->already_in_alarm = tr.already_in_alarm;
    out->x0 = tr.x0; out->y0 = tr.y0; out->x1 = tr.x1; out->y1 = tr.y1;
    out->n_points = (int32_t)(tr.contour.size() / 2);
    return GOFPID_OK;
}

GOFPID_API int gofpid_tracker_contour(gofpid_tracker *t, int i, int32_t *h_points, int point_capacity)
{
    if (!t || !h_points || i < 0 || i >= (int)t->tracks.size()) { set_error("no such track"); return GOFPID_ERR_INVALID; }
    const Track &tr = t->tracks[i];
    if ((size_t)std::max(point_capacity, 0) * 2 < tr.contour.size()) { set_error("output capacity too small: need %zu points", tr.contour.size() / 2); return GOFPID_ERR_INVALID; }
    if (!tr.contour.empty()) memcpy(h_points, tr.contour.data(), tr.contour.size() * sizeof(int32_t));
    return GOFPID_OK;
}

GOFPID_API int gofpid_tracker_alarm_state(gofpid_tracker *t, int *already_in_alarm, double *duration)
{
    if (!t) { set_error("null tracker"); return GOFPID_ERR_INVALID; }
    if (already_in_alarm) *already_in_alarm = t->sys_already_in_alarm;
    if (duration) *duration = t->duration;
    return GOFPID_OK;
}

} // extern "C"
