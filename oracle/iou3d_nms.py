"""ORACLE (test infrastructure): rotated BEV IoU and NMS -- SURVEY.md 8(f) N4.

Scalar float32 restatement of the reference's device code
  snvc/extension/iou3d_nms/src/iou3d_nms_kernel.cu:36-47   cross / check_rect_cross
                                                    :49-60   check_in_box2d (MARGIN 1e-2)
                                                    :62-93   intersection of two segments (EPS 1e-8)
                                                    :95-103  rotate_around_center, point_cmp (atan2 ordering)
                                                    :105-226 box_overlap (edge intersections + contained corners,
                                                             bubble sort around the centroid, fan area)
                                                    :228-235 iou_bev
  snvc/extension/iou3d_nms/src/iou3d_nms.cpp:131-177        nms_gpu: 64-bit suppression masks + greedy keep loop
  snvc/extension/iou3d_nms/iou3d_nms_utils.py:86-102        nms_gpu: sort by score, optional pre_maxsize
Boxes are [x, y, z, dx, dy, dz, heading].  PARITY UNPINNED against the reference itself: its op needs the torch
extension build and a GPU (no golden vectors are shipped); the restatement is cross-checked against an independent
float64 convex-polygon clipper (`iou_bev_exact`) in tests/test_oracle_iou3d_nms.py.
"""
import math

import numpy as np

F = np.float32
EPS = F(1e-8)
MARGIN = F(1e-2)


def _cross3(p1, p2, p0):
    return F(F(F(p1[0] - p0[0]) * F(p2[1] - p0[1])) - F(F(p2[0] - p0[0]) * F(p1[1] - p0[1])))


def _check_rect_cross(p1, p2, q1, q2):
    return (min(p1[0], p2[0]) <= max(q1[0], q2[0]) and min(q1[0], q2[0]) <= max(p1[0], p2[0]) and
            min(p1[1], p2[1]) <= max(q1[1], q2[1]) and min(q1[1], q2[1]) <= max(p1[1], p2[1]))


def _in_box(box, p):
    c, s = F(math.cos(-float(box[6]))), F(math.sin(-float(box[6])))
    rx = F(F(F(p[0] - box[0]) * c) + F(F(p[1] - box[1]) * F(-s)))
    ry = F(F(F(p[0] - box[0]) * s) + F(F(p[1] - box[1]) * c))
    return abs(rx) < F(F(box[3] / F(2)) + MARGIN) and abs(ry) < F(F(box[4] / F(2)) + MARGIN)


def _intersection(p1, p0, q1, q0):
    if not _check_rect_cross(p0, p1, q0, q1):
        return None
    s1, s2 = _cross3(q0, p1, p0), _cross3(p1, q1, p0)
    s3, s4 = _cross3(p0, q1, q0), _cross3(q1, p1, q0)
    if not (F(s1 * s2) > 0 and F(s3 * s4) > 0):
        return None
    s5 = _cross3(q1, p1, p0)
    if abs(F(s5 - s1)) > EPS:
        d = F(s5 - s1)
        return (F(F(F(s5 * q0[0]) - F(s1 * q1[0])) / d), F(F(F(s5 * q0[1]) - F(s1 * q1[1])) / d))
    a0, b0, c0 = F(p0[1] - p1[1]), F(p1[0] - p0[0]), F(F(p0[0] * p1[1]) - F(p1[0] * p0[1]))
    a1, b1, c1 = F(q0[1] - q1[1]), F(q1[0] - q0[0]), F(F(q0[0] * q1[1]) - F(q1[0] * q0[1]))
    D = F(F(a0 * b1) - F(a1 * b0))
    return (F(F(F(b0 * c1) - F(b1 * c0)) / D), F(F(F(a1 * c0) - F(a0 * c1)) / D))


def _corners(box):
    hx, hy = F(box[3] / F(2)), F(box[4] / F(2))
    pts = [(F(box[0] - hx), F(box[1] - hy)), (F(box[0] + hx), F(box[1] - hy)), (F(box[0] + hx), F(box[1] + hy)),
           (F(box[0] - hx), F(box[1] + hy))]
    c, s = F(math.cos(float(box[6]))), F(math.sin(float(box[6])))
    out = []
    for (x, y) in pts:
        nx = F(F(F(F(x - box[0]) * c) + F(F(y - box[1]) * F(-s))) + box[0])
        ny = F(F(F(F(x - box[0]) * s) + F(F(y - box[1]) * c)) + box[1])
        out.append((nx, ny))
    return out + [out[0]]


def box_overlap(box_a, box_b):
    box_a, box_b = np.asarray(box_a, F), np.asarray(box_b, F)
    ca, cb = _corners(box_a), _corners(box_b)
    pts, cx, cy = [], F(0), F(0)
    for i in range(4):
        for j in range(4):
            p = _intersection(ca[i + 1], ca[i], cb[j + 1], cb[j])
            if p is not None:
                cx, cy = F(cx + p[0]), F(cy + p[1])
                pts.append(p)
    for k in range(4):
        if _in_box(box_a, cb[k]):
            cx, cy = F(cx + cb[k][0]), F(cy + cb[k][1]); pts.append(cb[k])
        if _in_box(box_b, ca[k]):
            cx, cy = F(cx + ca[k][0]), F(cy + ca[k][1]); pts.append(ca[k])
    cnt = len(pts)
    if cnt == 0:
        return F(0)                       # (the reference divides 0/0 and sums an empty fan: area 0)
    cx, cy = F(cx / F(cnt)), F(cy / F(cnt))
    ang = lambda p: math.atan2(float(F(p[1] - cy)), float(F(p[0] - cx)))
    for j in range(cnt - 1):              # the reference's bubble sort, same comparison
        for i in range(cnt - j - 1):
            if F(ang(pts[i])) > F(ang(pts[i + 1])):
                pts[i], pts[i + 1] = pts[i + 1], pts[i]
    area = F(0)
    for k in range(cnt - 1):
        ax, ay = F(pts[k][0] - pts[0][0]), F(pts[k][1] - pts[0][1])
        bx, by = F(pts[k + 1][0] - pts[0][0]), F(pts[k + 1][1] - pts[0][1])
        area = F(area + F(F(ax * by) - F(ay * bx)))
    return F(abs(area) / F(2))


def iou_bev(box_a, box_b):
    box_a, box_b = np.asarray(box_a, F), np.asarray(box_b, F)
    sa, sb = F(box_a[3] * box_a[4]), F(box_b[3] * box_b[4])
    so = box_overlap(box_a, box_b)
    return F(so / max(F(F(sa + sb) - so), EPS))


def boxes_iou_bev(boxes_a, boxes_b):
    return np.array([[iou_bev(a, b) for b in boxes_b] for a in boxes_a], dtype=F)


def nms(boxes, scores, thresh, pre_maxsize=None):
    """iou3d_nms_utils.nms_gpu + iou3d_nms.cpp nms_gpu: indices (into the unsorted input) of the kept boxes."""
    order = np.argsort(-np.asarray(scores, F), kind="stable")
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    b = np.asarray(boxes, F)[order]
    removed = np.zeros(len(b), bool)
    keep = []
    for i in range(len(b)):
        if removed[i]:
            continue
        keep.append(i)
        for j in range(i + 1, len(b)):
            if not removed[j] and iou_bev(b[i], b[j]) > F(thresh):
                removed[j] = True
    return order[np.array(keep, dtype=np.int64)]


# ---- independent check: exact (float64) intersection area of two rotated rectangles, Sutherland-Hodgman clipping
def _rect64(box):
    x, y, dx, dy, r = (float(box[0]), float(box[1]), float(box[3]), float(box[4]), float(box[6]))
    c, s = math.cos(r), math.sin(r)
    return [(x + c * px - s * py, y + s * px + c * py) for px, py in ((-dx / 2, -dy / 2), (dx / 2, -dy / 2), (dx / 2, dy / 2), (-dx / 2, dy / 2))]


def iou_bev_exact(box_a, box_b):
    poly, clip = _rect64(box_a), _rect64(box_b)
    for i in range(4):
        a, b = clip[i], clip[(i + 1) % 4]
        inside = lambda p: (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) >= 0
        out = []
        for k in range(len(poly)):
            p, q = poly[k], poly[(k + 1) % len(poly)]
            if inside(p) != inside(q):
                t = ((b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])) / \
                    ((b[1] - a[1]) * (q[0] - p[0]) - (b[0] - a[0]) * (q[1] - p[1]))
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
            if inside(q):
                out.append(q)
        poly = out
        if not poly:
            return 0.0
    area = 0.5 * abs(sum(poly[k][0] * poly[(k + 1) % len(poly)][1] - poly[(k + 1) % len(poly)][0] * poly[k][1] for k in range(len(poly))))
    sa, sb = float(box_a[3]) * float(box_a[4]), float(box_b[3]) * float(box_b[4])
    return area / max(sa + sb - area, 1e-8)


def synthetic_boxes(n=96, seed=3, extent=20.0):
    """Car-sized boxes scattered (and clustered) in a BEV patch, random headings; scores in (0,1)."""
    rng = np.random.RandomState(seed)
    centres = rng.uniform(-extent, extent, size=(n // 3 + 1, 2))
    idx = rng.randint(0, len(centres), size=n)
    xy = centres[idx] + rng.normal(0, 0.6, size=(n, 2))
    boxes = np.stack([xy[:, 0], xy[:, 1], rng.uniform(-1, 1, n), rng.uniform(3.4, 4.6, n), rng.uniform(1.5, 2.0, n),
                      rng.uniform(1.4, 1.8, n), rng.uniform(-np.pi, np.pi, n)], axis=1).astype(F)
    return boxes, rng.uniform(0.05, 1.0, n).astype(F)
