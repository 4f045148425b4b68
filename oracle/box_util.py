"""TEST INFRASTRUCTURE (oracle): the `box_util` module the reference imports but does not ship
(sunrgbd_detection/roi_seg_box3d_dataset.py:15, box_pc_fit_dataset.py:17, train_boxpc.py:28, evaluate.py:20).
It is train/box_util.py of charlesq34/frustum-pointnets (the code base the reference was forked from; unpinned);
its published algorithm is restated here in numpy float64:

  polygon_clip          Sutherland-Hodgman clipping of a convex polygon by a convex polygon
  poly_area             shoelace formula
  convex_hull_intersection   clip + area (upstream: scipy ConvexHull(...).volume, which for the convex clip result is its
                        area; a clip result with fewer than 3 vertices, where upstream's qhull call raises, counts as 0)
  box3d_vol, box3d_iou  bird's-eye-view IoU of the x/z rectangles (corners 3,2,1,0) and the 3D IoU through the y overlap

and the callers' helpers get_3d_box / class2angle / class2size / compute_box3d_iou of roi_seg_box3d_dataset.py:64-139.
Pinned by closed forms in tests/test_oracle_cpu.py (axis-aligned boxes, 90-degree rotations, containment, symmetry) and
by a Monte-Carlo volume estimate."""
import numpy as np

from transferable3d_b200.constants import MEAN_DIMS_ARR, NUM_HEADING_BIN


def polygon_clip(subjectPolygon, clipPolygon):
    """Clip a polygon with another (convex) polygon; lists of (x, y); returns None when the result is empty."""
    def inside(p):
        return (cp2[0] - cp1[0]) * (p[1] - cp1[1]) > (cp2[1] - cp1[1]) * (p[0] - cp1[0])

    def computeIntersection():
        dc = [cp1[0] - cp2[0], cp1[1] - cp2[1]]
        dp = [s[0] - e[0], s[1] - e[1]]
        n1 = cp1[0] * cp2[1] - cp1[1] * cp2[0]
        n2 = s[0] * e[1] - s[1] * e[0]
        n3 = 1.0 / (dc[0] * dp[1] - dc[1] * dp[0])
        return [(n1 * dp[0] - n2 * dc[0]) * n3, (n1 * dp[1] - n2 * dc[1]) * n3]

    outputList = subjectPolygon
    cp1 = clipPolygon[-1]
    for clipVertex in clipPolygon:
        cp2 = clipVertex
        inputList = outputList
        outputList = []
        s = inputList[-1]
        for subjectVertex in inputList:
            e = subjectVertex
            if inside(e):
                if not inside(s):
                    outputList.append(computeIntersection())
                outputList.append(e)
            elif inside(s):
                outputList.append(computeIntersection())
            s = e
        cp1 = cp2
        if len(outputList) == 0:
            return None
    return outputList


def poly_area(x, y):
    return 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))


def convex_hull_intersection(p1, p2):
    inter_p = polygon_clip(p1, p2)
    if inter_p is None or len(inter_p) < 3:
        return None, 0.0
    a = np.asarray(inter_p, dtype=np.float64)
    return inter_p, float(poly_area(a[:, 0], a[:, 1]))


def box3d_vol(corners):
    a = np.sqrt(np.sum((corners[0, :] - corners[1, :]) ** 2))
    b = np.sqrt(np.sum((corners[1, :] - corners[2, :]) ** 2))
    c = np.sqrt(np.sum((corners[0, :] - corners[4, :]) ** 2))
    return a * b * c


def box3d_iou(corners1, corners2):
    """corners: (8,3) as produced by get_3d_box -> (iou_3d, iou_2d)."""
    corners1, corners2 = np.asarray(corners1, dtype=np.float64), np.asarray(corners2, dtype=np.float64)
    rect1 = [(corners1[i, 0], corners1[i, 2]) for i in range(3, -1, -1)]
    rect2 = [(corners2[i, 0], corners2[i, 2]) for i in range(3, -1, -1)]
    area1 = poly_area(np.array(rect1)[:, 0], np.array(rect1)[:, 1])
    area2 = poly_area(np.array(rect2)[:, 0], np.array(rect2)[:, 1])
    _, inter_area = convex_hull_intersection(rect1, rect2)
    iou_2d = inter_area / (area1 + area2 - inter_area)
    ymax = min(corners1[0, 1], corners2[0, 1])
    ymin = max(corners1[4, 1], corners2[4, 1])
    inter_vol = inter_area * max(0.0, ymax - ymin)
    vol1, vol2 = box3d_vol(corners1), box3d_vol(corners2)
    iou = inter_vol / (vol1 + vol2 - inter_vol)
    return iou, iou_2d


# ---- callers' helpers (roi_seg_box3d_dataset.py) ------------------------------------------------------------------
def roty(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def get_3d_box(box_size, heading_angle, center):
    """roi_seg_box3d_dataset.py:84-100 -> (8,3) corners in upright camera coordinates."""
    R = roty(heading_angle)
    l, w, h = box_size
    x_corners = [l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2]
    y_corners = [h / 2, h / 2, h / 2, h / 2, -h / 2, -h / 2, -h / 2, -h / 2]
    z_corners = [w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2]
    corners_3d = np.dot(R, np.vstack([x_corners, y_corners, z_corners]))
    corners_3d[0, :] += center[0]
    corners_3d[1, :] += center[1]
    corners_3d[2, :] += center[2]
    return np.transpose(corners_3d)


def class2angle(pred_cls, residual, num_class, to_label_format=True):
    """roi_seg_box3d_dataset.py:64-71."""
    angle = pred_cls * (2 * np.pi / float(num_class)) + residual
    if to_label_format and angle > np.pi:
        angle = angle - 2 * np.pi
    return angle


def class2size(pred_cls, residual):
    """roi_seg_box3d_dataset.py:79-82."""
    return MEAN_DIMS_ARR[pred_cls].astype(np.float64) + residual


def get_box3d_iou(center_A, box_size_A, heading_angle_A, center_B, box_size_B, heading_angle_B):
    """box_pc_fit_dataset.py:35-39."""
    return box3d_iou(get_3d_box(box_size_A, heading_angle_A, center_A), get_3d_box(box_size_B, heading_angle_B, center_B))


def compute_box3d_iou(center_pred, heading_logits, heading_residuals, size_logits, size_residuals, center_label,
                      heading_class_label, heading_residual_label, size_class_label, size_residual_label):
    """roi_seg_box3d_dataset.py:102-139 -> (iou2ds (B,), iou3ds (B,)) float32."""
    batch_size = heading_logits.shape[0]
    heading_class = np.argmax(heading_logits, 1)
    heading_residual = np.array([heading_residuals[i, heading_class[i]] for i in range(batch_size)])
    size_class = np.argmax(size_logits, 1)
    size_residual = np.vstack([size_residuals[i, size_class[i], :] for i in range(batch_size)])
    iou2d_list, iou3d_list = [], []
    for i in range(batch_size):
        heading_angle = class2angle(heading_class[i], heading_residual[i], NUM_HEADING_BIN)
        box_size = class2size(size_class[i], size_residual[i])
        corners_3d = get_3d_box(box_size, heading_angle, center_pred[i])
        heading_angle_label = class2angle(heading_class_label[i], heading_residual_label[i], NUM_HEADING_BIN)
        box_size_label = class2size(size_class_label[i], size_residual_label[i])
        corners_3d_label = get_3d_box(box_size_label, heading_angle_label, center_label[i])
        iou_3d, iou_2d = box3d_iou(corners_3d, corners_3d_label)
        iou3d_list.append(iou_3d)
        iou2d_list.append(iou_2d)
    return np.array(iou2d_list, dtype=np.float32), np.array(iou3d_list, dtype=np.float32)
