"""SUN-RGBD constants shared by the host mirror and the synthetic generators.

Reference: sunrgbd/sunrgbd_detection/roi_seg_box3d_dataset.py:18-33 (type2class order,
type_mean_size in (l,w,h), NUM_HEADING_BIN=12, NUM_SIZE_CLUSTER=10, NUM_CLASS=10) and
models/model_util.py:15-54 (KITTI set kept for the F-PointNet helpers, NUM_OBJECT_POINT=512).
"""
import numpy as np

type2class = {'bed': 0, 'table': 1, 'sofa': 2, 'chair': 3, 'toilet': 4, 'desk': 5,
              'dresser': 6, 'night_stand': 7, 'bookshelf': 8, 'bathtub': 9}
class2type = {type2class[t]: t for t in type2class}
type2onehotclass = dict(type2class)
type_mean_size = {'bathtub': np.array([0.765840, 1.398258, 0.472728]),
                  'bed': np.array([2.114256, 1.620300, 0.927272]),
                  'bookshelf': np.array([0.404671, 1.071108, 1.688889]),
                  'chair': np.array([0.591958, 0.552978, 0.827272]),
                  'desk': np.array([0.695190, 1.346299, 0.736364]),
                  'dresser': np.array([0.528526, 1.002642, 1.172878]),
                  'night_stand': np.array([0.500618, 0.632163, 0.683424]),
                  'sofa': np.array([0.923508, 1.867419, 0.845495]),
                  'table': np.array([0.791118, 1.279516, 0.718182]),
                  'toilet': np.array([0.699104, 0.454178, 0.756250])}
NUM_HEADING_BIN = 12
NUM_SIZE_CLUSTER = 10
NUM_CLASS = 10
NUM_OBJECT_POINT = 512

MEAN_DIMS_ARR = np.zeros((NUM_SIZE_CLUSTER, 3))
for _i in range(NUM_SIZE_CLUSTER):
    MEAN_DIMS_ARR[_i, :] = type_mean_size[class2type[_i]]
ORIENT_ANCHORS = np.arange(0, 2 * np.pi, 2 * np.pi / NUM_HEADING_BIN)

# KITTI set of models/model_util.py:15-34 (parse_output_to_tensors defaults there).
KITTI_NUM_SIZE_CLUSTER = 8
g_type2class = {'Car': 0, 'Van': 1, 'Truck': 2, 'Pedestrian': 3,
                'Person_sitting': 4, 'Cyclist': 5, 'Tram': 6, 'Misc': 7}
g_class2type = {g_type2class[t]: t for t in g_type2class}
g_type_mean_size = {'Car': np.array([3.88311640418, 1.62856739989, 1.52563191462]),
                    'Van': np.array([5.06763659, 1.9007158, 2.20532825]),
                    'Truck': np.array([10.13586957, 2.58549199, 3.2520595]),
                    'Pedestrian': np.array([0.84422524, 0.66068622, 1.76255119]),
                    'Person_sitting': np.array([0.80057803, 0.5983815, 1.27450867]),
                    'Cyclist': np.array([1.76282397, 0.59706367, 1.73698127]),
                    'Tram': np.array([16.17150617, 2.53246914, 3.53079012]),
                    'Misc': np.array([3.64300781, 1.54298177, 1.92320313])}
g_mean_size_arr = np.zeros((KITTI_NUM_SIZE_CLUSTER, 3))
for _i in range(KITTI_NUM_SIZE_CLUSTER):
    g_mean_size_arr[_i, :] = g_type_mean_size[g_class2type[_i]]

BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default epsilon (models/tf_util.py:1660)
