"""Class-hierarchy tables, restated literally from the reference (oracle side).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The reference hard-codes the 2-level hierarchy three times; this file restates
two of them as plain Python lists so the product's independently *derived*
tables (wlseg/hierarchy.py) can be checked against them:

* loss side:  code/estimator/define_losses_hierarchical.py:38-93
* model side: code/models/resnet50_extended_model_hierarchical.py:95-111
* head widths: code/models/resnet50_extended_model_hierarchical.py:81-83
"""

# Weak (Open Images) class ids, code/input_pipelines/open_images/input_subset_bboxes_v2.py:38-53
WEAK_CLASS_NAMES = [
    'bicycle', 'bus', 'car', 'motorcycle', 'train', 'truck',
    'human', 'man', 'woman', 'boy', 'girl',
    'traffic light', 'traffic sign', 'stop sign', 'void']

TABLES = {
    'cityscapes': dict(
        # define_losses_hierarchical.py:75-93
        cid_l1_vehicle=12,
        cid_l1_human=11,
        per_pixel_cids2l1_cids=[0, 1, 2, 3, 4, 5, 6, 7, 8, 9,
                                10, 11, 11, 12, 12, 12, 12, 12, 12, 13],
        per_bbox_cids2l1_cids=[12, 12, 12, 12, 12, 12, 11, 11, 11, 11,
                               11, 13, 13, 13, 13],
        per_pixel_cids2vehicle_cids=[6, 6, 6, 6, 6, 6, 6, 6, 6, 6,
                                     6, 6, 6, 0, 1, 2, 3, 4, 5, 6],
        per_bbox_cids2vehicle_cids=[5, 2, 0, 4, 3, 1, 6, 6, 6, 6, 6, 6, 6, 6, 6],
        per_pixel_cids2human_cids=[2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
                                   2, 0, 1, 2, 2, 2, 2, 2, 2, 2],
        per_bbox_cids2human_cids=[2, 2, 2, 2, 2, 2, 0, 0, 0, 0, 0, 2, 2, 2, 2],
        # resnet50_extended_model_hierarchical.py:106-111
        l1_cids2common_cids=[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 19],
        l2_vehicle_cids2common_cids=[13, 14, 15, 16, 17, 18, 19],
        l2_human_cids2common_cids=[11, 12, 19],
        # resnet50_extended_model_hierarchical.py:81-83
        head_widths=(14, 7, 3),
        num_classes=20,
    ),
    'vistas': dict(
        # define_losses_hierarchical.py:38-74
        cid_l1_vehicle=49,
        cid_l1_human=19,
        per_pixel_cids2l1_cids=[
            0, 1, 2, 3, 4, 5, 6, 7, 8, 9,
            10, 11, 12, 13, 14, 15, 16, 17, 18, 19,
            19, 19, 19, 20, 21, 22, 23, 24, 25, 26,
            27, 28, 29, 30, 31, 32, 33, 34, 35, 36,
            37, 38, 39, 40, 41, 42, 43, 44, 45, 46,
            47, 48, 49, 49, 49, 49, 49, 49, 49, 49,
            49, 49, 49, 50, 51, 52],
        per_bbox_cids2l1_cids=[49, 49, 49, 49, 49, 49, 19, 19, 19, 19,
                               19, 52, 52, 52, 52],
        per_pixel_cids2vehicle_cids=[
            11, 11, 11, 11, 11, 11, 11, 11, 11, 11,
            11, 11, 11, 11, 11, 11, 11, 11, 11, 11,
            11, 11, 11, 11, 11, 11, 11, 11, 11, 11,
            11, 11, 11, 11, 11, 11, 11, 11, 11, 11,
            11, 11, 11, 11, 11, 11, 11, 11, 11, 11,
            11, 11, 0, 1, 2, 3, 4, 5, 6, 7,
            8, 9, 10, 11, 11, 11],
        per_bbox_cids2vehicle_cids=[0, 2, 3, 5, 6, 9, 11, 11, 11, 11, 11, 11, 11, 11, 11],
        per_pixel_cids2human_cids=[
            4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
            4, 4, 4, 4, 4, 4, 4, 4, 4, 0,
            1, 2, 3, 4, 4, 4, 4, 4, 4, 4,
            4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
            4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
            4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
            4, 4, 4, 4, 4, 4],
        per_bbox_cids2human_cids=[4, 4, 4, 4, 4, 4, 0, 0, 0, 0, 0, 4, 4, 4, 4],
        # resnet50_extended_model_hierarchical.py:96-105
        l1_cids2common_cids=[
            0, 1, 2, 3, 4, 5, 6, 7, 8, 9,
            10, 11, 12, 13, 14, 15, 16, 17, 18, 19,
            23, 24, 25, 26, 27, 28, 29, 30, 31, 32,
            33, 34, 35, 36, 37, 38, 39, 40, 41, 42,
            43, 44, 45, 46, 47, 48, 49, 50, 51, 52,
            63, 64, 65],
        l2_vehicle_cids2common_cids=[52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 65],
        l2_human_cids2common_cids=[19, 20, 21, 22, 65],
        head_widths=(53, 12, 5),
        num_classes=66,
    ),
}
