"""Drop-in for the overlap half of pcdet/datasets/kitti/kitti_object_eval_python (rotate_iou.py, eval.py:80-155, 340-414)."""
