"""cnn_cort -- drop-in Python 3 mirror of the reference package's hot-path API
(sergivalverde/sub-cortical_segmentation, ``cnn_cort/``), backed by hand-written
sm_100a CUDA kernels in ``libsubcort_b200.so`` (C-ABI: ``include/subcort_b200.h``).

    from cnn_cort.load_options import load_options
    from cnn_cort.base import load_data, generate_training_set, load_test_names, test_scan
    from cnn_cort.nets import build_model

There is no CPU fallback: importing works anywhere, but the first call that needs the
network or the gather raises if the CUDA library or a B200 is missing.
"""
__all__ = ["base", "nets", "load_options", "nifti", "synthetic", "parallel"]
