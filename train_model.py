#!/usr/bin/env python
"""Example driver -- the Python 3 counterpart of the reference's train_model.py (same calls, same configuration.cfg
keys); everything below the imports runs on the B200 kernels.

    python train_model.py [configuration.cfg] [--train]
"""
import configparser
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "sub-cortical_segmentation_b200"))

from cnn_cort.base import generate_training_set, load_data, load_test_names, test_scan  # noqa: E402
from cnn_cort.load_options import load_options, print_options  # noqa: E402
from cnn_cort.nets import build_model  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    cfg_path = args[0] if args else os.path.join(os.getcwd(), "configuration.cfg")
    user_config = configparser.RawConfigParser()
    if not user_config.read(cfg_path):
        raise SystemExit("cannot read %s" % cfg_path)
    options = load_options(user_config)
    print_options(options)
    weights_path = os.path.join(os.getcwd(), "nets")

    if "--train" in sys.argv:
        x_axial, x_cor, x_sag, y, x_atlas, names = load_data(options)
        x_train_axial, x_train_cor, x_train_sag, x_train_atlas, y_train = generate_training_set(
            x_axial, x_cor, x_sag, x_atlas, y, options)
        net = build_model(weights_path, options)
        net.fit({'in1': x_train_axial, 'in2': x_train_cor, 'in3': x_train_sag, 'in4': x_train_atlas}, y_train)

    t1_test_paths, folder_names = load_test_names(options)
    options['net_verbose'] = 0
    net = build_model(weights_path, options)
    for t1, current_scan in zip(t1_test_paths, folder_names):
        t = test_scan(net, t1, options)
        print("    -->  tested subject :", current_scan, "(elapsed time:", t, "min.)")


if __name__ == "__main__":
    main()
