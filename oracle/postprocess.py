"""Oracle: connected-component post-processing (``cnn_cort/base.py:460-480``), restated with scipy.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Pinned by construction: the statements are the reference's own (scipy's
``ndimage.label`` / ``labeled_comprehension`` are the reference's dependencies, scipy 0.19 -> 1.x keeps their semantics).
"""
import numpy as np
from scipy import ndimage


def post_process_segmentation(atlas_mask, input_mask):
    """Per label 1..14: 6-connected components of ``input_mask == l`` (``ndimage.label``), voxels inside ``atlas_mask`` per
    component (``labeled_comprehension`` over ``np.unique(labels)``, background first), ``np.argmax`` of that list, paint
    ``labels == argmax`` with l (base.py:466-478).  Quirk Q12 included: argmax 0 selects the background of the class."""
    filtered_mask = np.zeros_like(input_mask)
    for l in range(1, 15):
        th_label = input_mask == l
        labels, num_labels = ndimage.label(th_label)
        label_list = np.unique(labels)
        num_elements = ndimage.labeled_comprehension(np.logical_and(th_label, atlas_mask), labels, label_list, np.sum, float, 0)
        argmax = np.argmax(num_elements)
        current_voxels = np.stack(np.where(labels == argmax), axis=1)
        filtered_mask[current_voxels[:, 0], current_voxels[:, 1], current_voxels[:, 2]] = l
    return filtered_mask
