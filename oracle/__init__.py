"""CPU oracle for the voxelwise hot path of sergivalverde/sub-cortical_segmentation.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline, never as the thing shipped.  The product
path (``sub-cortical_segmentation_b200/``) never imports this package and fails
loudly if its CUDA library is missing.

Pinning status
--------------
* gather (``oracle/gather.py``): PINNED.  ``tests/golden/make_golden.py``
  executes the reference's own ``get_patches`` / ``get_mask_voxels`` /
  ``generate_training_set`` source text (read from ``/root/reference`` at
  fixture-generation time, with the three mechanical py2->py3 token fixes the
  script lists) and stores its outputs in ``tests/golden/gather_golden.npz``;
  ``tests/test_oracle_gather.py`` checks this restatement against them.
* network (``oracle/network.py``): PARITY UNPINNED.  The arithmetic of the
  reference lives in Theano 0.9.0 / Lasagne 0.2.dev1 / nolearn 0.6.0
  (``requirements.txt:9,16,32``), none of which is present or installable here,
  and the reference ships no tests or golden vectors.  The restatement follows
  ``cnn_cort/nets.py:159-231`` plus the documented Lasagne layer semantics
   and is anchored by: the committed weight pickle loading into
  exactly the declared shapes, the flatten size 540, the one-hot-atlas ->
  class j+1 behaviour, and the dense-dilated == patchwise identity.
"""
