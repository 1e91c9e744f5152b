"""CPU oracle for the HaMeR per-crop hand-reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``hamer_yolo_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline.

It is a plain fp32 PyTorch-on-CPU / numpy restatement of the reference's algorithm for the path
(every function cites the reference file:line it follows) and does not read ``/root/reference`` at
run time, so it travels to the GPU box.

Pinning status
--------------
* backbone + decoder head + rot6d + projection + forward_step assembly: PINNED against the
  reference's own modules imported from ``/root/reference`` in the authoring container
  (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``).
* crop / warp: PINNED against OpenCV's ``cv2.warpAffine`` / ``cv2.getAffineTransform`` (the
  third-party functions the reference calls, ``hamer/hamer/datasets/utils.py:127,359``) which are
  installed in this image (``tests/test_oracle_crop.py``).
* MANO layer (smplx==0.1.28 ``lbs`` - third-party, pinned in ``setup.py:15``, source absent from the
  reference tree): restated from the published algorithm; cross-checked bit-for-bit against the
  reference's independent in-tree implementation of the same maths
  (``rootnet/KeypointFusion/manopth/manopth/manolayer.py:112-276``) on the real ``MANO_RIGHT.pkl``
  when ``/root/reference`` is present, and pinned by ``tests/golden/mano_lbs_real.npz`` (outputs of
  that in-tree implementation).  The reference holds no golden vectors of its own for this path, so
  at the smplx boundary the parity is "pinned to manopth", not to smplx itself.
"""
