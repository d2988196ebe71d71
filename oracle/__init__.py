"""CPU oracle for the Sequitr hot path -- TEST INFRASTRUCTURE ONLY.

This package is the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``sequitr_b200/`` imports it.

Parity status ("pins"):

* W1 / W2 weight maps (``weightmap_oracle``): restated in NumPy/SciPy following
  ``/root/reference/sequitr/pipeline.py:455-571`` and PINNED against the
  reference's own code, imported from ``/root/reference`` by
  ``scripts/make_golden.py`` (fixtures under ``tests/golden/``).
* L1 label-and-localise (``centroid_oracle``): restated around the same SciPy
  calls as ``/root/reference/sequitr/utils.py:505-578`` (the file itself is
  Python-2 only and cannot be imported).  The reference ships no tests or
  golden vectors for it, so this part is "parity unpinned" by reference
  fixtures; it is pinned to SciPy's behaviour (the reference's third-party
  dependency, version unpinned in ``README.md:27``).
* UNet (``unet_oracle`` + ``unet_ref.c``): the reference ships only the abstract
  topology (``networks/unet.py:224-322``); the layer arithmetic lives in
  TensorFlow 1.x (absent).  "Parity unpinned": restated with TF semantics
  (NHWC/HWIO, SAME padding, concat order ``[upsampled, skip]``) in torch-CPU fp32
  and in plain C with a fixed accumulation order.
"""
