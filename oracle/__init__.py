"""CPU oracle for the StyleMC hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU with plain torch ops (fp32 or fp64), the algorithm of the
reference's S-space synthesis + CLIP-loss path.  Nothing under ``stylemc_b200/`` may import it:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and there only as the checker / the timed CPU baseline.

Pinning status (see DESIGN.md "Oracle"):

* ``fir``, ``act``, ``conv.conv2d_resample``, ``conv.fma``, ``synthesis.block_forward`` /
  ``generate_image`` / ``get_styles`` / ``split_ws`` / ``get_temp_shapes``,
  ``direction.unprocess`` and ``direction.CLIPLoss`` are PINNED: ``oracle/pin_reference.py``
  runs the reference's own code from ``/root/reference`` (``torch_utils.ops.* impl='ref'``,
  ``utils.py``, ``find_direction.unprocess``, ``clip_loss.CLIPLoss``) on seeded inputs, checks
  the restatement against it and commits the reference's outputs as ``tests/golden/*.npz``.
* ``conv.modulated_conv2d`` and the ``synthesis`` layer classes restate NVlabs
  stylegan2-ada-pytorch ``training/networks.py`` (absent from the reference tree; it only
  arrives inside network pickles, persistence ``_version = 6``): PARITY UNPINNED against
  upstream source, anchored on the reference's call sites (``utils.py:13-53``), on
  ``legacy.py:173-202`` (parameter names/shapes) and on the in-tree e4e ``ModulatedConv2d``
  analogue.
* ``vit`` restates openai/CLIP ``clip/model.py`` (not vendored, unpinned git install,
  ``README.md:13``): PARITY UNPINNED against upstream source, cross-checked by weight copy
  against ``transformers.CLIPModel`` (same architecture) in ``pin_reference.py``.
"""
