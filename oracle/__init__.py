"""CPU oracle for the CoMA-UNet hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.
The product (``coma_unet_b200``) never imports it.

What it restates (plain PyTorch, fp32, NCDHW, runs on CPU):

* ``monai_blocks``  -- MONAI ``Convolution``/``ADN`` and ``attentionunet.{ConvBlock,
  UpConv,AttentionBlock,AttentionLayer}``.  MONAI is not vendored in the reference,
  not pinned (no requirements file) and not installed here, so this part is restated
  from MONAI's published source layout (>=1.0).  **parity unpinned** for these blocks.
* ``cond_conv``     -- ``CondConv.CondConvolution`` / ``CondConvBlock``.  The source is
  missing from the reference (only call sites exist,
  attn_unet_data_parallel.py:126,285-306,318-325,354-367); the behaviour is specified
  by this repo.  **parity unpinned**.
* ``model``         -- attn_unet_data_parallel.py:120-693 (module tree, tuple plumbing,
  prompt/modulator logic, return conventions).  **pinned**: tests/golden/ holds outputs
  of the reference's own classes (imported from /root/reference with the two packages
  above injected as ``monai`` / ``CondConv``), see tests/golden/make_golden.py.
* ``criterions``    -- criterions.py:124-211,485-644 (RoiMSE, RnCLoss,
  GenerativeContrastiveLoss).  **pinned** against the reference file itself, run on
  CPU in this container (tests/golden/make_golden.py).
"""
